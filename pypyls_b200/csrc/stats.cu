// Reductions over the resample axis: permutation p-values, bootstrap
// percentile intervals and bootstrap ratios.
//
//   pvals        compute.perm_sig   (pyls/compute.py:154-181)  strict '>'
//   percentile   compute.boot_ci    (pyls/compute.py:184-209)  numpy 'linear'
//                method incl. numpy's two-sided lerp, so results are bit-equal
//                to np.percentile on the same values
//   boot_ratio   compute.boot_rel   (pyls/compute.py:212-237)
#include <math.h>

#include "common.cuh"

namespace plsb {
namespace {

__global__ void pvals_kernel(const double *__restrict__ dperm, int count, int L,
                             const double *__restrict__ dorig, double *__restrict__ pvals) {
  __shared__ int red[32];
  const int l = blockIdx.x;
  const double o = dorig[l];
  int c = 0;
  for (int r = threadIdx.x; r < count; r += blockDim.x) c += dperm[(size_t)r * L + l] > o;
  for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    pvals[l] = ((double)t + 1.0) / ((double)count + 1.0);
  }
}

// ---- order statistics by selection ------------------------------------------------
// np.percentile needs two order statistics per quantile, not a sorted series.  A series
// (one (k, l) entry of the bootstrap distribution over all resamples) is read ONCE,
// coalesced, into registers as order-preserving 64-bit keys; the k-th smallest key is then
// found by a binary search over the key bits (most significant first: "how many keys are
// below prefix | bit?"), both quantiles at once: 64 counting passes over registers, one
// CTA barrier each, O(n) work and no data movement -- against O(n log^2 n) compare-
// exchanges through shared memory for the bitonic sort this replaces.  The interpolation
// reproduces numpy's _lerp bit for bit.
constexpr int SEL_THREADS = 512;

__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct QuantilePlan {
  int prev, next;   // order statistics (0-based) the quantile interpolates between
  double gamma;
};
__device__ __forceinline__ QuantilePlan plan_quantile(int n, double qpct) {
  // numpy 'linear': virtual index (n - 1) q, neighbours floor / floor + 1, clipped
  const double vi = (double)(n - 1) * (qpct / 100.0);
  QuantilePlan p;
  long long prev = (long long)floor(vi), next = prev + 1;
  if (vi >= (double)(n - 1)) prev = next = n - 1;
  if (vi < 0.0) prev = next = 0;
  p.prev = (int)prev;
  p.next = (int)next;
  p.gamma = vi - floor(vi);
  return p;
}
__device__ __forceinline__ double np_lerp(double a, double b, double gamma) {
  // numpy's _lerp, with explicit rounding of every product so that the compiler cannot
  // contract mul+add into an FMA (np.percentile does not)
  const double diff = __dsub_rn(b, a);
  double res = __dadd_rn(a, __dmul_rn(diff, gamma));
  if (gamma >= 0.5) res = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
  return res;
}

// KPT > 0: the series lives in registers (KPT keys per thread, n <= KPT * SEL_THREADS);
// KPT == 0: any length, keys re-read from global memory in every pass.
template <int KPT>
__global__ void __launch_bounds__(SEL_THREADS)
percentile_select_kernel(const double *__restrict__ src, long long ld, int n, int n_series,
                         double qlo, double qhi, double *__restrict__ lo,
                         double *__restrict__ hi) {
  constexpr int NW = SEL_THREADS / 32;
  __shared__ unsigned s_cnt[2][NW][2];
  __shared__ unsigned long long s_min[NW][2];
  __shared__ int s_nan;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const QuantilePlan pa = plan_quantile(n, qlo), pb = plan_quantile(n, qhi);
  constexpr int R = KPT > 0 ? KPT : 1;

  for (int j = blockIdx.x; j < n_series; j += gridDim.x) {
    const double *series = src + (size_t)j * ld;
    unsigned long long key[R];
    int has_nan = 0;
    if (KPT > 0) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int e = tid + i * SEL_THREADS;
        key[i] = ~0ull;   // padding sorts after every value
        if (e < n) {
          const double v = series[e];
          has_nan |= (v != v);
          key[i] = key_of(v);
        }
      }
    } else {
      for (int e = tid; e < n; e += SEL_THREADS) {
        const double v = series[e];
        has_nan |= (v != v);
      }
    }
    if (tid == 0) s_nan = 0;
    __syncthreads();
    if (has_nan) s_nan = 1;

    // binary search of the key bits for the two "prev" order statistics
    unsigned long long pre_a = 0, pre_b = 0;
    for (int bit = 63; bit >= 0; --bit) {
      const unsigned long long ca = pre_a | (1ull << bit), cb = pre_b | (1ull << bit);
      unsigned na = 0, nb = 0;
      if (KPT > 0) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          na += key[i] < ca;
          nb += key[i] < cb;
        }
      } else {
        for (int e = tid; e < n; e += SEL_THREADS) {
          const unsigned long long k = key_of(series[e]);
          na += k < ca;
          nb += k < cb;
        }
      }
      na = __reduce_add_sync(0xffffffffu, na);
      nb = __reduce_add_sync(0xffffffffu, nb);
      const int buf = bit & 1;
      if (lane == 0) {
        s_cnt[buf][warp][0] = na;
        s_cnt[buf][warp][1] = nb;
      }
      __syncthreads();
      unsigned ta = 0, tb = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        ta += s_cnt[buf][w][0];
        tb += s_cnt[buf][w][1];
      }
      // keys below the candidate <= k  <=>  the k-th smallest key is >= the candidate
      if ((int)ta <= pa.prev) pre_a = ca;
      if ((int)tb <= pb.prev) pre_b = cb;
    }

    // the following order statistic: the same value if it is repeated, else the smallest
    // key above it
    unsigned la = 0, lb = 0;
    unsigned long long ma = ~0ull, mb = ~0ull;
    if (KPT > 0) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        la += key[i] <= pre_a;
        lb += key[i] <= pre_b;
        if (key[i] > pre_a && key[i] < ma) ma = key[i];
        if (key[i] > pre_b && key[i] < mb) mb = key[i];
      }
    } else {
      for (int e = tid; e < n; e += SEL_THREADS) {
        const unsigned long long k = key_of(series[e]);
        la += k <= pre_a;
        lb += k <= pre_b;
        if (k > pre_a && k < ma) ma = k;
        if (k > pre_b && k < mb) mb = k;
      }
    }
    la = __reduce_add_sync(0xffffffffu, la);
    lb = __reduce_add_sync(0xffffffffu, lb);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const unsigned long long oa = __shfl_xor_sync(0xffffffffu, ma, s);
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, mb, s);
      ma = oa < ma ? oa : ma;
      mb = ob < mb ? ob : mb;
    }
    __syncthreads();   // the last counting pass has been read by everyone
    if (lane == 0) {
      s_cnt[0][warp][0] = la;
      s_cnt[0][warp][1] = lb;
      s_min[warp][0] = ma;
      s_min[warp][1] = mb;
    }
    __syncthreads();
    if (tid < 2) {
      const QuantilePlan &p = tid == 0 ? pa : pb;
      const unsigned long long pre = tid == 0 ? pre_a : pre_b;
      unsigned le = 0;
      unsigned long long mn = ~0ull;
      for (int w = 0; w < NW; ++w) {
        le += s_cnt[0][w][tid];
        mn = s_min[w][tid] < mn ? s_min[w][tid] : mn;
      }
      const unsigned long long nxt = (p.next == p.prev || (int)le > p.next) ? pre : mn;
      double res = np_lerp(value_of(pre), value_of(nxt), p.gamma);
      if (s_nan) res = __longlong_as_double(0x7ff8000000000000ll);   // like np.percentile
      (tid == 0 ? lo : hi)[j] = res;
    }
    __syncthreads();
  }
}

template <int KPT>
void launch_select(const double *src, long long ld, int n, int n_series, double qlo, double qhi,
                   double *lo, double *hi, int blocks, cudaStream_t st) {
  percentile_select_kernel<KPT><<<blocks, SEL_THREADS, 0, st>>>(src, ld, n, n_series, qlo, qhi,
                                                                 lo, hi);
}

__global__ void boot_ratio_kernel(const double *__restrict__ bs, const double *__restrict__ usum,
                                  const double *__restrict__ usq, size_t n_elem, int n_boot,
                                  int add_orig, double *__restrict__ bsr,
                                  double *__restrict__ se) {
  const double n = (double)(n_boot + (add_orig ? 1 : 0));
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n_elem;
       e += (size_t)gridDim.x * blockDim.x) {
    const double o = bs[e];
    double s1 = usum[e], s2 = usq[e];
    if (add_orig) {
      s1 += o;
      s2 += o * o;
    }
    const double sum2 = (s1 * s1) / n;
    const double e_se = sqrt(fabs(s2 - sum2) / (n - 1.0));
    se[e] = e_se;
    bsr[e] = o / e_se;
  }
}

}  // namespace

int launch_pvals(plsb_ctx *h, const double *dperm, int count, int L, const double *dorig,
                 double *pvals, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  pvals_kernel<<<L, 256, 0, st>>>(dperm, count, L, dorig, pvals);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

// series-major input: series j is the `count` values at series + j * ld
int launch_percentile_series(plsb_ctx *h, const double *series, long long ld, int count,
                             int n_series, double qlo, double qhi, double *lo, double *hi,
                             cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  PLSB_CHECK(count >= 1 && n_series >= 1 && ld >= count, PLSB_ERR_ARG,
             "percentile: empty input");
  const int blocks = std::min(n_series, 8 * h->sm_count);
  const int per = cdiv(count, SEL_THREADS);
  if (per <= 1) launch_select<1>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 2) launch_select<2>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 4) launch_select<4>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 8) launch_select<8>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 16) launch_select<16>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 32) launch_select<32>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else if (per <= 64) launch_select<64>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  else launch_select<0>(series, ld, count, n_series, qlo, qhi, lo, hi, blocks, st);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

// resample-major input (count, n_series), the layout the resampling drivers write: one
// tiled transpose into scratch (coalesced both ways), then the selection above
int launch_percentile(plsb_ctx *h, const double *distrib, int count, int n_series, double qlo,
                      double qhi, double *lo, double *hi, cudaStream_t st) {
  PLSB_CHECK(count >= 1 && n_series >= 1, PLSB_ERR_ARG, "percentile: empty input");
  if (n_series == 1)
    return launch_percentile_series(h, distrib, count, count, 1, qlo, qhi, lo, hi, st);
  PLSB_TRY(h->pctl.ensure(sizeof(double) * (size_t)count * n_series));
  PLSB_TRY(launch_transpose(h, distrib, count, n_series, n_series, h->pctl.as<double>(), st));
  return launch_percentile_series(h, h->pctl.as<double>(), count, count, n_series, qlo, qhi, lo,
                                  hi, st);
}

int launch_boot_ratio(plsb_ctx *h, const double *bs, const double *usum, const double *usq,
                      long long n_elem, int n_boot, int add_orig, double *bsr, double *se,
                      cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_elem <= 0) return PLSB_OK;
  const int blocks = (int)std::min<size_t>(((size_t)n_elem + 255) / 256, (size_t)h->sm_count * 16);
  boot_ratio_kernel<<<blocks, 256, 0, st>>>(bs, usum, usq, (size_t)n_elem, n_boot, add_orig, bsr,
                                            se);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
