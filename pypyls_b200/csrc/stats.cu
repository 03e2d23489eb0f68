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

__device__ double np_quantile_linear(const double *sorted, int n, double qpct) {
  const double q = qpct / 100.0;
  const double vi = (double)(n - 1) * q;
  long long prev = (long long)floor(vi), next = prev + 1;
  if (vi >= (double)(n - 1)) prev = next = n - 1;
  if (vi < 0.0) prev = next = 0;
  const double gamma = vi - floor(vi);
  const double a = sorted[prev], b = sorted[next];
  // numpy's _lerp, with explicit rounding of every product so that the
  // compiler cannot contract mul+add into an FMA (np.percentile does not)
  const double diff = __dsub_rn(b, a);
  double res = __dadd_rn(a, __dmul_rn(diff, gamma));
  if (gamma >= 0.5) res = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
  return res;
}

// one CTA per series: gather the series, bitonic sort, interpolate
__global__ void __launch_bounds__(1024)
percentile_kernel(const double *__restrict__ distrib, int count, int n_series, int n2, double qlo,
                  double qhi, double *__restrict__ lo, double *__restrict__ hi,
                  double *__restrict__ scratch) {
  extern __shared__ __align__(16) double sm[];
  double *buf = scratch ? scratch + (size_t)blockIdx.x * n2 : sm;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = blockIdx.x; j < n_series; j += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < n2; i += nt)
      buf[i] = i < count ? distrib[(size_t)i * n_series + j] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int s = k >> 1; s > 0; s >>= 1) {
        for (int i = tid; i < n2; i += nt) {
          const int ixs = i ^ s;
          if (ixs > i) {
            const bool asc = (i & k) == 0;
            const double a = buf[i], b = buf[ixs];
            if ((a > b) == asc) {
              buf[i] = b;
              buf[ixs] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    if (tid == 0) lo[j] = np_quantile_linear(buf, count, qlo);
    if (tid == 32) hi[j] = np_quantile_linear(buf, count, qhi);
  }
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

int launch_percentile(plsb_ctx *h, const double *distrib, int count, int n_series, double qlo,
                      double qhi, double *lo, double *hi, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  PLSB_CHECK(count >= 1 && n_series >= 1, PLSB_ERR_ARG, "percentile: empty input");
  int n2 = 2;
  while (n2 < count) n2 <<= 1;
  double *scratch = nullptr;
  size_t smem = sizeof(double) * (size_t)n2;
  int blocks = n_series;
  if (smem > 160 * 1024) {
    blocks = std::min(n_series, 2 * h->sm_count);
    PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)n2 * blocks));
    scratch = h->misc.as<double>();
    smem = 0;
  }
  PLSB_CUDA(cudaFuncSetAttribute(percentile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)std::max<size_t>(smem, 1024)));
  percentile_kernel<<<blocks, 1024, smem, st>>>(distrib, count, n_series, n2, qlo, qhi, lo, hi,
                                                scratch);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
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
