// Column scales of the bootstrap cross-correlations in one pass.
//
// compute.xcorr z-scores the gathered rows X[inds] of every cell (ddof = 1,
// pyls/compute.py:84) before the product; restated on multiplicities the
// scale of column b in cell g of resample r is
//     1 / ((n - 1) sigma'),  sigma'^2 = (s2 - s1^2 / n) / (n - 1),
//     s1 = sum_u c[u] Xn[u,b],  s2 = sum_u c[u] Xn[u,b]^2,
// c[u] = how often row u of the cell was drawn.  The generic route computes s1
// and s2 as two more (tiny) GEMMs and finishes them in colscale_kernel: three
// passes over (n J) x B matrices.  Here a CTA keeps one cell's rows of a column
// tile of Xn in shared memory, runs through its share of the resamples --
// multiplicities from the index table, only the rows that were drawn -- and
// writes the scale directly: X is read once, the output written once.
//
// Bootstrap tables resample subjects inside their group and keep conditions in
// place (pyls/base.py:134-143), so the sources of a cell's rows lie in the cell.
#include "common.cuh"

namespace plsb {
namespace {

constexpr int CS_THREADS = 128;
constexpr int CS_BATCH = 8;   // resamples per barrier round

template <int W>   // columns per thread
__global__ void __launch_bounds__(CS_THREADS)
colstats_kernel(const double *__restrict__ Xn, long long ldx, int B, const int32_t *__restrict__ idx,
                int S, int J, const int *__restrict__ cell_start, int n_res, int res_per_cta,
                double *__restrict__ scale) {
  extern __shared__ __align__(16) double sm[];
  const int g = blockIdx.y, tid = threadIdx.x;
  const int r0 = cell_start[g], ng = cell_start[g + 1] - r0;
  const int col0 = blockIdx.x * (CS_THREADS * W);
  double *Xs = sm;                                         // ng x (CS_THREADS * W)
  double *cw = Xs + (size_t)ng * CS_THREADS * W;           // CS_BATCH x ng compact: counts
  int *cnt = reinterpret_cast<int *>(cw + CS_BATCH * ng);  // CS_BATCH x ng
  int *lst = cnt + CS_BATCH * ng;                          // CS_BATCH x ng compact: rows
  int *nlst = lst + CS_BATCH * ng;                         // CS_BATCH
  for (int e = tid; e < ng * CS_THREADS * W; e += CS_THREADS) {
    const int u = e / (CS_THREADS * W), c = e - u * (CS_THREADS * W);
    Xs[e] = Xn[(size_t)(r0 + u) * ldx + col0 + c];          // columns >= B are zero padding
  }
  const int first = blockIdx.z * res_per_cta, last = min(first + res_per_cta, n_res);
  const double n = (double)ng, inv_n = 1.0 / n, inv_nm1 = 1.0 / (n - 1.0);
  for (int rb = first; rb < last; rb += CS_BATCH) {
    const int nb = min(CS_BATCH, last - rb);
    __syncthreads();
    for (int e = tid; e < nb * ng; e += CS_THREADS) cnt[e] = 0;
    __syncthreads();
    for (int e = tid; e < nb * ng; e += CS_THREADS) {
      const int k = e / ng, s = e - k * ng;
      const int u = idx[(size_t)(rb + k) * S + r0 + s] - r0;
      if (u >= 0 && u < ng) atomicAdd(&cnt[k * ng + u], 1);
    }
    __syncthreads();
    // compact list of the rows that were drawn (tile offset, count): a warp per resample
    for (int k = tid >> 5; k < nb; k += CS_THREADS / 32) {
      const int lane = tid & 31;
      int m = 0;
      for (int u0 = 0; u0 < ng; u0 += 32) {
        const int u = u0 + lane;
        const int c = u < ng ? cnt[k * ng + u] : 0;
        const unsigned mask = __ballot_sync(0xffffffffu, c != 0);
        if (c) {
          const int pos = k * ng + m + __popc(mask & ((1u << lane) - 1u));
          lst[pos] = u * (CS_THREADS * W);
          cw[pos] = (double)c;
        }
        m += __popc(mask);
      }
      if (lane == 0) nlst[k] = m;
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
      double s1[W], s2[W];
#pragma unroll
      for (int w = 0; w < W; ++w) s1[w] = s2[w] = 0.0;
      const int m = nlst[k];
      for (int i = 0; i < m; ++i) {
        const double c = cw[k * ng + i];
        const double *xr = Xs + lst[k * ng + i] + tid;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const double x = xr[w * CS_THREADS], cx = c * x;
          s1[w] += cx;
          s2[w] += cx * x;
        }
      }
      double *out = scale + ((size_t)(rb + k) * J + g) * ldx + col0 + tid;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int b = col0 + tid + w * CS_THREADS;
        const double var = (s2[w] - s1[w] * s1[w] * inv_n) * inv_nm1;
        out[w * CS_THREADS] = b < B ? rsqrt(var) * inv_nm1 : 0.0;
      }
    }
  }
}

}  // namespace

// scale (n_res * J rows, ldx) for bootstrap tables idx (n_res, S); returns PLSB_OK and
// sets *done = false (nothing launched) when a cell does not fit the shared-memory tile
int launch_colstats(plsb_ctx *h, const int32_t *idx, int n_res, double *scale, bool *done,
                    cudaStream_t st) {
  const Layout &l = h->lay;
  *done = false;
  int max_ng = 0;
  for (int j = 0; j < l.J; ++j) max_ng = std::max(max_ng, l.cell_start[j + 1] - l.cell_start[j]);
  if (n_res <= 0 || !tune_int("PLSB_COLSTATS", 1)) return PLSB_OK;
  constexpr int W = 1;
  const int cols = CS_THREADS * W;
  const size_t smem = sizeof(double) * ((size_t)max_ng * cols + (size_t)CS_BATCH * max_ng) +
                      sizeof(int) * ((size_t)2 * CS_BATCH * max_ng + CS_BATCH);
  if (smem > 100 * 1024 || l.ldx % cols != 0) return PLSB_OK;
  KernelTimer kt(h, KC_STATS, st);
  const int n_ct = l.ldx / cols;
  // enough CTAs for a few waves, every CTA re-using its X tile for many resamples
  int z = std::max(1, std::min(n_res / 64, (h->sm_count * 16) / std::max(1, n_ct * l.J)));
  const int per = cdiv(n_res, z);
  z = cdiv(n_res, per);
  PLSB_CUDA(cudaFuncSetAttribute(colstats_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  dim3 grid(n_ct, l.J, z);
  colstats_kernel<W><<<grid, CS_THREADS, smem, st>>>(h->Xglob.as<double>(), l.ldx, l.B, idx, l.S, l.J,
                                                    h->d_cell_start, n_res, per, scale);
  PLSB_LAUNCHED(h);
  *done = true;
  return PLSB_OK;
}

}  // namespace plsb
