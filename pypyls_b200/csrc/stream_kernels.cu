// Elementwise / reduction kernels around the stored cross-covariance matrices R
// (count*K rows of ldr doubles, one K x B block per resample).
//
//   finish_rowsq  rotated permutation singular values from the GEMM's row sums
//                 of squares (pyls/base.py:699-700)
//   colscale      1 / ((n-1) * sigma) of the resampled X columns from the count
//                 operand's sums and sums of squares (compute.xcorr's z-score of
//                 X[inds], pyls/compute.py:84, restated on multiplicities)
//   reduce_partials  sums the per-split partial (B, L) accumulators of accum_u
#include "common.cuh"

namespace plsb {
namespace {

__global__ void finish_rowsq_kernel(const double *__restrict__ rowsq, int n_splits, int M_pad,
                                    int n_rows, double *__restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_rows) return;
  double v = 0.0;
  for (int s = 0; s < n_splits; ++s) v += rowsq[(size_t)s * M_pad + m];
  out[m] = sqrt(v);
}

// out[m] = sqrt(sum_u T[m,u] * A[m,u]): quadratic forms a^T Kx a from T = A Kx
// (sample-space permutation singular values); one warp per row
__global__ void rowdot_sqrt_kernel(const double *__restrict__ T, long long ldt,
                                   const double *__restrict__ A, int lda, int n_cols,
                                   long long n_rows, double *__restrict__ out) {
  const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= n_rows) return;
  double v = 0.0;
  for (int u = lane; u < n_cols; u += 32) v += T[m * ldt + u] * A[m * lda + u];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[m] = sqrt(fmax(v, 0.0));
}

// rows come in groups of `rpr` per resample, the first J of a group are its cells; the
// row count of a (resample, cell) is cell_n[cell] or, with nrow, nrow[resample * J + cell]
__global__ void colscale_kernel(double *__restrict__ S1, const double *__restrict__ S2, int n_rows,
                                long long ld, int B, int J, int rpr,
                                const int *__restrict__ cell_n, const int *__restrict__ nrow) {
  // blockIdx.x: 256 columns, blockIdx.y (grid-stride): rows -- no 64-bit division per element
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ld) return;
  for (int row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const int cell = row % rpr;
    const size_t e = (size_t)row * ld + b;
    double out = 0.0;
    if (b < B && cell < J) {
      const double n = nrow ? (double)nrow[(size_t)(row / rpr) * J + cell] : (double)cell_n[cell];
      const double s1 = S1[e], s2 = S2[e];
      const double var = (s2 - s1 * s1 / n) / (n - 1.0);
      out = 1.0 / ((n - 1.0) * sqrt(var));
    }
    S1[e] = out;
  }
}

__global__ void reduce_partials_kernel(const double *__restrict__ P, int n_splits, size_t stride,
                                       size_t n, double *__restrict__ out) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < n_splits; ++s) v += P[(size_t)s * stride + e];
    out[e] += v;
  }
}

}  // namespace

int launch_reduce_partials(plsb_ctx *h, const double *P, int n_splits, size_t stride, size_t n,
                           double *out, cudaStream_t st) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 16);
  reduce_partials_kernel<<<blocks, 256, 0, st>>>(P, n_splits, stride, n, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_finish_rowsq(plsb_ctx *h, const double *rowsq, int n_splits, int M_pad, int n_rows,
                        double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_rows <= 0) return PLSB_OK;
  finish_rowsq_kernel<<<cdiv(n_rows, 256), 256, 0, st>>>(rowsq, n_splits, M_pad, n_rows, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_rowdot_sqrt(plsb_ctx *h, const double *T, long long ldt, const double *A, int lda,
                       int n_cols, long long n_rows, double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_rows <= 0) return PLSB_OK;
  const long long blocks = (n_rows * 32 + 255) / 256;
  rowdot_sqrt_kernel<<<(unsigned)blocks, 256, 0, st>>>(T, ldt, A, lda, n_cols, n_rows, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_colscale(plsb_ctx *h, double *S1, const double *S2, int n_rows, long long ld,
                    cudaStream_t st, int rows_per_resample, const int *nrow) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_rows <= 0) return PLSB_OK;
  const unsigned col_blocks = (unsigned)((ld + 255) / 256);
  // enough row blocks to fill the machine a few times over, every one striding the rows
  const unsigned row_blocks = (unsigned)std::min<long long>(
      n_rows, std::max<long long>(1, (long long)h->sm_count * 32 / col_blocks));
  const int rpr = rows_per_resample > 0 ? rows_per_resample : h->lay.J;
  colscale_kernel<<<dim3(col_blocks, row_blocks), 256, 0, st>>>(S1, S2, n_rows, ld, h->lay.B,
                                                                 h->lay.J, rpr, h->d_cell_n, nrow);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
