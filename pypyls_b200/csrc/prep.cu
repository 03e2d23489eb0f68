// One-time data preparation and the finishing steps of the original
// decomposition.
//
//   pad_copy        X (S,B) -> zero padded (S_pad, ldx)
//   prep_cells      per-cell z-score / centring of X (compute.xcorr's
//                   normalisation, pyls/compute.py:84-87, hoisted out of the
//                   permutation loop because X is fixed there) and the
//                   whole-column standardised copy the bootstrap contraction uses
//   colnorm, xproj  compute.normalize + `X @ U` (pyls/compute.py:97-126,
//                   pyls/base.py:364, pyls/types/behavioral.py:78)
//   normalize_flip  singular values / unit columns / sklearn svd_flip sign rule
//                   (pyls/compute.py:43-50)
#include "common.cuh"

namespace plsb {
namespace {

__global__ void pad_copy_kernel(const double *__restrict__ X, int S, int B,
                                double *__restrict__ out, int S_pad, int ldx) {
  const size_t total = (size_t)S_pad * ldx;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(e / ldx), b = (int)(e % ldx);
    out[e] = (s < S && b < B) ? X[(size_t)s * B + b] : 0.0;
  }
}

// out (cols, ld_out) = in (rows, cols)^T through a 32 x 32 tile; columns >= rows of out zeroed
__global__ void transpose_pad_kernel(const double *__restrict__ in, int rows, int cols,
                                     double *__restrict__ out, long long ld_out) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.y * 32;
  const long long r0 = (long long)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    if (c < cols && r < ld_out) out[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

__global__ void unpad_copy_kernel(const double *__restrict__ in, long long ld_in, int rows,
                                  int cols, double *__restrict__ out) {
  const size_t total = (size_t)rows * cols;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / cols, c = e % cols;
    out[e] = in[r * ld_in + c];
  }
}

// one thread per column; rows are walked cell by cell (coalesced across threads)
__global__ void prep_cells_kernel(const double *__restrict__ Xraw, double *__restrict__ Xcell,
                                  double *__restrict__ Xglob, int S, int B, int ldx, int J,
                                  const int *__restrict__ cell_start, int corr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (Xcell) {
    for (int c = 0; c < J; ++c) {
      const int r0 = cell_start[c], r1 = cell_start[c + 1], n = r1 - r0;
      double m = 0.0;
      for (int s = r0; s < r1; ++s) m += Xraw[(size_t)s * ldx + b];
      m /= n;
      double v = 0.0;
      for (int s = r0; s < r1; ++s) {
        const double d = Xraw[(size_t)s * ldx + b] - m;
        v += d * d;
      }
      const double sd = sqrt(v / (n - 1));
      for (int s = r0; s < r1; ++s) {
        const double d = Xraw[(size_t)s * ldx + b] - m;
        Xcell[(size_t)s * ldx + b] = corr ? d / sd : d;
      }
    }
  }
  if (Xglob) {
    double m = 0.0;
    for (int s = 0; s < S; ++s) m += Xraw[(size_t)s * ldx + b];
    m /= S;
    double v = 0.0;
    for (int s = 0; s < S; ++s) {
      const double d = Xraw[(size_t)s * ldx + b] - m;
      v += d * d;
    }
    const double sd = sqrt(v / (S - 1));
    for (int s = 0; s < S; ++s) {
      const double d = Xraw[(size_t)s * ldx + b] - m;
      Xglob[(size_t)s * ldx + b] = corr ? d / sd : d;
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double *red) {
  // blockDim.x multiple of 32, <= 1024
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

__global__ void colnorm_kernel(const double *__restrict__ U, int B, int L,
                               double *__restrict__ norms) {
  __shared__ double red[32];
  const int l = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double u = U[(size_t)b * L + l];
    v += u * u;
  }
  v = block_sum(v, red);
  if (threadIdx.x == 0) norms[l] = sqrt(v);
}

// out[s][l] = (1/norm_l) * sum_b X[s][b] U[b][l]; block = (row s, 8 columns of U)
__global__ void xproj_kernel(const double *__restrict__ X, int ldx, int B,
                             const double *__restrict__ U, int L,
                             const double *__restrict__ norms, double *__restrict__ out) {
  __shared__ double red[32];
  const int s = blockIdx.x, l0 = blockIdx.y * 8;
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double x = X[(size_t)s * ldx + b];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (l0 + i < L) acc[i] += x * U[(size_t)b * L + l0 + i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double t = block_sum(acc[i], red);
    if (threadIdx.x == 0 && l0 + i < L) {
      double sc = 1.0;
      if (norms) sc = norms[l0 + i] == 0.0 ? 0.0 : 1.0 / norms[l0 + i];
      out[(size_t)s * L + l0 + i] = t * sc;
    }
  }
}

// one block per column l of Uraw = R^T V: d_l = |col|, U = col / d_l with the
// sign that makes the largest-magnitude entry positive; V column flipped alike
__global__ void normalize_flip_kernel(const double *__restrict__ Uraw, int B, int L,
                                      const double *__restrict__ lam, double *__restrict__ U,
                                      double *__restrict__ V, int K, double *__restrict__ d) {
  __shared__ double red[32];
  __shared__ double s_best[32];
  __shared__ int s_idx[32];
  const int l = blockIdx.x;
  double v = 0.0, best = -1.0;
  int bidx = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double u = Uraw[(size_t)b * L + l];
    v += u * u;
    const double a = fabs(u);
    if (a > best) {  // strict: keeps the first maximum, like np.argmax
      best = a;
      bidx = b;
    }
  }
  v = block_sum(v, red);
  // arg-max across the block (ties -> smallest index)
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) {
      best = ob;
      bidx = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    s_best[warp] = best;
    s_idx[warp] = bidx;
  }
  __syncthreads();
  best = s_best[0];
  bidx = s_idx[0];
  for (int w = 1; w < nw; ++w)
    if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) {
      best = s_best[w];
      bidx = s_idx[w];
    }
  const double nrm = sqrt(v);
  // a singular value at rounding level (lam[0] = largest eigenvalue of R R^T)
  // has no direction: its column of U is set to zero instead of normalised noise
  const bool null_lv = !(nrm > 1e-10 * sqrt(lam[0]));
  const double sgn = (!null_lv && Uraw[(size_t)bidx * L + l] < 0.0) ? -1.0 : 1.0;
  const double sc = null_lv ? 0.0 : sgn / nrm;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    U[(size_t)b * L + l] = Uraw[(size_t)b * L + l] * sc;
  for (int k = threadIdx.x; k < K; k += blockDim.x) V[(size_t)k * L + l] *= sgn;
  if (threadIdx.x == 0) d[l] = nrm;
}

// G[i][j] = sum_b W[b][i] W[b][j] for a tall (B, K) matrix; one block per entry
__global__ void colgram_kernel(const double *__restrict__ W, int B, int K,
                               double *__restrict__ G) {
  __shared__ double red[32];
  const int i = blockIdx.x, j = blockIdx.y;
  double v = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    v += W[(size_t)b * K + i] * W[(size_t)b * K + j];
  v = block_sum(v, red);
  if (threadIdx.x == 0) G[(size_t)i * K + j] = v;
}

__global__ void matmul_small_kernel(const double *__restrict__ A, const double *__restrict__ Bm,
                                    int n, double *__restrict__ C) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int i = e / n, j = e - i * n;
    double v = 0.0;
    for (int k = 0; k < n; ++k) v += A[i * n + k] * Bm[k * n + j];
    C[e] = v;
  }
}

}  // namespace

int launch_colgram(plsb_ctx *h, const double *W, int B, int K, double *G, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  colgram_kernel<<<dim3(K, K), 256, 0, st>>>(W, B, K, G);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_matmul_small(plsb_ctx *h, const double *A, const double *Bm, int n, double *C,
                        cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  matmul_small_kernel<<<cdiv(n * n, 256), 256, 0, st>>>(A, Bm, n, C);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_pad_copy(plsb_ctx *h, const double *X, int S, int B, double *out, int S_pad, int ldx,
                    cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  const size_t total = (size_t)S_pad * ldx;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->sm_count * 16);
  pad_copy_kernel<<<blocks, 256, 0, st>>>(X, S, B, out, S_pad, ldx);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_transpose_pad(plsb_ctx *h, const double *in, int rows, int cols, double *out,
                         long long ld_out, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  if (cols <= 0 || ld_out <= 0) return PLSB_OK;
  dim3 grid((unsigned)((ld_out + 31) / 32), (unsigned)cdiv(cols, 32)), block(32, 8);
  transpose_pad_kernel<<<grid, block, 0, st>>>(in, rows, cols, out, ld_out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_unpad_copy(plsb_ctx *h, const double *in, long long ld_in, int rows, int cols,
                      double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  const size_t total = (size_t)rows * cols;
  if (total == 0) return PLSB_OK;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->sm_count * 16);
  unpad_copy_kernel<<<blocks, 256, 0, st>>>(in, ld_in, rows, cols, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_prep_cells(plsb_ctx *h, const double *Xraw, double *Xcell, double *Xglob,
                      cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  const Layout &l = h->lay;
  prep_cells_kernel<<<cdiv(l.B, 128), 128, 0, st>>>(Xraw, Xcell, Xglob, l.S, l.B, l.ldx, l.J,
                                                    h->d_cell_start, l.corr() ? 1 : 0);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_colnorm(plsb_ctx *h, const double *U, int B, int L, double *norms, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  colnorm_kernel<<<L, 256, 0, st>>>(U, B, L, norms);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_xproj(plsb_ctx *h, const double *Xmat, int ldx_, int S, int B, const double *U, int L,
                 const double *norms, double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  dim3 grid(S, cdiv(L, 8));
  xproj_kernel<<<grid, 256, 0, st>>>(Xmat, ldx_, B, U, L, norms, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_normalize_flip(plsb_ctx *h, const double *Uraw, int B, int L, const double *lam,
                          double *U, double *V, int K, double *d, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  normalize_flip_kernel<<<L, 256, 0, st>>>(Uraw, B, L, lam, U, V, K, d);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
