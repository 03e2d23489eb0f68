// Helpers of the "tall" orientation: analyses with more latent rows than features
// (K = cells x behaviours > B).  compute.svd then decomposes the cross-covariance the other
// way round (pyls/compute.py:46-50) and every small problem lives on the FEATURE side:
// with R (K x B) of a resample and R' = R^T (B x K),
//     G' = R' R'^T = R^T R = U d^2 U^T            (B x B; U = feature-side singular vectors)
//     permutation, rotated:  H' = R' V_orig,  temp = H'^T U d^-1 = V_orig^T V_perm,
//         Q = polar(temp)^T,  M = U Q,  ssd_j = |(V_perm d Q)[:, j]| = sqrt(m_j^T G' m_j)
//     bootstrap:  temp = U_orig^T U,  Q = polar(temp)^T,  U_boot d Q = U d Q  (B x B): the
//         accumulators u_sum / u_square are sums of B x B matrices, no B-sized pass at all
// The big GEMM and the operand builders are the ones of the wide orientation; R is
// transposed per resample (it is small: B < K) and the streaming / small-matrix kernels run
// with the roles of K and B exchanged.
#include "common.cuh"

namespace plsb {
namespace {

// out[r] (cols x ld_out) = in[r] (rows x cols, pitch ld_in)^T, padding columns zeroed
__global__ void transpose_batch_kernel(const double *__restrict__ in, int rows, int cols,
                                       long long ld_in, long long in_stride,
                                       double *__restrict__ out, long long ld_out,
                                       long long out_stride) {
  __shared__ double tile[32][33];
  const int r = blockIdx.z;
  const double *src = in + (size_t)r * in_stride;
  double *dst = out + (size_t)r * out_stride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;   // c0: input column / output row
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int row = r0 + i, col = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (row < rows && col < cols) ? src[(size_t)row * ld_in + col] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int orow = c0 + i, ocol = r0 + threadIdx.x;
    if (orow < cols && ocol < ld_out) dst[(size_t)orow * ld_out + ocol] = tile[threadIdx.x][i];
  }
}

// out[r][j] = sqrt(m_j^T G[r] m_j), G (count, n, n), M (count, n, L) dense; one CTA per resample
__global__ void quadform_sqrt_kernel(const double *__restrict__ G, const double *__restrict__ M,
                                     int n, int L, double *__restrict__ out) {
  extern __shared__ double gm[];   // n x L: G M
  const int r = blockIdx.x;
  const double *Gr = G + (size_t)r * n * n, *Mr = M + (size_t)r * n * L;
  for (int e = threadIdx.x; e < n * L; e += blockDim.x) {
    const int a = e / L, j = e - a * L;
    double v = 0.0;
    for (int b = 0; b < n; ++b) v += Gr[(size_t)a * n + b] * Mr[(size_t)b * L + j];
    gm[e] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    double v = 0.0;
    for (int a = 0; a < n; ++a) v += Mr[(size_t)a * L + j] * gm[a * L + j];
    out[(size_t)r * L + j] = sqrt(fmax(v, 0.0));
  }
}

// usum[e] += sum_r M[r][e], usq[e] += sum_r M[r][e]^2 in resample order (deterministic)
__global__ void accum_small_kernel(const double *__restrict__ M, int count, long long n_elem,
                                   double *__restrict__ usum, double *__restrict__ usq) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double s1 = 0.0, s2 = 0.0;
  for (int r = 0; r < count; ++r) {
    const double v = M[(size_t)r * n_elem + e];
    s1 += v;
    s2 += v * v;
  }
  usum[e] += s1;
  usq[e] += s2;
}

}  // namespace

int launch_transpose_batch(plsb_ctx *h, const double *in, int rows, int cols, long long ld_in,
                           long long in_stride, double *out, long long ld_out,
                           long long out_stride, int count, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  if (count <= 0) return PLSB_OK;
  // columns of `out` beyond `rows` must be zero as well: cover the whole pitch
  dim3 block(32, 8);
  for (int off = 0; off < count; off += 65535) {
    const int n = std::min(65535, count - off);
    dim3 grid(cdiv(cols, 32), (unsigned)((ld_out + 31) / 32), n);
    transpose_batch_kernel<<<grid, block, 0, st>>>(in + (size_t)off * in_stride, rows, cols, ld_in,
                                                   in_stride, out + (size_t)off * out_stride,
                                                   ld_out, out_stride);
    PLSB_LAUNCHED(h);
  }
  return PLSB_OK;
}

int launch_quadform_sqrt(plsb_ctx *h, const double *G, const double *M, int count, int n, int L,
                         double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  const size_t smem = sizeof(double) * (size_t)n * L;
  PLSB_CHECK(smem <= 200 * 1024, PLSB_ERR_ARG,
             "quadratic forms of %d x %d matrices exceed shared memory", n, L);
  PLSB_CUDA(cudaFuncSetAttribute(quadform_sqrt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  quadform_sqrt_kernel<<<count, 256, smem, st>>>(G, M, n, L, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_accum_small(plsb_ctx *h, const double *M, int count, long long n_elem, double *usum,
                       double *usq, cudaStream_t st) {
  KernelTimer kt(h, KC_ACCUM, st);
  if (count <= 0 || n_elem <= 0) return PLSB_OK;
  accum_small_kernel<<<(unsigned)((n_elem + 255) / 256), 256, 0, st>>>(M, count, n_elem, usum,
                                                                       usq);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
