// Generic (any K, L) versions of the two streaming passes over the stored
// cross-covariances, for analyses with more latent variables than the
// fragment-table kernels of gram_proj.cu / accum_u.cu are instantiated for
// (K > MAX_K = 80; the reference's own integration shapes have K = 100 ... 400,
// pyls/tests/types/test_svd.py:6-19):
//
//   gram_proj_generic   G[r] = R[r] R[r]^T  (K x K),  H[r] = R[r] U_orig  (K x L)
//                       -- both are "A B^T" products contracted over the B features,
//                       which is the contiguous axis of R and of U_orig^T
//   accum_u_generic     U[r] = R[r]^T M[r]  (B x L);  u_sum += sum_r U[r],
//                       u_square += sum_r U[r]^2   (pyls/base.py:510-511)
//
// Plain tiled DMMA kernels (mma.sync.m8n8k4.f64): 64 x 64 CTA tiles, 4 warps of
// 32 x 32, 16-deep k chunks staged in shared memory with conflict-free pitches.
// They are correct for every shape and reasonably fast; the tuned kernels stay in
// charge of K <= 80 (every benchmark configuration).
#include "common.cuh"

namespace plsb {
namespace {

constexpr int LT = 64;          // tile edge
constexpr int LK = 16;          // k chunk
constexpr int LTHREADS = 128;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// C[r] (K x N) = R[r] (K x B) * P^T with the rows of P taken from R[r] itself (n < K: the
// Gram matrix) and from the projection block UoT (n >= K), contraction over the columns.
// grid: (tiles over K) x (tiles over N = K + L) x count
__global__ void __launch_bounds__(LTHREADS)
gram_proj_generic_kernel(const double *__restrict__ R, long long ldr, int K, int B,
                         const double *__restrict__ UoT, long long uot_stride, int uot_div, int L,
                         double *__restrict__ G, double *__restrict__ H) {
  constexpr int LDS = LK + 4;   // 20: fragment loads (8 rows x 4 k) hit 32 distinct banks
  __shared__ __align__(16) double As[LT * LDS];
  __shared__ __align__(16) double Bs[LT * LDS];
  const int r = blockIdx.z, m0 = blockIdx.x * LT, n0 = blockIdx.y * LT;
  const int N = K + L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const double *Rr = R + (size_t)r * K * ldr;
  const double *Pr = UoT ? UoT + (size_t)(r / uot_div) * uot_stride : nullptr;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int b0 = 0; b0 < B; b0 += LK) {
    // 64 rows x 16 columns of each operand, one 16-byte piece per thread and pass
    for (int e = tid; e < LT * (LK / 2); e += LTHREADS) {
      const int row = e / (LK / 2), seg = e - row * (LK / 2);
      const int col = b0 + seg * 2;
      double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
      const int m = m0 + row, n = n0 + row;
      if (m < K) {
        const double *src = Rr + (size_t)m * ldr + col;
        if (col + 1 < B) a = *reinterpret_cast<const double2 *>(src);
        else if (col < B) a.x = src[0];
      }
      if (n < N) {
        const double *src = n < K ? Rr + (size_t)n * ldr + col
                                  : Pr + (size_t)(n - K) * ldr + col;
        if (col + 1 < B) b = *reinterpret_cast<const double2 *>(src);
        else if (col < B) b.x = src[0];
      }
      *reinterpret_cast<double2 *>(As + row * LDS + seg * 2) = a;
      *reinterpret_cast<double2 *>(Bs + row * LDS + seg * 2) = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[(wm * 32 + i * 8 + g) * LDS + kk + q];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn * 32 + j * 8 + g) * LDS + kk + q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int m = m0 + wm * 32 + i * 8 + g, n = n0 + wn * 32 + j * 8 + 2 * q + c;
        if (m >= K || n >= N) continue;
        if (n < K)
          G[(size_t)r * K * K + (size_t)m * K + n] = acc[i][j][c];
        else if (H)
          H[(size_t)r * K * L + (size_t)m * L + (n - K)] = acc[i][j][c];
      }
}

// One CTA per 64 (features) x 64 (latent variables) tile of u_sum / u_square; it runs
// through ALL resamples, accumulating sum U and sum U^2 in registers (fixed order:
// deterministic), and adds them to the outputs once.
__global__ void __launch_bounds__(LTHREADS)
accum_u_generic_kernel(const double *__restrict__ R, long long ldr, int count, int K, int B,
                       const double *__restrict__ M, int ldm, int L, double *__restrict__ usum,
                       double *__restrict__ usq) {
  constexpr int LDT = LT + 4;   // 68: (4 k) x (8 rows) fragment loads are conflict free
  __shared__ __align__(16) double Rs[LK * LDT];   // [k][feature]
  __shared__ __align__(16) double Ms[LK * LDT];   // [k][latent variable]
  const int b0 = blockIdx.x * LT, l0 = blockIdx.y * LT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  double s1[4][4][2], s2[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s1[i][j][0] = s1[i][j][1] = s2[i][j][0] = s2[i][j][1] = 0.0;

  for (int r = 0; r < count; ++r) {
    const double *Rr = R + (size_t)r * K * ldr;
    const double *Mr = M + (size_t)r * K * ldm;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < K; k0 += LK) {
      for (int e = tid; e < LK * LT; e += LTHREADS) {
        const int k = e / LT, c = e - k * LT;
        const bool in = k0 + k < K;
        Rs[k * LDT + c] = (in && b0 + c < B) ? Rr[(size_t)(k0 + k) * ldr + b0 + c] : 0.0;
        Ms[k * LDT + c] = (in && l0 + c < L) ? Mr[(size_t)(k0 + k) * ldm + l0 + c] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < LK; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = Rs[(kk + q) * LDT + wm * 32 + i * 8 + g];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Ms[(kk + q) * LDT + wn * 32 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const double u = acc[i][j][c];
          s1[i][j][c] += u;
          s2[i][j][c] += u * u;
        }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int b = b0 + wm * 32 + i * 8 + g, l = l0 + wn * 32 + j * 8 + 2 * q + c;
        if (b < B && l < L) {
          usum[(size_t)b * L + l] += s1[i][j][c];
          usq[(size_t)b * L + l] += s2[i][j][c];
        }
      }
}

}  // namespace

int launch_gram_proj_generic(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                             const double *UoT, int L, double *G, double *H, cudaStream_t st,
                             long long uot_stride, int uot_div) {
  const bool proj = UoT && H;
  const int N = K + (proj ? L : 0);
  PLSB_CHECK(ldr % 2 == 0, PLSB_ERR_ARG, "gram_proj: odd row pitch");
  for (int off = 0; off < count; off += 65535) {   // grid.z limit
    const int n = std::min(65535, count - off);
    dim3 grid(cdiv(K, LT), cdiv(N, LT), n);
    gram_proj_generic_kernel<<<grid, LTHREADS, 0, st>>>(
        R + (size_t)off * K * ldr, ldr, K, B,
        proj ? UoT + (size_t)(off / std::max(uot_div, 1)) * uot_stride : nullptr, uot_stride,
        std::max(uot_div, 1), proj ? L : 0, G + (size_t)off * K * K,
        proj ? H + (size_t)off * K * L : nullptr);
    PLSB_LAUNCHED(h);
  }
  return PLSB_OK;
}

int launch_accum_u_generic(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                           const double *M, int ldm, int L, double *usum, double *usq,
                           cudaStream_t st) {
  dim3 grid(cdiv(B, LT), cdiv(L, LT));
  accum_u_generic_kernel<<<grid, LTHREADS, 0, st>>>(R, ldr, count, K, B, M, ldm, L, usum, usq);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
