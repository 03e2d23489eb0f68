// Small dense decompositions, one CTA per resample.
//
// The reference runs sklearn's randomized_svd on the (K, B) cross-covariance
// of every resample (pyls/compute.py:36-49) and a second SVD inside
// compute.procrustes (pyls/compute.py:260-262).  Both collapse onto K x K
// problems once G = R R^T and H = R U_orig are known:
//
//   eigen_kernel      G = V diag(lam) V^T                    (Jacobi)
//   rotation_kernel   d = sqrt(lam);  temp = H^T V d^-1      (= U_orig^T U_boot)
//                     Q = polar(temp)^T   (= P^T N^T of compute.py:262; Newton-Schulz)
//                     M = V Q             so that U_boot d Q = R^T M
//
// eigen_kernel: cyclic two-sided Jacobi with a round-robin (tournament)
// ordering.  The n/2 plane rotations of one round act on disjoint index pairs,
// so the update A <- J^T A J decomposes into independent 2 x 2 blocks, one per
// (pair a, pair b): a thread reads the four entries of its block, applies pair
// b's rotation from the right and pair a's from the left, and writes the block
// and its mirror image (jacobi.cuh).  Two K x K tiles of shared memory.
//
// rotation_kernel: the orthogonal polar factor comes from the Newton-Schulz
// iteration X <- X (3 I - X^T X) / 2 -- matrix products only, done as
// mma.sync.m8n8k4.f64 (DMMA) fragments out of shared memory.
//
// Numerically null directions (mean-centred PLS always has one: the cell
// means minus their mean have rank J-1) are removed from the rotation: their
// d^-1 is 0 and the rows of temp that belong to null ORIGINAL latent variables
// are zeroed; exact-zero rows / columns stay zero under Newton-Schulz, so Q is
// the Procrustes rotation of the non-null subspace and the null columns of
// R^T M are 0.  (The reference rotates with whatever unit vectors its
// randomized SVD returns for the null directions, which perturbs the other
// latent variables; see DESIGN.md.)
#include "common.cuh"
#include "jacobi.cuh"

namespace plsb {
namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;

// eigen-decomposition of the symmetric K x K matrices G[r]:
//   V_out (K,K): eigenvectors in columns, sorted by descending eigenvalue
//   lam_out (K): eigenvalues sorted descending (sqrt_lam: their square roots)
__global__ void __launch_bounds__(SM_THREADS)
eigen_kernel(const double *__restrict__ G, int ldg, long long g_stride, int K, int sqrt_lam,
             double *__restrict__ V_out, double *__restrict__ lam_out) {
  extern __shared__ __align__(16) double sm[];
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  double *bufA = sm;                 // G -> diag(lam)
  double *bufV = bufA + ne * ld;     // V
  double *lam = bufV + ne * ld;      // ne
  JacobiScratch sc;
  sc.cst = lam + ne;                                   // 3*half
  sc.pq = reinterpret_cast<int *>(sc.cst + 3 * half);  // 2*half
  int *rank = sc.pq + 2 * half;                        // ne
  sc.blk = reinterpret_cast<short2 *>(rank + ne);      // half*(half+1)/2
  const int r = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const double *Gr = G + (size_t)r * g_stride;

  // block table: bi -> (a, b), a <= b
  for (int a = tid; a < half; a += nthr) {
    int bi = a * half - a * (a - 1) / 2;   // blocks of the rows before a
    for (int b = a; b < half; ++b) sc.blk[bi++] = make_short2((short)a, (short)b);
  }
  for (int e = tid; e < ne * ne; e += nthr) {
    const int i = e / ne, j = e - i * ne;
    double g = 0.0;
    // symmetrise: both triangles must agree exactly for the two-sided updates
    if (i < K && j < K) g = 0.5 * (Gr[(size_t)i * ldg + j] + Gr[(size_t)j * ldg + i]);
    bufA[i * ld + j] = g;
    bufV[i * ld + j] = (i == j && i < K) ? 1.0 : 0.0;
  }
  jacobi_sym(bufA, bufV, K, ld, sc);   // diag(bufA) = lam, bufV = V
  for (int j = tid; j < K; j += nthr) lam[j] = fmax(bufA[j * ld + j], 0.0);
  __syncthreads();
  // descending rank of every eigenvalue (ties broken by index)
  for (int j = tid; j < K; j += nthr) {
    int rk = 0;
    for (int i = 0; i < K; ++i) rk += (lam[i] > lam[j]) || (lam[i] == lam[j] && i < j);
    rank[j] = rk;
  }
  __syncthreads();
  if (lam_out)
    for (int j = tid; j < K; j += nthr)
      lam_out[(size_t)r * K + rank[j]] = sqrt_lam ? sqrt(lam[j]) : lam[j];
  if (V_out)
    for (int e = tid; e < K * K; e += nthr) {
      const int i = e / K, j = e - i * K;
      V_out[(size_t)r * K * K + (size_t)i * K + rank[j]] = bufV[i * ld + j];
    }
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// C[m][n] = sum_k Aop(m,k) Bop(k,n) for an LP x LP x LP product out of shared
// memory, Aop(m,k) = A[m*sam + k*sak], Bop(k,n) = B[k*sbk + n*sbn].  The
// (LP/8)^2 output fragments are dealt round-robin to the warps; `epi(row, col,
// value)` is called for every output element.  No barrier inside.
template <typename Epi>
__device__ __forceinline__ void small_mm(const double *A, int sam, int sak, const double *B,
                                         int sbk, int sbn, int LP, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int nf = LP / 8;
  for (int f = warp; f < nf * nf; f += SM_WARPS) {
    const int i0 = (f / nf) * 8, j0 = (f % nf) * 8;
    double c0 = 0.0, c1 = 0.0;
    const double *pa = A + (i0 + g) * sam + q * sak;
    const double *pb = B + q * sbk + (j0 + g) * sbn;
    for (int k0 = 0; k0 < LP; k0 += 4) dmma_8x8x4(c0, c1, pa[k0 * sak], pb[k0 * sbk]);
    epi(i0 + g, j0 + 2 * q, c0);
    epi(i0 + g, j0 + 2 * q + 1, c1);
  }
}

// M[r] (K,L) from V[r] (K,K), lam[r] (K) (eigen_kernel's output) and H[r] (K,L)
__global__ void __launch_bounds__(SM_THREADS)
rotation_kernel(const double *__restrict__ H, const double *__restrict__ V_in,
                const double *__restrict__ lam_in, int K, int L,
                const double *__restrict__ dorig, double *__restrict__ M_out, int ldm) {
  extern __shared__ __align__(16) double sm[];
  const int LP = (K + 7) & ~7, ld = LP + 4;
  double *Vs = sm;                // V          [LP][ld]
  double *Hs = Vs + LP * ld;      // H, later the second iterate
  double *Xs = Hs + LP * ld;      // temp = X_0
  double *Bs = Xs + LP * ld;      // (3 I - X^T X) / 2
  double *dinv = Bs + LP * ld;    // LP
  double *omask = dinv + LP;      // LP: 1 for non-null original latent variables
  const int r = blockIdx.x, tid = threadIdx.x;

  for (int e = tid; e < LP * LP; e += SM_THREADS) {
    const int i = e / LP, j = e - i * LP;
    const bool in = i < K && j < K;
    Vs[i * ld + j] = in ? V_in[(size_t)r * K * K + (size_t)i * K + j] : 0.0;
    Hs[i * ld + j] = in ? H[(size_t)r * K * L + (size_t)i * L + j] : 0.0;
  }
  if (tid < LP) {
    double lmax = 0.0, domax = 0.0;
    for (int i = 0; i < K; ++i) lmax = fmax(lmax, lam_in[(size_t)r * K + i]);
    if (dorig)
      for (int i = 0; i < L; ++i) domax = fmax(domax, dorig[i]);
    double di = 0.0, om = 0.0;
    if (tid < K) {
      const double l = lam_in[(size_t)r * K + tid];
      // d^-1 with a guard for numerically null directions
      di = (l > 1e-14 * lmax && l > 0.0) ? rsqrt(l) : 0.0;
      // a numerically null ORIGINAL latent variable has no direction to rotate onto
      om = (!dorig || dorig[tid] > 1e-10 * domax) ? 1.0 : 0.0;
    }
    dinv[tid] = di;
    omask[tid] = om;
  }
  __syncthreads();
  // temp[i][j] = omask_i * sum_k H[k][i] V[k][j] * dinv_j
  small_mm(Hs, 1, ld, Vs, ld, 1, LP,
           [&](int i, int j, double v) { Xs[i * ld + j] = v * omask[i] * dinv[j]; });
  __syncthreads();

  // Newton-Schulz:  X <- X (3 I - X^T X) / 2.  Singular values of temp are
  // cosines of principal angles (<= 1 < sqrt(3)): it converges, quadratically
  // at the end.
  double *X = Xs, *Y = Hs;
  for (int it = 0; it < 100; ++it) {
    small_mm(X, 1, ld, X, ld, 1, LP, [&](int i, int j, double v) {
      Bs[i * ld + j] = (i == j ? 1.5 : 0.0) - 0.5 * v;
    });
    __syncthreads();
    if (it == 0) {
      // |X^T X|_inf bounds sigma_max^2; the iteration needs sigma_max < sqrt(3).
      // (The polar factor does not depend on a positive scale of X.)
      double rs = 0.0;
      if (tid < LP)
        for (int j = 0; j < LP; ++j) rs += fabs((tid == j ? 3.0 : 0.0) - 2.0 * Bs[tid * ld + j]);
      // max over the CTA through an integer OR of the comparison is not enough:
      // reduce the maximum with shuffles + shared memory
      for (int o = 16; o > 0; o >>= 1) rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
      __shared__ double s_red[SM_WARPS];
      if ((tid & 31) == 0) s_red[tid >> 5] = rs;
      __syncthreads();
      double s = 0.0;
      for (int w = 0; w < SM_WARPS; ++w) s = fmax(s, s_red[w]);
      if (s > 2.8) {
        const double f2 = 2.8 / s, f = sqrt(f2);
        for (int e = tid; e < LP * LP; e += SM_THREADS) {
          const int i = e / LP, j = e - i * LP;
          X[i * ld + j] *= f;
          const double xtx = (i == j ? 3.0 : 0.0) - 2.0 * Bs[i * ld + j];
          Bs[i * ld + j] = (i == j ? 1.5 : 0.0) - 0.5 * f2 * xtx;
        }
      }
      __syncthreads();
    }
    int changed = 0;
    small_mm(X, ld, 1, Bs, ld, 1, LP, [&](int i, int j, double v) {
      Y[i * ld + j] = v;
      changed |= fabs(v - X[i * ld + j]) > 1e-14;
    });
    const int more = __syncthreads_or(changed);
    double *t = X;
    X = Y;
    Y = t;
    if (!more) break;
  }
  // Q = polar(temp)^T;  M[a][j] = sum_i V[a][i] X[j][i]
  small_mm(Vs, ld, 1, X, 1, ld, LP, [&](int a, int j, double v) {
    if (a < K && j < L) M_out[((size_t)r * K + a) * ldm + j] = v;
  });
  // padding columns of a padded output pitch
  for (int e = tid; e < K * (ldm - L); e += SM_THREADS) {
    const int a = e / (ldm - L), j = L + e - a * (ldm - L);
    M_out[((size_t)r * K + a) * ldm + j] = 0.0;
  }
}

int launch_eigen(plsb_ctx *h, const double *G, int count, int K, int sqrt_lam, double *V,
                 double *lam, cudaStream_t st, int ldg = 0, long long g_stride = 0) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= MAX_K, PLSB_ERR_ARG, "small decomposition: K=%d outside [1,%d]", K,
             MAX_K);
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  const int nb = half * (half + 1) / 2;
  const size_t smem = sizeof(double) * (2 * (size_t)ne * ld + ne + 3 * half) +
                      sizeof(int) * (2 * half + ne) + sizeof(short2) * nb + 16;
  PLSB_CUDA(cudaFuncSetAttribute(eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  // one thread per 2 x 2 block of a Jacobi round (measured: fewer, busier threads lose)
  const int threads = std::min(SM_THREADS, std::max(32, round_up(
      tune_int("PLSB_EIGEN_THREADS", nb), 32)));
  if (ldg <= 0) ldg = K;
  if (g_stride <= 0) g_stride = (long long)K * K;
  eigen_kernel<<<count, threads, smem, st>>>(G, ldg, g_stride, K, sqrt_lam, V, lam);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_rotation(plsb_ctx *h, const double *H, const double *V, const double *lam, int count,
                    int K, int L, const double *dorig, double *M, int ldm, cudaStream_t st) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(L == K, PLSB_ERR_ARG, "small decomposition: L=%d must equal K=%d", L, K);
  const int LP = round_up(K, 8), ld = LP + 4;
  const size_t smem = sizeof(double) * (4 * (size_t)LP * ld + 2 * LP);
  PLSB_CUDA(cudaFuncSetAttribute(rotation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  rotation_kernel<<<count, SM_THREADS, smem, st>>>(H, V, lam, K, L, dorig, M, ldm);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace

int launch_small_decomp(plsb_ctx *h, const double *G, const double *H, int count, int K, int L,
                        const double *dorig, double *M, int ldm, double *lam, cudaStream_t st) {
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(ldm >= L, PLSB_ERR_ARG, "small decomposition: output pitch %d < L=%d", ldm, L);
  const size_t kk = (size_t)K * K;
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)count * (kk + K)));
  double *V = h->misc.as<double>(), *lam_s = V + (size_t)count * kk;
  PLSB_TRY(launch_eigen(h, G, count, K, 0, V, lam_s, st));
  PLSB_TRY(launch_rotation(h, H, V, lam_s, count, K, L, dorig, M, ldm, st));
  if (lam)
    PLSB_CUDA(cudaMemcpyAsync(lam, lam_s, sizeof(double) * (size_t)count * K,
                              cudaMemcpyDeviceToDevice, st));
  return PLSB_OK;
}

int launch_sym_eig(plsb_ctx *h, const double *G, int count, int K, double *V, double *lam,
                   int sqrt_lam, cudaStream_t st, int ldg, long long g_stride) {
  return launch_eigen(h, G, count, K, sqrt_lam, V, lam, st, ldg, g_stride);
}

}  // namespace plsb
