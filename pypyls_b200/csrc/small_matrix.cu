// Small dense decompositions, one CTA per resample.
//
// The reference runs sklearn's randomized_svd on the (K, B) cross-covariance
// of every resample (pyls/compute.py:36-49) and a second SVD inside
// compute.procrustes (pyls/compute.py:260-262).  Both collapse onto K x K
// symmetric eigenproblems once G = R R^T and H = R U_orig are known:
//
//   G = V diag(lam) V^T                      Jacobi on G
//   d = sqrt(lam);  temp = H^T V d^-1        (= U_orig^T U_boot)
//   temp^T temp = W diag(mu) W^T             Jacobi on temp^T temp
//   N s = temp W;  Q = W N^T                 (= P^T N^T of compute.py:262)
//   M = V Q                                  so that U_boot d Q = R^T M
//
// The eigen-solver is the cyclic two-sided Jacobi method with a round-robin
// (tournament) ordering: the n/2 plane rotations of one round act on disjoint
// index pairs, so their parameters come from three matrix entries each and
// the column / row updates of all pairs run concurrently across the CTA with
// three barriers per round and no reductions.
//
// Numerically null directions (mean-centred PLS always has one: the cell
// means minus their mean have rank J-1) are removed from the rotation: their
// d^-1 is 0 and the rows of temp that belong to null ORIGINAL latent variables
// are zeroed, so Q is the Procrustes rotation of the non-null subspace and the
// null columns of R^T M are 0.  (The reference rotates with whatever unit
// vectors its randomized SVD returns for the null directions, which perturbs
// the other latent variables; see DESIGN.md.)
#include "common.cuh"

namespace plsb {
namespace {

constexpr int SM_THREADS = 256;
constexpr int MAX_SWEEPS = 30;
constexpr double JACOBI_TOL = 1e-14;
constexpr double JACOBI_EPS = 1e-15;

// Diagonalises the symmetric n x n matrix A (row-major, leading dimension ld,
// both triangles stored) in place: A <- J^T A J, V <- V J over all rotations.
// Called by every thread of the CTA.
//
// One round = the n/2 disjoint pairs of the tournament schedule.  Every warp
// owns up to NPW pairs of the round: lane j derives the rotation of the warp's
// j-th pair from (a_pp, a_qq, a_pq) -- entries no other pair's column update
// touches -- the parameters are broadcast with shuffles, then the warp rotates
// the columns of its pairs (lanes stride over the rows) and, after one barrier,
// the rows (lanes stride over the columns).  Two barriers per round.
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int NPW = (MAX_K / 2 + SM_WARPS - 1) / SM_WARPS;   // pairs per warp per round

__device__ void jacobi_sym(double *A, double *V, int n, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ne = n + (n & 1);       // even number of players (last may be a dummy)
  const int half = ne / 2;
  __syncthreads();
  if (n < 2) return;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    int any = 0;
    for (int round = 0; round < ne - 1; ++round) {
      // rotation of pair (warp + lane * SM_WARPS), lanes < NPW
      int my_p = -1, my_q = 0;
      double my_c = 1.0, my_s = 0.0;
      const int my_pi = warp + lane * SM_WARPS;
      if (lane < NPW && my_pi < half) {
        int a, b;
        if (my_pi == 0) {
          a = ne - 1;
          b = round;
        } else {
          a = (round + my_pi) % (ne - 1);
          b = (round - my_pi + (ne - 1)) % (ne - 1);
        }
        const int p = min(a, b), q = max(a, b);
        if (q < n) {
          const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
          // rotate while the off-diagonal entry is above both the relative
          // (graded-matrix) threshold and the rounding level of the larger
          // diagonal entry -- below that a rotation only shuffles noise
          const double thr = fmax(JACOBI_TOL * sqrt(fabs(app * aqq)),
                                  JACOBI_EPS * fmax(fabs(app), fabs(aqq)));
          if (fabs(apq) > thr) {
            const double zeta = (aqq - app) / (2.0 * apq);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            my_c = rsqrt(1.0 + t * t);
            my_s = my_c * t;
            my_p = p;
            my_q = q;
          }
        }
      }
      int pp[NPW], qq[NPW];
      double cc[NPW], ss[NPW];
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        pp[j] = __shfl_sync(0xffffffffu, my_p, j);
        qq[j] = __shfl_sync(0xffffffffu, my_q, j);
        cc[j] = __shfl_sync(0xffffffffu, my_c, j);
        ss[j] = __shfl_sync(0xffffffffu, my_s, j);
      }
      // column rotations  A <- A J,  V <- V J
      int rotated = 0;
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        if (pp[j] < 0) continue;
        rotated = 1;
        const int p = pp[j], q = qq[j];
        const double c = cc[j], s = ss[j];
        for (int i = lane; i < n; i += 32) {
          const double x = A[i * ld + p], y = A[i * ld + q];
          A[i * ld + p] = c * x - s * y;
          A[i * ld + q] = s * x + c * y;
          const double vx = V[i * ld + p], vy = V[i * ld + q];
          V[i * ld + p] = c * vx - s * vy;
          V[i * ld + q] = s * vx + c * vy;
        }
      }
      any |= __syncthreads_or(rotated);
      // row rotations  A <- J^T A
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        if (pp[j] < 0) continue;
        const int p = pp[j], q = qq[j];
        const double c = cc[j], s = ss[j];
        for (int i = lane; i < n; i += 32) {
          const double x = A[p * ld + i], y = A[q * ld + i];
          A[p * ld + i] = c * x - s * y;
          A[q * ld + i] = s * x + c * y;
        }
      }
      __syncthreads();
    }
    if (!any) break;
  }
}

// mode 0: full decomposition -> M (K,L) [+ lam sorted descending if lam_out]
// mode 1: eigen only -> V_out (K,K) columns sorted by descending eigenvalue, lam_out sorted
//         (sqrt_lam != 0 writes sqrt(lam) instead: singular values of R)
__global__ void __launch_bounds__(SM_THREADS)
small_decomp_kernel(const double *__restrict__ G, const double *__restrict__ H, int K, int L,
                    int mode, int sqrt_lam, const double *__restrict__ dorig,
                    double *__restrict__ M_out, double *__restrict__ V_out,
                    double *__restrict__ lam_out) {
  extern __shared__ __align__(16) double sm[];
  const int ld = K | 1;
  double *bufA = sm;                 // G, later temp^T temp, later N s
  double *bufV = bufA + K * ld;      // V
  double *bufT = bufV + K * ld;      // temp, later Q
  double *bufW = bufT + K * ld;      // H, later W
  double *lam = bufW + K * ld;       // K
  double *aux = lam + K;             // K
  int *rank = reinterpret_cast<int *>(aux + K);  // K
  const int r = blockIdx.x, tid = threadIdx.x;
  const double *Gr = G + (size_t)r * K * K;

  for (int e = tid; e < K * K; e += SM_THREADS) {
    const int i = e / K, j = e - i * K;
    // symmetrise: both triangles must agree exactly for the two-sided updates
    bufA[i * ld + j] = 0.5 * (Gr[(size_t)i * K + j] + Gr[(size_t)j * K + i]);
    bufV[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  jacobi_sym(bufA, bufV, K, ld);   // diag(bufA) = lam, bufV = V
  for (int j = tid; j < K; j += SM_THREADS) lam[j] = fmax(bufA[j * ld + j], 0.0);
  __syncthreads();
  // descending rank of every eigenvalue (ties broken by index)
  for (int j = tid; j < K; j += SM_THREADS) {
    int rk = 0;
    for (int i = 0; i < K; ++i) rk += (lam[i] > lam[j]) || (lam[i] == lam[j] && i < j);
    rank[j] = rk;
  }
  __syncthreads();
  if (lam_out)
    for (int j = tid; j < K; j += SM_THREADS)
      lam_out[(size_t)r * K + rank[j]] = sqrt_lam ? sqrt(lam[j]) : lam[j];
  if (mode == 1) {
    if (V_out)
      for (int e = tid; e < K * K; e += SM_THREADS) {
        const int i = e / K, j = e - i * K;
        V_out[(size_t)r * K * K + (size_t)i * K + rank[j]] = bufV[i * ld + j];
      }
    return;
  }

  // d^-1 with a guard for numerically null directions
  double lmax = 0.0;
  for (int i = 0; i < K; ++i) lmax = fmax(lmax, lam[i]);
  double dorig_max = 0.0;
  if (dorig)
    for (int i = 0; i < L; ++i) dorig_max = fmax(dorig_max, dorig[i]);
  __syncthreads();
  for (int j = tid; j < K; j += SM_THREADS)
    aux[j] = (lam[j] > 1e-14 * lmax && lam[j] > 0.0) ? 1.0 / sqrt(lam[j]) : 0.0;
  // bufW <- H (K x L)
  const double *Hr = H + (size_t)r * K * L;
  for (int e = tid; e < K * L; e += SM_THREADS) {
    const int k = e / L, i = e - k * L;
    bufW[k * ld + i] = Hr[e];
  }
  __syncthreads();
  // temp[i][j] = sum_k H[k][i] V[k][j] / d_j   (L x L)
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int i = e / L, j = e - i * L;
    double v = 0.0;
    for (int k = 0; k < K; ++k) v += bufW[k * ld + i] * bufV[k * ld + j];
    // a numerically null ORIGINAL latent variable has no direction to rotate onto
    if (dorig && !(dorig[i] > 1e-10 * dorig_max)) v = 0.0;
    bufT[i * ld + j] = v * aux[j];
  }
  __syncthreads();
  // S = temp^T temp -> bufA;  W <- I
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int i = e / L, j = e - i * L;
    if (j >= i) {
      double v = 0.0;
      for (int k = 0; k < L; ++k) v += bufT[k * ld + i] * bufT[k * ld + j];
      bufA[i * ld + j] = v;
      bufA[j * ld + i] = v;
    }
    bufW[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  jacobi_sym(bufA, bufW, L, ld);   // bufW = W (right singular vectors)
  // N s = temp W -> bufA
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int i = e / L, k = e - i * L;
    double v = 0.0;
    for (int j = 0; j < L; ++j) v += bufT[i * ld + j] * bufW[j * ld + k];
    bufA[i * ld + k] = v;
  }
  __syncthreads();
  // singular values = column norms of temp W
  for (int k = tid; k < L; k += SM_THREADS) {
    double v = 0.0;
    for (int i = 0; i < L; ++i) v += bufA[i * ld + k] * bufA[i * ld + k];
    lam[k] = sqrt(v);
  }
  __syncthreads();
  double smax = 0.0;
  for (int i = 0; i < L; ++i) smax = fmax(smax, lam[i]);
  __syncthreads();
  for (int k = tid; k < L; k += SM_THREADS)
    aux[k] = (lam[k] > 1e-12 * smax && lam[k] > 0.0) ? 1.0 / lam[k] : 0.0;
  __syncthreads();
  // Q[i][j] = sum_k W[i][k] N[j][k]  -> bufT
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int i = e / L, j = e - i * L;
    double v = 0.0;
    for (int k = 0; k < L; ++k) v += bufW[i * ld + k] * bufA[j * ld + k] * aux[k];
    bufT[i * ld + j] = v;
  }
  __syncthreads();
  // M[a][j] = sum_i V[a][i] Q[i][j]
  for (int e = tid; e < K * L; e += SM_THREADS) {
    const int a = e / L, j = e - a * L;
    double v = 0.0;
    for (int i = 0; i < K; ++i) v += bufV[a * ld + i] * bufT[i * ld + j];
    M_out[(size_t)r * K * L + e] = v;
  }
}

int launch_small(plsb_ctx *h, const double *G, const double *H, int count, int K, int L, int mode,
                 int sqrt_lam, const double *dorig, double *M, double *V, double *lam,
                 cudaStream_t st) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= MAX_K, PLSB_ERR_ARG, "small decomposition: K=%d outside [1,%d]", K,
             MAX_K);
  PLSB_CHECK(mode == 1 || L == K, PLSB_ERR_ARG, "small decomposition: L=%d must equal K=%d", L, K);
  const int ld = K | 1;
  const size_t smem = sizeof(double) * (4 * (size_t)K * ld + 2 * K) + sizeof(int) * K;
  PLSB_CUDA(cudaFuncSetAttribute(small_decomp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  small_decomp_kernel<<<count, SM_THREADS, smem, st>>>(G, H, K, L, mode, sqrt_lam, dorig, M, V,
                                                       lam);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace

int launch_small_decomp(plsb_ctx *h, const double *G, const double *H, int count, int K, int L,
                        const double *dorig, double *M, double *lam, cudaStream_t st) {
  return launch_small(h, G, H, count, K, L, 0, 0, dorig, M, nullptr, lam, st);
}

int launch_sym_eig(plsb_ctx *h, const double *G, int count, int K, double *V, double *lam,
                   int sqrt_lam, cudaStream_t st) {
  return launch_small(h, G, nullptr, count, K, K, 1, sqrt_lam, nullptr, nullptr, V, lam, st);
}

}  // namespace plsb
