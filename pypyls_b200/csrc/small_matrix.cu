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
constexpr double JACOBI_TOL = 1e-15;

// Diagonalises the symmetric n x n matrix A (row-major, leading dimension ld,
// both triangles stored) in place: A <- J^T A J, V <- V J over all rotations.
// pq[2*half] and cs[2*half] are scratch.  Called by every thread of the CTA.
__device__ void jacobi_sym(double *A, double *V, int n, int ld, int *pq, double *cs, int *s_flag) {
  const int tid = threadIdx.x;
  const int ne = n + (n & 1);       // even number of players (last may be a dummy)
  const int half = ne / 2;
  __syncthreads();
  if (n < 2) return;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int round = 0; round < ne - 1; ++round) {
      // phase 0: rotation parameters of the pairs of this round
      for (int pi = tid; pi < half; pi += SM_THREADS) {
        int a, b;
        if (pi == 0) {
          a = ne - 1;
          b = round;
        } else {
          a = (round + pi) % (ne - 1);
          b = (round - pi + (ne - 1)) % (ne - 1);
        }
        const int p = min(a, b), q = max(a, b);
        double c = 1.0, s = 0.0;
        int active = 0;
        if (q < n) {
          const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
          if (apq != 0.0 && fabs(apq) > JACOBI_TOL * sqrt(fabs(app * aqq))) {
            const double zeta = (aqq - app) / (2.0 * apq);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = c * t;
            active = 1;
          }
        }
        pq[2 * pi] = active ? p : -1;
        pq[2 * pi + 1] = q;
        cs[2 * pi] = c;
        cs[2 * pi + 1] = s;
        if (active) *s_flag = 1;
      }
      __syncthreads();
      // phase 1: column rotations  A <- A J,  V <- V J
      for (int e = tid; e < half * n; e += SM_THREADS) {
        const int pi = e / n, i = e - pi * n;
        const int p = pq[2 * pi];
        if (p < 0) continue;
        const int q = pq[2 * pi + 1];
        const double c = cs[2 * pi], s = cs[2 * pi + 1];
        const double x = A[i * ld + p], y = A[i * ld + q];
        A[i * ld + p] = c * x - s * y;
        A[i * ld + q] = s * x + c * y;
        const double vx = V[i * ld + p], vy = V[i * ld + q];
        V[i * ld + p] = c * vx - s * vy;
        V[i * ld + q] = s * vx + c * vy;
      }
      __syncthreads();
      // phase 2: row rotations  A <- J^T A
      for (int e = tid; e < half * n; e += SM_THREADS) {
        const int pi = e / n, j = e - pi * n;
        const int p = pq[2 * pi];
        if (p < 0) continue;
        const int q = pq[2 * pi + 1];
        const double c = cs[2 * pi], s = cs[2 * pi + 1];
        const double x = A[p * ld + j], y = A[q * ld + j];
        A[p * ld + j] = c * x - s * y;
        A[q * ld + j] = s * x + c * y;
      }
      __syncthreads();
    }
    if (*s_flag == 0) break;
    __syncthreads();
  }
  __syncthreads();
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
  double *cs = aux + K;              // K + 2
  int *rank = reinterpret_cast<int *>(cs + K + 2);  // K
  int *pq = rank + K;                // K + 2
  __shared__ int s_flag;
  const int r = blockIdx.x, tid = threadIdx.x;
  const double *Gr = G + (size_t)r * K * K;

  for (int e = tid; e < K * K; e += SM_THREADS) {
    const int i = e / K, j = e - i * K;
    // symmetrise: both triangles must agree exactly for the two-sided updates
    bufA[i * ld + j] = 0.5 * (Gr[(size_t)i * K + j] + Gr[(size_t)j * K + i]);
    bufV[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  jacobi_sym(bufA, bufV, K, ld, pq, cs, &s_flag);   // diag(bufA) = lam, bufV = V
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
  jacobi_sym(bufA, bufW, L, ld, pq, cs, &s_flag);   // bufW = W (right singular vectors)
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
  const size_t smem =
      sizeof(double) * (4 * (size_t)K * ld + 3 * K + 2) + sizeof(int) * (2 * K + 2);
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
