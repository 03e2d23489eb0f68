// Small dense decompositions, one CTA per resample, one WARP per column pair.
//
// The reference runs sklearn's randomized_svd on the (K, B) cross-covariance
// of every resample (pyls/compute.py:36-49) and a second SVD inside
// compute.procrustes (pyls/compute.py:260-262).  Both collapse onto K x K
// problems once G = R R^T and H = R U_orig are known:
//
//   G = V diag(lam) V^T                      one-sided Jacobi on the columns of G
//   d = sqrt(lam);  temp = H^T V d^-1        (= U_orig^T U_boot)
//   temp = N s Vr^T                          one-sided Jacobi on the columns of temp
//   Q = Vr N^T  (= P^T N^T of compute.py:262);   M = V Q
//
// Numerically null directions (mean-centred PLS always has one: the cell
// means minus their mean have rank J-1) are removed from the rotation: their
// d^-1 is 0 and the rows of temp that belong to null ORIGINAL latent variables
// are zeroed, so Q is the Procrustes rotation of the non-null subspace and the
// null columns of R^T M are 0.  (The reference rotates with whatever unit
// vectors its randomized SVD returns for the null directions, which perturbs
// the other latent variables by O(sqrt(K/B)); see DESIGN.md.)
//
// so that U_boot d Q = R^T M.  A column pair (p, q) is handled by one warp:
// lanes stride over the rows, the three inner products are reduced with warp
// shuffles, and the pairs of one round-robin round are independent, so the
// warps of a CTA rotate them concurrently.
#include "common.cuh"

namespace plsb {
namespace {

constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int MAX_SWEEPS = 40;
constexpr double JACOBI_TOL = 1e-14;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Orthogonalises the n columns of W (column j at W + j*ld) by plane rotations
// applied from the right; the same rotations are accumulated into Vacc.
// Called by every thread of the CTA.
__device__ void jacobi_onesided(double *W, double *Vacc, int n, int ld, int *s_flag) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ne = n + (n & 1);       // even number of players (last may be a dummy)
  const int half = ne / 2;
  __syncthreads();
  if (n < 2) return;
  for (int sweep = 0; sweep < MAX_SWEEPS; ++sweep) {
    __syncthreads();
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int round = 0; round < ne - 1; ++round) {
      for (int pi = warp; pi < half; pi += SM_WARPS) {
        int a, b;
        if (pi == 0) {
          a = ne - 1;
          b = round;
        } else {
          a = (round + pi) % (ne - 1);
          b = (round - pi + (ne - 1)) % (ne - 1);
        }
        const int p = min(a, b), q = max(a, b);
        if (q >= n) continue;
        double *wp = W + p * ld, *wq = W + q * ld;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = lane; i < n; i += 32) {
          const double x = wp[i], y = wq[i];
          alpha += x * x;
          beta += y * y;
          gamma += x * y;
        }
        alpha = warp_sum(alpha);
        beta = warp_sum(beta);
        gamma = warp_sum(gamma);
        if (alpha == 0.0 || beta == 0.0) continue;
        if (fabs(gamma) <= JACOBI_TOL * sqrt(alpha * beta)) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = lane; i < n; i += 32) {
          const double x = wp[i], y = wq[i];
          wp[i] = c * x - s * y;
          wq[i] = s * x + c * y;
        }
        double *vp = Vacc + p * ld, *vq = Vacc + q * ld;
        for (int i = lane; i < n; i += 32) {
          const double x = vp[i], y = vq[i];
          vp[i] = c * x - s * y;
          vq[i] = s * x + c * y;
        }
        if (lane == 0) *s_flag = 1;
      }
      __syncthreads();
    }
    if (*s_flag == 0) break;
  }
  __syncthreads();
}

// column norms of W into nrm[0..n), by warps
__device__ void col_norms(const double *W, int n, int ld, double *nrm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < n; j += SM_WARPS) {
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v += W[j * ld + i] * W[j * ld + i];
    v = warp_sum(v);
    if (lane == 0) nrm[j] = sqrt(v);
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
  double *bufA = sm;
  double *bufV = bufA + K * ld;
  double *bufT = bufV + K * ld;
  double *bufZ = bufT + K * ld;
  double *lam = bufZ + K * ld;   // K
  double *aux = lam + K;         // K
  int *rank = reinterpret_cast<int *>(aux + K);  // K
  __shared__ int s_flag;
  const int r = blockIdx.x, tid = threadIdx.x;
  const double *Gr = G + (size_t)r * K * K;

  for (int e = tid; e < K * K; e += SM_THREADS) {
    const int j = e / K, i = e - j * K;
    bufA[j * ld + i] = Gr[(size_t)i * K + j];
    bufV[j * ld + i] = (i == j) ? 1.0 : 0.0;
  }
  jacobi_onesided(bufA, bufV, K, ld, &s_flag);   // bufA = V diag(lam), bufV = V
  col_norms(bufA, K, ld, lam);
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
        const int j = e / K, i = e - j * K;
        V_out[(size_t)r * K * K + (size_t)i * K + rank[j]] = bufV[j * ld + i];
      }
    return;
  }

  // d^-1 with a guard for numerically null directions
  {
    double lmax = 0.0;
    for (int i = 0; i < K; ++i) lmax = fmax(lmax, lam[i]);
    __syncthreads();
    for (int j = tid; j < K; j += SM_THREADS)
      aux[j] = (lam[j] > 1e-14 * lmax && lam[j] > 0.0) ? 1.0 / sqrt(lam[j]) : 0.0;
  }
  // bufA <- H^T (bufA[i*ld + k] = H[k][i]);  bufZ <- I
  const double *Hr = H + (size_t)r * K * L;
  __syncthreads();
  for (int e = tid; e < K * L; e += SM_THREADS) {
    const int k = e / L, i = e - k * L;
    bufA[i * ld + k] = Hr[e];
  }
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int j = e / L, i = e - j * L;
    bufZ[j * ld + i] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  double dorig_max = 0.0;
  if (dorig)
    for (int i = 0; i < L; ++i) dorig_max = fmax(dorig_max, dorig[i]);
  // temp[i][j] = sum_k H[k][i] V[k][j] / d_j, stored column-major in bufT
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int j = e / L, i = e - j * L;
    double v = 0.0;
    for (int k = 0; k < K; ++k) v += bufA[i * ld + k] * bufV[j * ld + k];
    // a numerically null ORIGINAL latent variable has no direction to rotate onto
    if (dorig && !(dorig[i] > 1e-10 * dorig_max)) v = 0.0;
    bufT[j * ld + i] = v * aux[j];
  }
  jacobi_onesided(bufT, bufZ, L, ld, &s_flag);   // bufT = N diag(s), bufZ = Vr
  col_norms(bufT, L, ld, lam);                   // lam <- s
  {
    double smax = 0.0;
    for (int i = 0; i < L; ++i) smax = fmax(smax, lam[i]);
    __syncthreads();
    for (int j = tid; j < L; j += SM_THREADS)
      aux[j] = (lam[j] > 1e-12 * smax && lam[j] > 0.0) ? 1.0 / lam[j] : 0.0;
  }
  __syncthreads();
  // Q[i][j] = sum_k Vr[i][k] N[j][k]  -> bufA[j*ld + i]
  for (int e = tid; e < L * L; e += SM_THREADS) {
    const int j = e / L, i = e - j * L;
    double v = 0.0;
    for (int k = 0; k < L; ++k) v += bufZ[k * ld + i] * bufT[k * ld + j] * aux[k];
    bufA[j * ld + i] = v;
  }
  __syncthreads();
  // M[a][j] = sum_i V[a][i] Q[i][j]
  for (int e = tid; e < K * L; e += SM_THREADS) {
    const int a = e / L, j = e - a * L;
    double v = 0.0;
    for (int i = 0; i < K; ++i) v += bufV[i * ld + a] * bufA[j * ld + i];
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
