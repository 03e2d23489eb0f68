// Small dense decompositions, one CTA per resample.
//
// The reference runs sklearn's randomized_svd on the (K, B) cross-covariance
// of every resample (pyls/compute.py:36-49) and a second SVD inside
// compute.procrustes (pyls/compute.py:260-262).  Both collapse onto K x K
// symmetric eigenproblems / polar factors once G = R R^T and H = R U_orig are known:
//
//   G = V diag(lam) V^T                      Jacobi on G
//   d = sqrt(lam);  temp = H^T V d^-1        (= U_orig^T U_boot)
//   Q = polar(temp)^T                        (= P^T N^T of compute.py:262; Newton-Schulz)
//   M = V Q                                  so that U_boot d Q = R^T M
//
// The eigen-solver is the cyclic two-sided Jacobi method with a round-robin
// (tournament) ordering.  The n/2 plane rotations of one round act on disjoint
// index pairs, so the update A <- J^T A J decomposes into independent 2 x 2
// blocks, one per (pair a, pair b): a thread reads the four entries of its
// block, applies pair b's rotation from the right and pair a's from the left,
// and writes the block and its mirror image.  Only blocks with a <= b are
// computed (symmetry), the parameters of a rotation come from the diagonal
// block of its pair, and a round needs two barriers.
//
// Numerically null directions (mean-centred PLS always has one: the cell
// means minus their mean have rank J-1) are removed from the rotation: their
// d^-1 is 0 and the rows of temp that belong to null ORIGINAL latent variables
// are zeroed, so Q is the Procrustes rotation of the non-null subspace and the
// null columns of R^T M are 0.  (The reference rotates with whatever unit
// vectors its randomized SVD returns for the null directions, which perturbs
// the other latent variables; see DESIGN.md.)
#include "common.cuh"
#include "jacobi.cuh"

namespace plsb {
namespace {

constexpr int SM_THREADS = 256;
// mode 0: full decomposition -> M (K,L) [+ lam sorted descending if lam_out]
// mode 1: eigen only -> V_out (K,K) columns sorted by descending eigenvalue, lam_out sorted
//         (sqrt_lam != 0 writes sqrt(lam) instead: singular values of R).  Needs two
//         K x K tiles of shared memory only, so twice as many CTAs fit on an SM.
// mode 2: rotation only: V_out / lam_out are INPUTS (what mode 1 wrote) -> M (K,L)
// The drivers run mode 1 then mode 2: the Jacobi stage is latency bound and gains
// from the higher occupancy.
__global__ void __launch_bounds__(SM_THREADS)
small_decomp_kernel(const double *__restrict__ G, const double *__restrict__ H, int K, int L,
                    int mode, int sqrt_lam, const double *__restrict__ dorig,
                    double *__restrict__ M_out, double *__restrict__ V_out,
                    double *__restrict__ lam_out) {
  extern __shared__ __align__(16) double sm[];
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  const int nb = half * (half + 1) / 2;
  double *bufA = sm;                 // G, later temp^T temp, later N s
  double *bufV = bufA + ne * ld;     // V
  double *bufT = bufV + ne * ld;     // temp, later Q          (modes 0, 2 only)
  double *bufW = bufT + ne * ld;     // H, later W             (modes 0, 2 only)
  double *lam = bufA + (mode == 1 ? 2 : 4) * ne * ld;   // ne
  double *aux = lam + ne;            // ne
  JacobiScratch sc;
  sc.cst = aux + ne;                                   // 3*half
  sc.pq = reinterpret_cast<int *>(sc.cst + 3 * half);  // 2*half
  int *rank = sc.pq + 2 * half;                        // ne
  sc.blk = reinterpret_cast<short2 *>(rank + ne);      // nb
  const int r = blockIdx.x, tid = threadIdx.x;
  const double *Gr = G + (size_t)r * K * K;

  // block table: bi -> (a, b), a <= b
  for (int a = tid; a < half; a += SM_THREADS) {
    int bi = a * half - a * (a - 1) / 2;   // blocks of the rows before a
    for (int b = a; b < half; ++b) sc.blk[bi++] = make_short2((short)a, (short)b);
  }
  if (mode == 2) {
    for (int e = tid; e < K * K; e += SM_THREADS) {
      const int i = e / K, j = e - i * K;
      bufV[i * ld + j] = V_out[(size_t)r * K * K + e];
    }
    for (int j = tid; j < K; j += SM_THREADS) lam[j] = lam_out[(size_t)r * K + j];
    __syncthreads();
  } else {
  for (int e = tid; e < ne * ne; e += SM_THREADS) {
    const int i = e / ne, j = e - i * ne;
    double g = 0.0;
    // symmetrise: both triangles must agree exactly for the two-sided updates
    if (i < K && j < K) g = 0.5 * (Gr[(size_t)i * K + j] + Gr[(size_t)j * K + i]);
    bufA[i * ld + j] = g;
    bufV[i * ld + j] = (i == j && i < K) ? 1.0 : 0.0;
  }
  jacobi_sym(bufA, bufV, K, ld, sc);   // diag(bufA) = lam, bufV = V
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
  }
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
  // Orthogonal polar factor of temp by the Newton-Schulz iteration
  //   X <- X (3 I - X^T X) / 2,   X_0 = temp
  // (singular values of temp are cosines of principal angles, <= 1 < sqrt(3), so
  // it converges, quadratically at the end; exact-zero rows / columns -- the
  // null directions -- stay zero, which is the rotation of the non-null
  // subspace).  Only matrix products: no rotations, no barriers per pair.
  double *X = bufT, *Y = bufW;
  for (int it = 0; it < 100; ++it) {
    for (int e = tid; e < L * L; e += SM_THREADS) {
      const int i = e / L, j = e - i * L;
      if (j < i) continue;
      double v = 0.0;
      for (int k = 0; k < L; ++k) v += X[k * ld + i] * X[k * ld + j];
      bufA[i * ld + j] = v;
      bufA[j * ld + i] = v;
    }
    __syncthreads();
    if (it == 0) {
      // |X^T X|_inf bounds sigma_max^2; the iteration needs sigma_max < sqrt(3).
      // temp from orthonormal factors never triggers this (the polar factor
      // does not depend on a positive scale of X).
      for (int i = tid; i < L; i += SM_THREADS) {
        double rs = 0.0;
        for (int j = 0; j < L; ++j) rs += fabs(bufA[i * ld + j]);
        aux[i] = rs;
      }
      __syncthreads();
      double s = 0.0;
      for (int i = 0; i < L; ++i) s = fmax(s, aux[i]);
      if (s > 2.8) {
        const double f2 = 2.8 / s, f = sqrt(f2);
        for (int e = tid; e < L * L; e += SM_THREADS) {
          const int i = e / L, j = e - i * L;
          X[i * ld + j] *= f;
          bufA[i * ld + j] *= f2;
        }
      }
      __syncthreads();
    }
    int changed = 0;
    for (int e = tid; e < L * L; e += SM_THREADS) {
      const int i = e / L, j = e - i * L;
      double v = 0.0;
      for (int k = 0; k < L; ++k) v += X[i * ld + k] * bufA[k * ld + j];
      const double x = X[i * ld + j];
      const double y = 1.5 * x - 0.5 * v;
      Y[i * ld + j] = y;
      changed |= fabs(y - x) > 1e-14;
    }
    const int more = __syncthreads_or(changed);
    double *t = X;
    X = Y;
    Y = t;
    if (!more) break;
  }
  // Q = polar(temp)^T;  M[a][j] = sum_i V[a][i] Q[i][j] = sum_i V[a][i] X[j][i]
  for (int e = tid; e < K * L; e += SM_THREADS) {
    const int a = e / L, j = e - a * L;
    double v = 0.0;
    for (int i = 0; i < K; ++i) v += bufV[a * ld + i] * X[j * ld + i];
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
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  const int nb = half * (half + 1) / 2;
  const size_t smem = sizeof(double) * ((mode == 1 ? 2 : 4) * (size_t)ne * ld + 2 * ne + 3 * half) +
                      sizeof(int) * (2 * half + ne) + sizeof(short2) * nb + 16;
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
  if (count <= 0) return PLSB_OK;
  // stage 1 (eigen-decomposition, 2 tiles of shared memory) then stage 2 (rotation)
  const size_t kk = (size_t)K * K;
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)count * (kk + K)));
  double *V = h->misc.as<double>(), *lam_s = V + (size_t)count * kk;
  PLSB_TRY(launch_small(h, G, nullptr, count, K, K, 1, 0, nullptr, nullptr, V, lam_s, st));
  PLSB_TRY(launch_small(h, G, H, count, K, L, 2, 0, dorig, M, V, lam_s, st));
  if (lam)
    PLSB_CUDA(cudaMemcpyAsync(lam, lam_s, sizeof(double) * (size_t)count * K,
                              cudaMemcpyDeviceToDevice, st));
  return PLSB_OK;
}

int launch_sym_eig(plsb_ctx *h, const double *G, int count, int K, double *V, double *lam,
                   int sqrt_lam, cudaStream_t st) {
  return launch_small(h, G, nullptr, count, K, K, 1, sqrt_lam, nullptr, nullptr, V, lam, st);
}

}  // namespace plsb
