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
// `sm`: the CTA's work space -- shared memory, or (matrices too large for it) a slice of
// a global scratch buffer that stays L2 resident
__device__ void eigen_body(int r, double *sm, const double *__restrict__ G, int ldg,
                           long long g_stride, int K, int sqrt_lam, double *__restrict__ V_out,
                           double *__restrict__ lam_out) {
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  double *bufA = sm;                 // G -> diag(lam)
  double *bufV = bufA + ne * ld;     // V
  double *lam = bufV + ne * ld;      // ne
  JacobiScratch sc;
  sc.cst = lam + ne;                                   // 3*half
  sc.pq = reinterpret_cast<int *>(sc.cst + 3 * half);  // 2*half
  int *rank = sc.pq + 2 * half;                        // ne
  sc.blk = reinterpret_cast<short2 *>(rank + ne);      // half*(half+1)/2
  const int tid = threadIdx.x, nthr = blockDim.x;
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

__global__ void __launch_bounds__(SM_THREADS)
eigen_kernel(const double *__restrict__ G, int ldg, long long g_stride, int K, int sqrt_lam,
             double *__restrict__ V_out, double *__restrict__ lam_out,
             const int *__restrict__ todo, int count, double *gscratch, size_t gs_stride) {
  extern __shared__ __align__(16) double sm_dyn[];
  double *ws = gscratch ? gscratch + (size_t)blockIdx.x * gs_stride : sm_dyn;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    if (todo && !todo[r]) continue;   // done by the Newton-Schulz fast path
    __syncthreads();                  // the work space of the previous matrix is free
    eigen_body(r, ws, G, ldg, g_stride, K, sqrt_lam, V_out, lam_out);
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
// tall != 0 (more latent rows than features, the K x K problem is the FEATURE-side Gram
// matrix): H holds U_orig itself (shared by all resamples, h_stride = 0), temp = U_orig^T V
// without the d^-1 factor and M = V diag(sqrt(lam)) Q, i.e. compute.procrustes(U_orig,
// U_boot, d) = U_boot d Q itself (pyls/compute.py:260-262)
__device__ void rotation_body(int r, double *sm, const double *__restrict__ H,
                              const double *__restrict__ V_in, const double *__restrict__ lam_in,
                              int K, int L, const double *__restrict__ dorig,
                              double *__restrict__ M_out, int ldm, int tall, long long h_stride) {
  const int LP = (K + 7) & ~7, ld = LP + 4;
  double *Vs = sm;                // V          [LP][ld]
  double *Hs = Vs + LP * ld;      // H, later the second iterate
  double *Xs = Hs + LP * ld;      // temp = X_0
  double *Bs = Xs + LP * ld;      // (3 I - X^T X) / 2
  double *dinv = Bs + LP * ld;    // LP
  double *omask = dinv + LP;      // LP: 1 for non-null original latent variables
  double *dsc = omask + LP;       // LP: sqrt(lam) (tall mode), 0 for null directions
  const int tid = threadIdx.x;

  for (int e = tid; e < LP * LP; e += SM_THREADS) {
    const int i = e / LP, j = e - i * LP;
    const bool in = i < K && j < K;
    Vs[i * ld + j] = in ? V_in[(size_t)r * K * K + (size_t)i * K + j] : 0.0;
    Hs[i * ld + j] = in ? H[(size_t)r * h_stride + (size_t)i * L + j] : 0.0;
  }
  {
    // eigenvalues arrive sorted descending: lam[0] is the largest
    const double lmax = lam_in[(size_t)r * K];
    double domax = 0.0;
    if (dorig)
      for (int i = 0; i < L; ++i) domax = fmax(domax, dorig[i]);
    for (int i = tid; i < LP; i += SM_THREADS) {
      double di = 0.0, om = 0.0;
      if (i < K) {
        const double l = lam_in[(size_t)r * K + i];
        // d^-1 with a guard for numerically null directions
        const bool live = l > 1e-14 * lmax && l > 0.0;
        di = live ? (tall ? 1.0 : rsqrt(l)) : 0.0;
        // a numerically null ORIGINAL latent variable has no direction to rotate onto
        om = (!dorig || dorig[i] > 1e-10 * domax) ? 1.0 : 0.0;
        dsc[i] = live ? sqrt(l) : 0.0;
      } else {
        dsc[i] = 0.0;
      }
      dinv[i] = di;
      omask[i] = om;
    }
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
      for (int i = tid; i < LP; i += SM_THREADS) {
        double t = 0.0;
        for (int j = 0; j < LP; ++j) t += fabs((i == j ? 3.0 : 0.0) - 2.0 * Bs[i * ld + j]);
        rs = fmax(rs, t);
      }
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
  if (tall) {
    // M = V diag(d) Q: scale column i of X (= row i of Q) by d_i first
    __syncthreads();
    for (int e = tid; e < LP * LP; e += SM_THREADS) {
      const int j = e / LP, i = e - j * LP;
      X[j * ld + i] *= dsc[i];
    }
    __syncthreads();
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

__global__ void __launch_bounds__(SM_THREADS)
rotation_kernel(const double *__restrict__ H, const double *__restrict__ V_in,
                const double *__restrict__ lam_in, int K, int L,
                const double *__restrict__ dorig, double *__restrict__ M_out, int ldm,
                const int *__restrict__ todo, int count, double *gscratch, size_t gs_stride,
                int tall, long long h_stride) {
  extern __shared__ __align__(16) double sm_dyn[];
  double *ws = gscratch ? gscratch + (size_t)blockIdx.x * gs_stride : sm_dyn;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    if (todo && !todo[r]) continue;   // done by the Newton-Schulz fast path
    __syncthreads();
    rotation_body(r, ws, H, V_in, lam_in, K, L, dorig, M_out, ldm, tall, h_stride);
  }
}

// Fast path of the bootstrap rotation without an eigen-decomposition.
//
//   M = V Q  with  Q = polar(H^T V lam^-1/2)^T   is   M = polar(G^-1/2 H)
//
// (V polar(Y) = polar(V Y) for orthogonal V), and both factors come out of
// Newton-Schulz iterations -- DMMA matrix products only:
//   G^-1/2:  Y_0 = G / c, Z_0 = I;  T = (3 I - Z Y) / 2,  Y <- Y T,  Z <- T Z
//            (Z -> (G / c)^-1/2; c = |G|_inf >= lam_max, so |I - Y_0| < 1)
//   polar:   X_0 = Z H;  X <- X (3 I - X^T X) / 2     (sigma(X_0) <= 1: cosines)
// The first iteration converges in about log_2.25(cond(G)) + 5 steps; a
// resample whose G does not get there in NS_MAX_IT steps (ill-conditioned or
// rank-deficient bootstrap samples), or an analysis whose original
// decomposition has null latent variables (mean-centred PLS), is left to the
// Jacobi path: todo[r] = 1.
constexpr int NS_MAX_IT = 24;

__global__ void __launch_bounds__(SM_THREADS)
ns_rotation_kernel(const double *__restrict__ G, const double *__restrict__ H, int K, int L,
                   const double *__restrict__ dorig, double *__restrict__ M_out, int ldm,
                   int *__restrict__ todo) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[SM_WARPS];
  __shared__ int s_flag;
  const int LP = (K + 7) & ~7, ld = LP + 4;
  const int r = blockIdx.x, tid = threadIdx.x;
  double *Y = sm, *Z = Y + LP * ld, *T = Z + LP * ld, *S = T + LP * ld, *Hs = S + LP * ld;

  // null ORIGINAL latent variables take the other path (see rotation_kernel)
  if (tid == 0) {
    int bad = 0;
    if (dorig) {
      double domax = 0.0;
      for (int i = 0; i < L; ++i) domax = fmax(domax, dorig[i]);
      for (int i = 0; i < L; ++i) bad |= !(dorig[i] > 1e-10 * domax);
    }
    s_flag = bad;
  }
  // c = max row sum of |G|
  double rs = 0.0;
  if (tid < K)
    for (int j = 0; j < K; ++j) rs += fabs(G[(size_t)r * K * K + (size_t)tid * K + j]);
  for (int o = 16; o > 0; o >>= 1) rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = rs;
  __syncthreads();
  double c = 0.0;
  for (int w = 0; w < SM_WARPS; ++w) c = fmax(c, s_red[w]);
  if (s_flag || !(c > 0.0)) {
    if (tid == 0) todo[r] = 1;
    return;
  }
  const double ic = 1.0 / c;
  for (int e = tid; e < LP * LP; e += SM_THREADS) {
    const int i = e / LP, j = e - i * LP;
    const bool in = i < K && j < K;
    // the padding block is the identity: it stays the identity under the iteration
    Y[i * ld + j] = in ? 0.5 * ic * (G[(size_t)r * K * K + (size_t)i * K + j] +
                                     G[(size_t)r * K * K + (size_t)j * K + i])
                       : (i == j ? 1.0 : 0.0);
    Z[i * ld + j] = i == j ? 1.0 : 0.0;
    Hs[i * ld + j] = (in && j < L) ? H[(size_t)r * K * L + (size_t)i * L + j] : 0.0;
  }
  __syncthreads();
  bool ok = false;
  for (int it = 0; it < NS_MAX_IT; ++it) {
    double dev = 0.0;
    small_mm(Z, ld, 1, Y, ld, 1, LP, [&](int i, int j, double v) {
      const double t = (i == j ? 1.5 : 0.0) - 0.5 * v;
      T[i * ld + j] = t;
      dev = fmax(dev, fabs(t - (i == j ? 1.0 : 0.0)));
    });
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    __syncthreads();                       // s_red is free again, T is complete
    if ((tid & 31) == 0) s_red[tid >> 5] = dev;
    __syncthreads();
    dev = 0.0;
    for (int w = 0; w < SM_WARPS; ++w) dev = fmax(dev, s_red[w]);
    // |T - I| = (1 - x) / 2 for the slowest eigenvalue x of Z Y, which grows by 2.25 per
    // step while it is small and needs ~5 steps from 1/2 to rounding level: give up as
    // soon as that cannot fit (rank-deficient / ill-conditioned G: x_0 ~ 1e-16 .. 1e-7)
    const double x = fmax(1.0 - 2.0 * dev, 1e-300);
    const double need = x < 0.5 ? log(0.5 / x) * (1.0 / 0.8109302162163288) + 5.0 : 0.0;
    if (!(dev < 0.75) || it + need > NS_MAX_IT) break;
    small_mm(Y, ld, 1, T, ld, 1, LP, [&](int i, int j, double v) { S[i * ld + j] = v; });
    __syncthreads();
    // Z <- T Z into the buffer of the old Y
    small_mm(T, ld, 1, Z, ld, 1, LP, [&](int i, int j, double v) { Y[i * ld + j] = v; });
    __syncthreads();
    double *oldz = Z;
    Z = Y;
    Y = S;
    S = oldz;
    if (dev <= 1e-10) {
      ok = true;
      break;
    }
  }
  if (!ok) {
    if (tid == 0) todo[r] = 1;
    return;
  }
  if (tid == 0) todo[r] = 0;
  // X_0 = c^-1/2 Z H = G^-1/2 H
  double *X = S, *X2 = Y, *Bs = T;
  const double isc = sqrt(ic);
  small_mm(Z, ld, 1, Hs, ld, 1, LP, [&](int i, int j, double v) { X[i * ld + j] = v * isc; });
  __syncthreads();
  for (int it = 0; it < 100; ++it) {
    small_mm(X, 1, ld, X, ld, 1, LP, [&](int i, int j, double v) {
      Bs[i * ld + j] = (i == j ? 1.5 : 0.0) - 0.5 * v;
    });
    __syncthreads();
    if (it == 0) {
      // |X^T X|_inf bounds sigma_max^2; the iteration needs sigma_max < sqrt(3)
      double q = 0.0;
      if (tid < LP)
        for (int j = 0; j < LP; ++j) q += fabs((tid == j ? 3.0 : 0.0) - 2.0 * Bs[tid * ld + j]);
      for (int o = 16; o > 0; o >>= 1) q = fmax(q, __shfl_xor_sync(0xffffffffu, q, o));
      if ((tid & 31) == 0) s_red[tid >> 5] = q;
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
      X2[i * ld + j] = v;
      changed |= fabs(v - X[i * ld + j]) > 1e-14;
    });
    const int more = __syncthreads_or(changed);
    double *t = X;
    X = X2;
    X2 = t;
    if (!more) break;
  }
  for (int e = tid; e < K * ldm; e += SM_THREADS) {
    const int a = e / ldm, j = e - a * ldm;
    M_out[((size_t)r * K + a) * ldm + j] = j < L ? X[a * ld + j] : 0.0;
  }
}

// matrices whose work space exceeds this run out of a global (L2 resident) scratch slice
// per CTA instead of shared memory: any K, at global-memory speed
constexpr size_t SMALL_SMEM_MAX = 220 * 1024;

int launch_eigen(plsb_ctx *h, const double *G, int count, int K, int sqrt_lam, double *V,
                 double *lam, cudaStream_t st, int ldg = 0, long long g_stride = 0,
                 const int *todo = nullptr) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= 16384, PLSB_ERR_ARG, "small decomposition: K=%d", K);
  const int ne = K + (K & 1), ld = ne | 1, half = ne / 2;
  const int nb = half * (half + 1) / 2;
  const size_t smem = sizeof(double) * (2 * (size_t)ne * ld + ne + 3 * half) +
                      sizeof(int) * (2 * half + ne) + sizeof(short2) * nb + 16;
  // one thread per 2 x 2 block of a Jacobi round (measured: fewer, busier threads lose)
  const int threads = std::min(SM_THREADS, std::max(32, round_up(
      tune_int("PLSB_EIGEN_THREADS", nb), 32)));
  if (ldg <= 0) ldg = K;
  if (g_stride <= 0) g_stride = (long long)K * K;
  if (smem <= SMALL_SMEM_MAX) {
    PLSB_CUDA(cudaFuncSetAttribute(eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    eigen_kernel<<<count, threads, smem, st>>>(G, ldg, g_stride, K, sqrt_lam, V, lam, todo, count,
                                               nullptr, 0);
  } else {
    const size_t stride = (smem + 7) / 8;   // doubles
    const int ctas = std::min(count, 2 * h->sm_count);
    PLSB_TRY(h->big.ensure(sizeof(double) * stride * ctas));
    eigen_kernel<<<ctas, threads, 0, st>>>(G, ldg, g_stride, K, sqrt_lam, V, lam, todo, count,
                                           h->big.as<double>(), stride);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_rotation(plsb_ctx *h, const double *H, const double *V, const double *lam, int count,
                    int K, int L, const double *dorig, double *M, int ldm, cudaStream_t st,
                    const int *todo = nullptr, int tall = 0) {
  KernelTimer kt(h, KC_SMALL, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(L == K, PLSB_ERR_ARG, "small decomposition: L=%d must equal K=%d", L, K);
  const int LP = round_up(K, 8), ld = LP + 4;
  const size_t smem = sizeof(double) * (4 * (size_t)LP * ld + 3 * LP);
  const long long h_stride = tall ? 0 : (long long)K * L;
  if (smem <= SMALL_SMEM_MAX) {
    PLSB_CUDA(cudaFuncSetAttribute(rotation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rotation_kernel<<<count, SM_THREADS, smem, st>>>(H, V, lam, K, L, dorig, M, ldm, todo, count,
                                                     nullptr, 0, tall, h_stride);
  } else {
    const size_t stride = smem / 8;
    const int ctas = std::min(count, 2 * h->sm_count);
    PLSB_TRY(h->big.ensure(sizeof(double) * stride * ctas));
    rotation_kernel<<<ctas, SM_THREADS, 0, st>>>(H, V, lam, K, L, dorig, M, ldm, todo, count,
                                                 h->big.as<double>(), stride, tall, h_stride);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace

int launch_small_decomp(plsb_ctx *h, const double *G, const double *H, int count, int K, int L,
                        const double *dorig, double *M, int ldm, double *lam, cudaStream_t st) {
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(ldm >= L, PLSB_ERR_ARG, "small decomposition: output pitch %d < L=%d", ldm, L);
  const size_t kk = (size_t)K * K;
  PLSB_TRY(h->misc.ensure(sizeof(double) * (size_t)count * (kk + K) + sizeof(int) * (size_t)count));
  double *V = h->misc.as<double>(), *lam_s = V + (size_t)count * kk;
  // Newton-Schulz fast path (no eigen-decomposition); what it leaves -- and every
  // resample when the caller wants the eigenvalues -- takes the Jacobi path
  int *todo = nullptr;
  const int LP = round_up(K, 8);
  const size_t ns_smem = sizeof(double) * 5 * (size_t)LP * (LP + 4);
  if (!lam && L == K && ns_smem <= 200 * 1024 && tune_int("PLSB_NS_ROTATION", 1)) {
    KernelTimer kt(h, KC_SMALL, st);
    todo = reinterpret_cast<int *>(lam_s + (size_t)count * K);
    PLSB_CUDA(cudaFuncSetAttribute(ns_rotation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)ns_smem));
    ns_rotation_kernel<<<count, SM_THREADS, ns_smem, st>>>(G, H, K, L, dorig, M, ldm, todo);
    PLSB_LAUNCHED(h);
  }
  PLSB_TRY(launch_eigen(h, G, count, K, 0, V, lam_s, st, 0, 0, todo));
  PLSB_TRY(launch_rotation(h, H, V, lam_s, count, K, L, dorig, M, ldm, st, todo));
  if (lam)
    PLSB_CUDA(cudaMemcpyAsync(lam, lam_s, sizeof(double) * (size_t)count * K,
                              cudaMemcpyDeviceToDevice, st));
  return PLSB_OK;
}

// M[r] = V[r] diag(sqrt(lam[r])) polar(U_orig^T V[r])^T for the feature-side decomposition of
// analyses with more latent rows than features (K here = number of features = L)
int launch_rotation_tall(plsb_ctx *h, const double *Uorig, const double *V, const double *lam,
                         int count, int K, const double *dorig, double *M, cudaStream_t st) {
  return launch_rotation(h, Uorig, V, lam, count, K, K, dorig, M, K, st, nullptr, 1);
}

int launch_sym_eig(plsb_ctx *h, const double *G, int count, int K, double *V, double *lam,
                   int sqrt_lam, cudaStream_t st, int ldg, long long g_stride) {
  return launch_eigen(h, G, count, K, sqrt_lam, V, lam, st, ldg, g_stride);
}

}  // namespace plsb
