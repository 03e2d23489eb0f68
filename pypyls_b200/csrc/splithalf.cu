// Split-half reliability of the singular vectors (BasePLS.split_half,
// pyls/base.py:714-770; called once on the original data, base.py:373-380, and
// inside every permutation, base.py:704-708).
//
// The reference, for every half/half mask i of a (permuted) data set with
// decomposition R = U d V^T:
//     D1, D2  = gen_covcorr of either half               (K x B each)
//     ucorr_i = corr_columns(D1^T vd, D2^T vd)           vd = V d^-1    (B x L)
//     vcorr_i = corr_columns(D1 ud,   D2 ud)             ud = U d^-1    (K x L)
// and averages over the masks.  With U = R^T V d^-1 none of the B-sized
// projections has to be formed: for the stacked halves Z = [D1; D2] (2K x B)
//     Gz = Z Z^T                    (2K x 2K)
//     Hz = Z [R^T | 1]              (2K x (K + 1))
// give   (D_a^T vd_j) . (D_b^T vd_j) = vd_j^T Gz[a,b] vd_j,
//        sum_b (D_a^T vd_j)_b       = vd_j^T Hz[a, :, K],
//        D_a ud_j                   = Hz[a, :, :K] V d^-2 e_j,
// i.e. one gram_proj pass over Z with the permutation's own [R; 1^T; 0] block
// as projection target (gram_proj needs as many projection rows as Z has rows).
//
//   projblock_kernel   (n, K, ldx) stored R  ->  (n, 2K, ldx) blocks [R; 1^T; 0]
//   splithalf_score_kernel   one CTA per permutation: loops over the splits of
//                      this pass and accumulates mean ucorr / vcorr
//
// Numerically null latent variables (d_j <= 1e-7 d_max) have no direction: the
// reference divides by d_j ~ 1e-16 there and reports rounding noise; this
// kernel reports 0 for them (documented in DESIGN.md).
#include "common.cuh"

namespace plsb {
namespace {

__global__ void projblock_kernel(const double *__restrict__ R, int K, int B, long long ldx,
                                 long long n_elem, double *__restrict__ PB) {
  const long long per = 2ll * K * ldx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_elem;
       e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / per, rem = e - p * per;
    const int row = (int)(rem / ldx), c = (int)(rem - (long long)row * ldx);
    double v = 0.0;
    if (row < K)
      v = R[(p * K + row) * ldx + c];
    else if (row == K && c < B)
      v = 1.0;
    PB[e] = v;
  }
}

// G, H: (n_perm * ns, 2K, 2K) from gram_proj; V (K,K) eigenvectors in columns and d (K)
// singular values per permutation (v_stride / d_stride 0: shared by all)
// `sm`: the CTA's work space (shared memory, or a slice of a global scratch buffer when
// 11 K^2 doubles do not fit); thread loops run over the latent variables, so any K works
__device__ void splithalf_score_body(int p, double *sm, const double *__restrict__ G,
                                     const double *__restrict__ H, int K, int B, int ns,
                                     const double *__restrict__ V, long long v_stride,
                                     const double *__restrict__ d, long long d_stride,
                                     double inv_nsplit, double *__restrict__ ucorr,
                                     double *__restrict__ vcorr) {
  const int K2 = 2 * K, tid = threadIdx.x, nt = blockDim.x;
  double *vd = sm;                    // K*K   V d^-1
  double *w = vd + K * K;             // K*K   V d^-2
  double *Gs = w + K * K;             // K2*K2
  double *Hs = Gs + K2 * K2;          // K2*(K+1)
  double *Q = Hs + K2 * (K + 1);      // 3*K*K partial quadratic forms, then b1 | b2
  __shared__ double s_dmax;
  const double *Vp = V + (size_t)p * v_stride, *dp = d + (size_t)p * d_stride;
  if (tid == 0) {
    double m = 0.0;
    for (int j = 0; j < K; ++j) m = fmax(m, dp[j]);
    s_dmax = m;
  }
  __syncthreads();
  for (int e = tid; e < K * K; e += nt) {
    const int j = e % K;
    const double dj = dp[j];
    const double di = (dj > 1e-7 * s_dmax && dj > 0.0) ? 1.0 / dj : 0.0;
    vd[e] = Vp[e] * di;
    w[e] = Vp[e] * di * di;
  }
  double *uc_all = Q + 3 * (size_t)K * K;   // K: ucorr of the current split
  for (int i = 0; i < ns; ++i) {
    const double *Gp = G + ((size_t)p * ns + i) * K2 * K2;
    const double *Hp = H + ((size_t)p * ns + i) * K2 * K2;
    __syncthreads();
    for (int e = tid; e < K2 * K2; e += nt) Gs[e] = Gp[e];
    for (int e = tid; e < K2 * (K + 1); e += nt) {
      const int a = e / (K + 1), c = e - a * (K + 1);
      Hs[e] = Hp[(size_t)a * K2 + c];
    }
    __syncthreads();
    // partial quadratic forms: Q[0] -> q11, Q[1] -> q12, Q[2] -> q22
    for (int e = tid; e < K * K; e += nt) {
      const int a = e / K, j = e - a * K;
      double p1 = 0.0, p2 = 0.0, p3 = 0.0;
      for (int b = 0; b < K; ++b) {
        const double v = vd[b * K + j];
        p1 += Gs[a * K2 + b] * v;
        p2 += Gs[a * K2 + K + b] * v;
        p3 += Gs[(K + a) * K2 + K + b] * v;
      }
      const double va = vd[a * K + j];
      Q[e] = va * p1;
      Q[K * K + e] = va * p2;
      Q[2 * K * K + e] = va * p3;
    }
    __syncthreads();
    for (int j = tid; j < K; j += nt) {
      double q11 = 0.0, q12 = 0.0, q22 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int a = 0; a < K; ++a) {
        q11 += Q[a * K + j];
        q12 += Q[K * K + a * K + j];
        q22 += Q[2 * K * K + a * K + j];
        s1 += vd[a * K + j] * Hs[a * (K + 1) + K];
        s2 += vd[a * K + j] * Hs[(K + a) * (K + 1) + K];
      }
      const double c11 = q11 - s1 * s1 / B, c22 = q22 - s2 * s2 / B, c12 = q12 - s1 * s2 / B;
      uc_all[j] = c12 / sqrt(c11 * c22);
    }
    __syncthreads();
    // b1 = H1 w, b2 = H2 w   (K x K each), stored in Q
    for (int e = tid; e < K2 * K; e += nt) {
      const int a = e / K, j = e - a * K;
      double v = 0.0;
      for (int c = 0; c < K; ++c) v += Hs[a * (K + 1) + c] * w[c * K + j];
      Q[e] = v;
    }
    __syncthreads();
    for (int j = tid; j < K; j += nt) {
      double m1 = 0.0, m2 = 0.0;
      for (int a = 0; a < K; ++a) {
        m1 += Q[a * K + j];
        m2 += Q[(K + a) * K + j];
      }
      m1 /= K;
      m2 /= K;
      double s11 = 0.0, s22 = 0.0, s12 = 0.0;
      for (int a = 0; a < K; ++a) {
        const double x = Q[a * K + j] - m1, y = Q[(K + a) * K + j] - m2;
        s11 += x * x;
        s22 += y * y;
        s12 += x * y;
      }
      const bool null_lv = !(dp[j] > 1e-7 * s_dmax && dp[j] > 0.0);
      const double vc = s12 / sqrt(s11 * s22);
      // compute.efficient_corr clips rounding overshoot (pyls/compute.py:389); splits are
      // added in order, one owner thread per latent variable: deterministic
      if (!null_lv) {
        ucorr[(size_t)p * K + j] += fmin(1.0, fmax(-1.0, uc_all[j])) * inv_nsplit;
        vcorr[(size_t)p * K + j] += fmin(1.0, fmax(-1.0, vc)) * inv_nsplit;
      }
    }
  }
}

// G, H: (n_perm * ns, 2K, 2K) from gram_proj; V (K,K) eigenvectors in columns and d (K)
// singular values per permutation (v_stride / d_stride 0: shared by all)
__global__ void __launch_bounds__(256)
splithalf_score_kernel(const double *__restrict__ G, const double *__restrict__ H, int K, int B,
                       int ns, const double *__restrict__ V, long long v_stride,
                       const double *__restrict__ d, long long d_stride, double inv_nsplit,
                       double *__restrict__ ucorr, double *__restrict__ vcorr, int n_perm,
                       double *gscratch, size_t gs_stride) {
  extern __shared__ __align__(16) double sm_dyn[];
  double *ws = gscratch ? gscratch + (size_t)blockIdx.x * gs_stride : sm_dyn;
  for (int p = blockIdx.x; p < n_perm; p += gridDim.x) {
    __syncthreads();
    splithalf_score_body(p, ws, G, H, K, B, ns, V, v_stride, d, d_stride, inv_nsplit, ucorr,
                         vcorr);
  }
}

}  // namespace

int launch_projblock(plsb_ctx *h, const double *R, int n, double *PB, cudaStream_t st) {
  KernelTimer kt(h, KC_PREP, st);
  const Layout &l = h->lay;
  const long long n_elem = (long long)n * 2 * l.K * l.ldx;
  if (n_elem <= 0) return PLSB_OK;
  const int blocks = (int)std::min<long long>((n_elem + 255) / 256, (long long)h->sm_count * 32);
  projblock_kernel<<<blocks, 256, 0, st>>>(R, l.K, l.B, l.ldx, n_elem, PB);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_splithalf_score(plsb_ctx *h, const double *G, const double *H, int n_perm, int ns,
                           const double *V, long long v_stride, const double *d,
                           long long d_stride, int n_split, double *ucorr, double *vcorr,
                           cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_perm <= 0 || ns <= 0) return PLSB_OK;
  const Layout &l = h->lay;
  const int K = l.K, K2 = 2 * K;
  const size_t smem = sizeof(double) * (2 * (size_t)K * K + (size_t)K2 * K2 +
                                       (size_t)K2 * (K + 1) + 3 * (size_t)K * K + K);
  if (smem <= 200 * 1024) {
    PLSB_CUDA(cudaFuncSetAttribute(splithalf_score_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    splithalf_score_kernel<<<n_perm, 256, smem, st>>>(G, H, K, l.B, ns, V, v_stride, d, d_stride,
                                                      1.0 / n_split, ucorr, vcorr, n_perm,
                                                      nullptr, 0);
  } else {
    // beyond shared memory: a global (L2 resident) work space per CTA
    const size_t stride = smem / sizeof(double);
    const int ctas = std::min(n_perm, 2 * h->sm_count);
    PLSB_TRY(h->big.ensure(sizeof(double) * stride * ctas));
    splithalf_score_kernel<<<ctas, 256, 0, st>>>(G, H, K, l.B, ns, V, v_stride, d, d_stride,
                                                 1.0 / n_split, ucorr, vcorr, n_perm,
                                                 h->big.as<double>(), stride);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
