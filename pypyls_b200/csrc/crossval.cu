// Cross-validation of behavioural PLS (BehavioralPLS.crossval / _single_crossval,
// pyls/types/behavioral.py:82-170, and compute.rescale_test, pyls/compute.py:129-151).
//
// For a train / test split the reference decomposes the TRAINING cross-covariance
// R = U d V^T, standardises the test rows of X with the training statistics of
// their cell (zmap) and predicts  Y_pred = X_resc U V_cell^T + mean(Y_train).
// With P = X_resc R^T and G = R R^T this is
//
//     Y_pred = P G^-1/2 [:, columns of the cell] + mean(Y_train)
//
// (U V_cell^T = R^T V d^-1 V_cell^T and V d^-1 V^T = G^-1/2), so no B-sized
// singular vectors are needed: the engine stacks R (K rows) and the
// standardised test rows into one matrix per split, and a single gram_proj pass
// over it yields both G (leading K x K block) and P (rows K.., columns < K).
//
//   rescale_test  writes the standardised test rows under the K rows of R
//   cv_score      G^-1/2 from the eigen-decomposition, the prediction, and the
//                 Pearson r / R^2 of every behaviour against the held-out rows
#include "common.cuh"

namespace plsb {
namespace {

// one CTA per (split, test-row slot): row K + slot of the split's stacked matrix
__global__ void rescale_test_kernel(const int32_t *__restrict__ mask, int S, int B, int J, int K,
                                    int max_test, const int *__restrict__ cell_of_row,
                                    const double *__restrict__ Xglob, int ldx,
                                    const double *__restrict__ S1, const double *__restrict__ S2,
                                    int rpr, const int *__restrict__ ntrain,
                                    double *__restrict__ Z, int stride) {
  const int r = blockIdx.x / max_test, slot = blockIdx.x - r * max_test;
  __shared__ int s_row;
  if (threadIdx.x == 0) {
    int seen = 0, row = -1;
    for (int s = 0; s < S; ++s)
      if (mask[(size_t)r * S + s] == 0) {
        if (seen == slot) {
          row = s;
          break;
        }
        ++seen;
      }
    s_row = row;
  }
  __syncthreads();
  const int s = s_row;
  double *out = Z + ((size_t)r * stride + K + slot) * ldx;
  if (s < 0) {   // fewer test rows than slots: zero row
    for (int b = threadIdx.x; b < ldx; b += blockDim.x) out[b] = 0.0;
    return;
  }
  const int g = cell_of_row[s];
  const double n = (double)ntrain[(size_t)r * J + g];
  const double *s1 = S1 + ((size_t)r * rpr + g) * ldx, *s2 = S2 + ((size_t)r * rpr + g) * ldx;
  const double *x = Xglob + (size_t)s * ldx;
  for (int b = threadIdx.x; b < ldx; b += blockDim.x) {
    double v = 0.0;
    if (b < B) {
      // zmap(X_test, compare=X_train, ddof=1) from the sums over the training rows
      const double m = s1[b] / n;
      const double var = (s2[b] - s1[b] * m) / (n - 1.0);
      v = (x[b] - m) / sqrt(var);
    }
    out[b] = v;
  }
}

// one split; `sm`: work space (shared memory or a global scratch slice)
__device__ void cv_score_body(int r, double *sm, const int32_t *__restrict__ mask, int S, int T,
                              int J, int K, int max_test, const int *__restrict__ cell_of_row,
                              const double *__restrict__ Y, const double *__restrict__ Gz,
                              int stride, const double *__restrict__ V,
                              const double *__restrict__ lam, const double *__restrict__ ytrain,
                              double *__restrict__ r_out, double *__restrict__ r2_out) {
  double *W = sm;                       // K x K: G^-1/2
  double *pred = W + K * K;             // max_test x T
  double *dinv = pred + max_test * T;   // K
  int *rows = reinterpret_cast<int *>(dinv + K);   // max_test
  __shared__ int s_nte;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double *Vr = V + (size_t)r * K * K, *lr = lam + (size_t)r * K;
  const double *G = Gz + (size_t)r * stride * stride;

  if (tid == 0) {
    int n = 0;
    for (int s = 0; s < S; ++s)
      if (mask[(size_t)r * S + s] == 0 && n < max_test) rows[n++] = s;
    s_nte = n;
  }
  {
    const double lmax = lr[0];   // eigenvalues arrive sorted descending
    for (int i = tid; i < K; i += nt) {
      const double l = lr[i];
      // numerically null directions carry no prediction (cf. rotation_kernel)
      dinv[i] = (l > 1e-14 * lmax && l > 0.0) ? rsqrt(sqrt(l)) : 0.0;   // lam^-1/4
    }
  }
  __syncthreads();
  const int nte = s_nte;
  // W = V diag(lam^-1/2) V^T
  for (int e = tid; e < K * K; e += nt) {
    const int i = e / K, j = e - i * K;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
      const double d2 = dinv[k] * dinv[k];
      acc += Vr[i * K + k] * d2 * Vr[j * K + k];
    }
    W[e] = acc;
  }
  __syncthreads();
  // pred[j][t] = sum_k P[j][k] W[k][cell(j) * T + t] + mean(Y_train of the cell)[t]
  for (int e = tid; e < nte * T; e += nt) {
    const int j = e / T, t = e - j * T;
    const int g = cell_of_row[rows[j]];
    const double *P = G + (size_t)(K + j) * stride;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc += P[k] * W[k * K + g * T + t];
    pred[e] = acc + ytrain[((size_t)r * J + g) * T + t];
  }
  __syncthreads();
  // Pearson r (compute.efficient_corr) and R^2 (sklearn r2_score, raw values) per behaviour
  for (int t = tid; t < T; t += nt) {
    double my = 0.0, mp = 0.0;
    for (int j = 0; j < nte; ++j) {
      my += Y[(size_t)rows[j] * T + t];
      mp += pred[j * T + t];
    }
    my /= nte;
    mp /= nte;
    double syy = 0.0, spp = 0.0, syp = 0.0, sse = 0.0;
    for (int j = 0; j < nte; ++j) {
      const double y = Y[(size_t)rows[j] * T + t], p = pred[j * T + t];
      syy += (y - my) * (y - my);
      spp += (p - mp) * (p - mp);
      syp += (y - my) * (p - mp);
      sse += (y - p) * (y - p);
    }
    r_out[(size_t)r * T + t] = fmin(1.0, fmax(-1.0, syp / sqrt(syy * spp)));
    r2_out[(size_t)r * T + t] = 1.0 - sse / syy;
  }
}

__global__ void cv_score_kernel(const int32_t *__restrict__ mask, int S, int T, int J, int K,
                                int max_test, const int *__restrict__ cell_of_row,
                                const double *__restrict__ Y, const double *__restrict__ Gz,
                                int stride, const double *__restrict__ V,
                                const double *__restrict__ lam,
                                const double *__restrict__ ytrain, double *__restrict__ r_out,
                                double *__restrict__ r2_out, int count, double *gscratch,
                                size_t gs_stride) {
  extern __shared__ __align__(16) double sm_dyn[];
  double *ws = gscratch ? gscratch + (size_t)blockIdx.x * gs_stride : sm_dyn;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    __syncthreads();
    cv_score_body(r, ws, mask, S, T, J, K, max_test, cell_of_row, Y, Gz, stride, V, lam, ytrain,
                  r_out, r2_out);
  }
}

}  // namespace

int launch_rescale_test(plsb_ctx *h, const int32_t *mask, int count, int max_test,
                        const double *S1, const double *S2, int rows_per_resample,
                        const int *ntrain, double *Z, int stride, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (count <= 0 || max_test <= 0) return PLSB_OK;
  const Layout &l = h->lay;
  rescale_test_kernel<<<count * max_test, 256, 0, st>>>(
      mask, l.S, l.B, l.J, l.K, max_test, h->d_cell_of_row, h->Xglob.as<double>(), l.ldx, S1, S2,
      rows_per_resample, ntrain, Z, stride);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_cv_score(plsb_ctx *h, const int32_t *mask, int count, int max_test, const double *Gz,
                    int stride, const double *V, const double *lam, const double *ytrain,
                    double *r_out, double *r2_out, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (count <= 0) return PLSB_OK;
  const Layout &l = h->lay;
  const size_t smem = sizeof(double) * ((size_t)l.K * l.K + (size_t)max_test * l.T + l.K) +
                      sizeof(int) * (size_t)max_test + 16;
  if (smem <= 200 * 1024) {
    PLSB_CUDA(cudaFuncSetAttribute(cv_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    cv_score_kernel<<<count, 256, smem, st>>>(mask, l.S, l.T, l.J, l.K, max_test, h->d_cell_of_row,
                                              h->Y.as<double>(), Gz, stride, V, lam, ytrain, r_out,
                                              r2_out, count, nullptr, 0);
  } else {
    const size_t gs = (smem + 7) / 8;
    const int ctas = std::min(count, 2 * h->sm_count);
    PLSB_TRY(h->big.ensure(sizeof(double) * gs * ctas));
    cv_score_kernel<<<ctas, 256, 0, st>>>(mask, l.S, l.T, l.J, l.K, max_test, h->d_cell_of_row,
                                          h->Y.as<double>(), Gz, stride, V, lam, ytrain, r_out,
                                          r2_out, count, h->big.as<double>(), gs);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
