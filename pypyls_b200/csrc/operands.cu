// Per-resample left operands of the cross-covariance contraction.
//
// The reference gathers rows of X and Y for every resample and then forms
// `Yn.T @ Xn` (pyls/base.py:569,599; pyls/compute.py:84-92).  Here the row
// gather is moved to the small operand: for a resample with source rows
// src[s], X[src]^T Z = X^T W with W[u,:] = sum_{s: src[s]=u} Z[s,:], so every
// resample is a few rows of a tall matrix A that multiplies the SAME data
// matrix.  One CTA builds the rows of one resample.
//
//   behavioural (pyls/types/behavioral.py:27-52)
//     ROT    A[r*L+j, u] = sum_t zy[u,t] V[(cell(u),t), j] / (n-1)   (rotated perms:
//            |R^T v_j| = |Xcell^T a_j|, pyls/base.py:696-700)
//     PLAIN  A[r*K+(g,t), u] = [u in g] zy[u,t] / (n-1)              (R itself)
//     BOOT   A[r*K+(g,t), u] = sum_{s in g, src[s]=u} zy[s,t]        (x 1/(n-1) for covariance)
//            Ac[r*J+g, u]    = #{s in g : src[s]=u}                  (column statistics)
//            distrib[r]      = per-cell xcorr(Sx[src], Y[src])       (behavioral.py:54-80)
//     TRAIN  like BOOT for a train / test split (cross-validation,
//            pyls/types/behavioral.py:125-170): idx[s] != 0 marks a training row;
//            all cell statistics run over the training rows only, and the
//            per-cell training counts and Y means are written out
//     HALF   one half of a split-half mask (pyls/base.py:714-770) of permuted data:
//            operand row block r is half (r & 1) of mask r >> 1; the rows of the half
//            are a TRAIN set whose behaviours come through the permutation `yidx`
//     zy = Y[src] z-scored (ddof=1) or centred within each cell; permutations may
//     instead bring their own Y (pre-permuted matrices, pyls/base.py:636-639, 689-692).
//   mean-centred (pyls/types/meancentered.py:50-125, pyls/compute.py:267-357)
//     AR[j,u] = sum_{s: src[s]=u} C[j,s];  ROT: A = V^T AR;  PLAIN/BOOT: A = AR;
//     distrib[r] = AR @ Sx.  HALF: C is rebuilt from the rows of the half (cell mean of
//     the half minus the half's centring mean) and scattered through the permutation.
#include "common.cuh"

namespace plsb {
namespace {

struct BuildParams {
  const int32_t *idx;
  int S, T, J, K, L, lda, corr, kind;
  long long cellpad_w, cellpad_c;   // > 0: rows grouped by cell, cell g starts at g * cellpad
  const double *Y;
  const double *Yperm;              // optional (count, S, T): resample r uses Yperm[r] as its Y
  const int *cell_start, *cell_of_row;
  const double *Vo, *Sx, *Cmat;
  double *A, *Ac, *distrib;
  int *ntrain;        // TRAIN / HALF: (count, J) training rows per cell
  double *ytrain;     // TRAIN: (count, J, T) mean of Y over the training rows of every cell
  // HALF: idx = masks (., n_split, S); operand block r -> mask (r >> 1): permutation
  // (r >> 1) / half_ns, split half_s0 + (r >> 1) % half_ns; side r & 1 (0: mask != 0)
  const int32_t *yidx;   // optional (n_perm, S) permutation table of the data being split
  int half_ns, half_s0, half_nsplit;
  int n_cond, mean_centering;
};

// is position s inside half r, and which permutation does half r belong to
__device__ __forceinline__ bool half_member(const BuildParams &p, int r, int s, int *perm) {
  const int m = r >> 1, pl = m / p.half_ns, sp = p.half_s0 + (m - pl * p.half_ns);
  *perm = pl;
  const bool v = p.idx[((size_t)pl * p.half_nsplit + sp) * p.S + s] != 0;
  return (r & 1) ? !v : v;
}

__global__ void build_behavioral_kernel(BuildParams p) {
  extern __shared__ __align__(16) double sm[];
  const int S = p.S, T = p.T, J = p.J, K = p.K, L = p.L, lda = p.lda;
  double *Yp = sm;                 // S*T
  double *ymean = Yp + S * T;      // J*T
  double *yistd = ymean + J * T;   // J*T
  double *sxm = yistd + J * T;     // J*L
  double *sxi = sxm + J * L;       // J*L
  int *src = reinterpret_cast<int *>(sxi + J * L);  // S
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;

  const bool half = p.kind == BUILD_HALF;
  const bool train = p.kind == BUILD_TRAIN || half;
  int yr = r;                            // which pre-permuted Y / permutation this block uses
  for (int s = tid; s < S; s += nt) {
    int v;
    if (half) {
      v = half_member(p, r, s, &yr) ? s : -1;
    } else {
      v = p.idx ? p.idx[(size_t)r * S + s] : s;
      if (train) v = v != 0 ? s : -1;    // a training row is its own source, a test row has none
    }
    src[s] = v;
  }
  if (half) yr = (r >> 1) / p.half_ns;
  __syncthreads();
  for (int e = tid; e < S * T; e += nt) {
    const int s = e / T, t = e - s * T;
    double y = 0.0;
    if (src[s] >= 0) {
      if (p.Yperm)
        y = p.Yperm[((size_t)yr * S + s) * T + t];
      else if (half)
        y = p.Y[(size_t)(p.yidx ? p.yidx[(size_t)yr * S + s] : s) * T + t];
      else
        y = p.Y[(size_t)src[s] * T + t];
    }
    Yp[e] = y;
  }
  __syncthreads();
  // cell statistics over the rows that have a source (all of them except in TRAIN)
  for (int c = tid; c < J * T; c += nt) {
    const int g = c / T, t = c - g * T;
    const int r0 = p.cell_start[g], r1 = p.cell_start[g + 1];
    int n = 0;
    double m = 0.0;
    for (int s = r0; s < r1; ++s)
      if (src[s] >= 0) {
        m += Yp[s * T + t];
        ++n;
      }
    m /= n;
    double v = 0.0;
    for (int s = r0; s < r1; ++s)
      if (src[s] >= 0) {
        const double d = Yp[s * T + t] - m;
        v += d * d;
      }
    ymean[c] = m;
    yistd[c] = p.corr ? 1.0 / sqrt(v / (n - 1)) : 1.0;
    if (train) {
      if (p.ytrain) p.ytrain[(size_t)r * J * T + c] = m;
      if (t == 0) p.ntrain[(size_t)r * J + g] = n;
    }
  }
  __syncthreads();
  for (int e = tid; e < S * T; e += nt) {
    const int s = e / T, t = e - s * T;
    const int c = p.cell_of_row[s] * T + t;
    Yp[e] = src[s] >= 0 ? (Yp[e] - ymean[c]) * yistd[c] : 0.0;
  }
  __syncthreads();

  if (p.kind == BUILD_ROT) {
    double *Ar = p.A + (size_t)r * L * lda;
    for (int e = tid; e < L * lda; e += nt) {
      const int j = e / lda, u = e - j * lda;
      double val = 0.0;
      if (u < S) {
        const int g = p.cell_of_row[u];
        const int n = p.cell_start[g + 1] - p.cell_start[g];
        for (int t = 0; t < T; ++t) val += Yp[u * T + t] * p.Vo[(size_t)(g * T + t) * L + j];
        val /= (n - 1);
      }
      Ar[e] = val;
    }
    return;
  }
  if (p.kind == BUILD_PLAIN) {
    for (int e = tid; e < K * lda; e += nt) {
      const int row = e / lda, u = e - row * lda;
      const int g = row / T, t = row - g * T;
      double val = 0.0;
      if (u < S && p.cell_of_row[u] == g) {
        const int n = p.cell_start[g + 1] - p.cell_start[g];
        val = Yp[u * T + t] / (n - 1);
      }
      const size_t orow = p.cellpad_w ? (size_t)g * p.cellpad_w + (size_t)r * T + t
                                      : (size_t)r * K + row;
      p.A[orow * lda + u] = val;
    }
    return;
  }

  // ---- BUILD_BOOT ----  (A == nullptr: only the bootstrap distribution is wanted)
  if (p.A) {
    for (int e = tid; e < K * lda; e += nt) {
      const int row = e / lda, u = e - row * lda;
      const int g = row / T, t = row - g * T;
      const int r0 = p.cell_start[g], r1 = p.cell_start[g + 1];
      double val = 0.0;
      int n = 0;
      if (u < S)
        for (int s = r0; s < r1; ++s) {
          if (src[s] == u) val += Yp[s * T + t];
          n += src[s] >= 0;
        }
      if (!p.corr) val /= (n - 1);
      const size_t orow = p.cellpad_w ? (size_t)g * p.cellpad_w + (size_t)r * T + t
                                      : (size_t)r * K + row;
      p.A[orow * lda + u] = val;
    }
    if (p.Ac) {
      for (int e = tid; e < J * lda; e += nt) {
        const int g = e / lda, u = e - g * lda;
        const int r0 = p.cell_start[g], r1 = p.cell_start[g + 1];
        int cnt = 0;
        if (u < S)
          for (int s = r0; s < r1; ++s) cnt += (src[s] == u);
        const size_t orow = p.cellpad_c ? (size_t)g * p.cellpad_c + r : (size_t)r * J + g;
        p.Ac[orow * lda + u] = (double)cnt;
      }
    }
  }
  if (!p.distrib) return;
  // per-cell statistics of the gathered scores Sx[src]
  for (int c = tid; c < J * L; c += nt) {
    const int g = c / L, l = c - g * L;
    const int r0 = p.cell_start[g], r1 = p.cell_start[g + 1], n = r1 - r0;
    double m = 0.0;
    for (int s = r0; s < r1; ++s) m += p.Sx[(size_t)src[s] * L + l];
    m /= n;
    double v = 0.0;
    for (int s = r0; s < r1; ++s) {
      const double d = p.Sx[(size_t)src[s] * L + l] - m;
      v += d * d;
    }
    sxm[c] = m;
    sxi[c] = p.corr ? 1.0 / sqrt(v / (n - 1)) : 1.0;
  }
  __syncthreads();
  double *Dr = p.distrib + (size_t)r * K * L;
  for (int e = tid; e < K * L; e += nt) {
    const int row = e / L, l = e - row * L;
    const int g = row / T, t = row - g * T;
    const int r0 = p.cell_start[g], r1 = p.cell_start[g + 1], n = r1 - r0;
    const double m = sxm[g * L + l], is = sxi[g * L + l];
    double acc = 0.0;
    for (int s = r0; s < r1; ++s)
      acc += Yp[s * T + t] * ((p.Sx[(size_t)src[s] * L + l] - m) * is);
    Dr[e] = acc / (n - 1);
  }
}

__global__ void build_meancentered_kernel(BuildParams p) {
  extern __shared__ __align__(16) double sm[];
  const int S = p.S, J = p.J, L = p.L, lda = p.lda;
  double *AR = sm;                                   // J*S
  int *src = reinterpret_cast<int *>(AR + J * S);    // S
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (p.kind == BUILD_HALF) {
    // rows of the half: cell means of the half minus the half's centring mean
    // (compute.get_mean_center on X[perm][half], pyls/compute.py:267-357), as an
    // operator on the rows of X
    __shared__ int nh[MAX_K + 1];          // rows of the half per cell, [J] their total
    int *inh = src + S;                    // S flags
    int pl = 0;
    for (int s = tid; s < S; s += nt) {
      inh[s] = half_member(p, r, s, &pl) ? 1 : 0;
      src[s] = p.yidx ? p.yidx[(size_t)((r >> 1) / p.half_ns) * S + s] : s;
    }
    __syncthreads();
    for (int j = tid; j <= J; j += nt) {
      int n = 0;
      const int r0 = j < J ? p.cell_start[j] : 0, r1 = j < J ? p.cell_start[j + 1] : S;
      for (int s = r0; s < r1; ++s) n += inh[s];
      nh[j] = n;
    }
    __syncthreads();
    const int n_cond = p.n_cond, n_groups = J / n_cond;
    for (int e = tid; e < J * S; e += nt) AR[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < J * S; e += nt) {
      const int j = e / S, s = e - j * S;
      if (!inh[s]) continue;
      const int cs = p.cell_of_row[s];
      double val = cs == j ? 1.0 / nh[j] : 0.0;
      if (p.mean_centering == 0) {
        const int g = j / n_cond;
        if (cs / n_cond == g) {
          int ng = 0;
          for (int c = 0; c < n_cond; ++c) ng += nh[g * n_cond + c];
          val -= 1.0 / ng;
        }
      } else if (p.mean_centering == 1) {
        if (cs % n_cond == j % n_cond) val -= 1.0 / ((double)n_groups * nh[cs]);
      } else {
        val -= 1.0 / nh[J];
      }
      AR[j * S + src[s]] = val;            // src is a permutation: one writer per element
    }
    __syncthreads();
    double *Ar = p.A + (size_t)r * J * lda;
    for (int e = tid; e < J * lda; e += nt) {
      const int j = e / lda, u = e - j * lda;
      Ar[e] = u < S ? AR[j * S + u] : 0.0;
    }
    return;
  }
  for (int s = tid; s < S; s += nt) src[s] = p.idx ? p.idx[(size_t)r * S + s] : s;
  __syncthreads();
  for (int e = tid; e < J * S; e += nt) {
    const int j = e / S, u = e - j * S;
    double val = 0.0;
    for (int s = 0; s < S; ++s)
      if (src[s] == u) val += p.Cmat[(size_t)j * S + s];
    AR[e] = val;
  }
  __syncthreads();
  if (p.kind == BUILD_ROT) {
    double *Ar = p.A + (size_t)r * L * lda;
    for (int e = tid; e < L * lda; e += nt) {
      const int l = e / lda, u = e - l * lda;
      double val = 0.0;
      if (u < S)
        for (int j = 0; j < J; ++j) val += p.Vo[(size_t)j * L + l] * AR[j * S + u];
      Ar[e] = val;
    }
    return;
  }
  if (p.A) {
    double *Ar = p.A + (size_t)r * J * lda;
    for (int e = tid; e < J * lda; e += nt) {
      const int j = e / lda, u = e - j * lda;
      Ar[e] = u < S ? AR[j * S + u] : 0.0;
    }
  }
  if (p.kind == BUILD_BOOT && p.distrib) {
    double *Dr = p.distrib + (size_t)r * J * L;
    for (int e = tid; e < J * L; e += nt) {
      const int j = e / L, l = e - j * L;
      double acc = 0.0;
      for (int u = 0; u < S; ++u) acc += AR[j * S + u] * p.Sx[(size_t)u * L + l];
      Dr[e] = acc;
    }
  }
}

// row_map / kranges of a cell-grouped operand: operand row g*cellpad + i (i =
// r*rows_pc + t) is output row r*stride_r + g*rows_pc + t; rows i >= n*rows_pc
// are padding (-1).  Every 128-row tile lies inside one cell and contracts
// only over that cell's rows of the data matrix.
__global__ void build_maps_kernel(int n, int rows_pc, int stride_r, int J, long long cellpad,
                                  const int4 *__restrict__ cell_kr, int *__restrict__ row_map,
                                  int4 *__restrict__ kranges) {
  const long long total = (long long)J * cellpad;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(m / cellpad);
    const long long i = m - g * cellpad;
    int out = -1;
    if (i < (long long)n * rows_pc) {
      const int r = (int)(i / rows_pc), t = (int)(i - (long long)r * rows_pc);
      out = r * stride_r + g * rows_pc + t;
    }
    row_map[m] = out;
    if (m % GEMM_BM == 0) kranges[m / GEMM_BM] = cell_kr[g];
  }
}

}  // namespace

int launch_build_maps(plsb_ctx *h, int n, int rows_pc, int stride_r, long long cellpad,
                      int *row_map, int4 *kranges, cudaStream_t st) {
  KernelTimer kt(h, KC_BUILD, st);
  const long long total = (long long)h->lay.J * cellpad;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 8);
  build_maps_kernel<<<blocks, 256, 0, st>>>(n, rows_pc, stride_r, h->lay.J, cellpad, h->d_cell_kr,
                                            row_map, kranges);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_build(plsb_ctx *h, int kind, const int32_t *idx, const double *yperm, int count,
                 double *A, double *Ac, double *distrib, long long cellpad_w, long long cellpad_c,
                 cudaStream_t st, int *ntrain, double *ytrain, const HalfSpec *half) {
  KernelTimer kt(h, KC_BUILD, st);
  const Layout &l = h->lay;
  if (count <= 0) return PLSB_OK;
  BuildParams p;
  p.idx = idx;
  p.S = l.S; p.T = l.T; p.J = l.J; p.K = l.K; p.L = l.L; p.lda = l.S_pad;
  p.corr = l.corr() ? 1 : 0;
  p.kind = kind;
  p.Y = h->Y.as<double>();
  p.Yperm = yperm;
  PLSB_CHECK(!yperm || (l.behavioral() && kind != BUILD_BOOT && kind != BUILD_TRAIN), PLSB_ERR_ARG,
             "pre-permuted Y matrices only apply to behavioural permutations");
  PLSB_CHECK(kind != BUILD_TRAIN || (l.behavioral() && idx && ntrain && ytrain), PLSB_ERR_ARG,
             "train / test operands need a behavioural analysis, masks and output buffers");
  PLSB_CHECK(kind != BUILD_HALF || (half && idx && half->ns >= 1 && (!l.behavioral() || ntrain)),
             PLSB_ERR_ARG, "split-half operands need masks and a split layout");
  p.yidx = half ? half->yidx : nullptr;
  p.half_ns = half ? half->ns : 1;
  p.half_s0 = half ? half->s0 : 0;
  p.half_nsplit = half ? half->n_split : 1;
  p.n_cond = l.n_cond;
  p.mean_centering = l.mean_centering;
  p.cell_start = h->d_cell_start;
  p.cell_of_row = h->d_cell_of_row;
  p.Vo = h->Vo.as<double>();
  p.Sx = h->Sx.as<double>();
  p.Cmat = h->Cmat.as<double>();
  p.A = A; p.Ac = Ac; p.distrib = distrib;
  p.ntrain = ntrain; p.ytrain = ytrain;
  p.cellpad_w = cellpad_w; p.cellpad_c = cellpad_c;
  if (kind == BUILD_ROT || (kind == BUILD_BOOT && distrib))
    PLSB_CHECK(h->has_original, PLSB_ERR_STATE, "operand builder needs the original decomposition");
  size_t smem;
  if (l.behavioral()) {
    smem = sizeof(double) * ((size_t)l.S * l.T + 2 * (size_t)l.J * l.T + 2 * (size_t)l.J * l.L) +
           sizeof(int) * (size_t)l.S;
    PLSB_CHECK(smem <= 200 * 1024, PLSB_ERR_ARG,
               "behavioural operand builder needs %zu bytes of shared memory (S*T too large)", smem);
    PLSB_CUDA(cudaFuncSetAttribute(build_behavioral_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_behavioral_kernel<<<count, 256, smem, st>>>(p);
  } else {
    smem = sizeof(double) * (size_t)l.J * l.S + sizeof(int) * 2 * (size_t)l.S;
    PLSB_CHECK(smem <= 200 * 1024, PLSB_ERR_ARG,
               "mean-centred operand builder needs %zu bytes of shared memory (J*S too large)", smem);
    PLSB_CUDA(cudaFuncSetAttribute(build_meancentered_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_meancentered_kernel<<<count, 256, smem, st>>>(p);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
