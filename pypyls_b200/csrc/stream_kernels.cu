// Streaming kernels over the stored cross-covariance matrices R (count*K rows
// of ldr doubles, one K x B block per resample).
//
//   finish_rowsq  rotated permutation singular values from the GEMM's row sums
//                 of squares (pyls/base.py:699-700)
//   colscale      1 / ((n-1) * sigma) of the resampled X columns from the count
//                 operand's sums and sums of squares (compute.xcorr's z-score of
//                 X[inds], pyls/compute.py:84, restated on multiplicities)
//   gram_proj     G = R R^T (K,K) and H = R U_orig (K,L) per resample: all the
//                 small SVD + Procrustes step needs (pyls/compute.py:36-49, 260)
//   accum_u       u_sum += R^T M, u_square += (R^T M)^2 summed over resamples
//                 (pyls/base.py:510-511 with compute.procrustes folded into M)
#include "common.cuh"

namespace plsb {
namespace {

__global__ void finish_rowsq_kernel(const double *__restrict__ rowsq, int n_splits, int M_pad,
                                    int n_rows, double *__restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_rows) return;
  double v = 0.0;
  for (int s = 0; s < n_splits; ++s) v += rowsq[(size_t)s * M_pad + m];
  out[m] = sqrt(v);
}

__global__ void colscale_kernel(double *__restrict__ S1, const double *__restrict__ S2, int n_rows,
                                long long ld, int B, int J, const int *__restrict__ cell_n) {
  const size_t total = (size_t)n_rows * ld;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / ld), b = (int)(e % ld);
    double out = 0.0;
    if (b < B) {
      const double n = (double)cell_n[row % J];
      const double s1 = S1[e], s2 = S2[e];
      const double var = (s2 - s1 * s1 / n) / (n - 1.0);
      out = 1.0 / ((n - 1.0) * sqrt(var));
    }
    S1[e] = out;
  }
}

// ---- gram_proj --------------------------------------------------------------
constexpr int GP_BC = 32;       // columns of R per staged chunk
constexpr int GP_THREADS = 256;
constexpr int GP_MAXT = 4;      // 4x4 output tiles per thread (K <= 80)

__global__ void __launch_bounds__(GP_THREADS)
gram_proj_kernel(const double *__restrict__ R, long long ldr, int K, int B,
                 const double *__restrict__ Uo, int L, double *__restrict__ G,
                 double *__restrict__ H, int NC, int NCp, int tiles_c, int n_tiles, int slices) {
  extern __shared__ __align__(16) double sm[];
  double *Cs = sm;  // [GP_BC][NCp]; columns: K rows of R, then L columns of Uo
  const int r = blockIdx.x, tid = threadIdx.x;
  const double *Rr = R + (size_t)r * K * ldr;

  int my_tile[GP_MAXT];
  int my_slice = 0, n_my = 0;
  if (slices == 1) {
    for (int t = tid; t < n_tiles && n_my < GP_MAXT; t += GP_THREADS) my_tile[n_my++] = t;
  } else {
    const int sl = tid / n_tiles;
    if (sl < slices) {
      my_tile[0] = tid - sl * n_tiles;
      my_slice = sl;
      n_my = 1;
    }
  }
  double acc[GP_MAXT][4][4];
#pragma unroll
  for (int a = 0; a < GP_MAXT; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.0;

  for (int b0 = 0; b0 < B; b0 += GP_BC) {
    __syncthreads();
    for (int e = tid; e < NCp * GP_BC; e += GP_THREADS) {
      double v = 0.0;
      int c, b;
      if (e < K * GP_BC) {          // R part: consecutive threads walk b (coalesced)
        c = e / GP_BC;
        b = e - c * GP_BC;
        if (b0 + b < B) v = Rr[(size_t)c * ldr + b0 + b];
      } else {                      // Uo part (and zero padding up to NCp)
        const int e2 = e - K * GP_BC;
        const int w = NCp - K;
        b = e2 / w;
        c = K + (e2 - b * w);
        if (c < NC && b0 + b < B) v = Uo[(size_t)(b0 + b) * L + (c - K)];
      }
      Cs[b * NCp + c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < GP_MAXT; ++a) {
      if (a < n_my) {
        const int ti = my_tile[a] / tiles_c, tc = my_tile[a] - ti * tiles_c;
        const double *pa = Cs + ti * 4, *pb = Cs + tc * 4;
        for (int b = my_slice; b < GP_BC; b += slices) {
          const double2 a01 = *reinterpret_cast<const double2 *>(pa + b * NCp);
          const double2 a23 = *reinterpret_cast<const double2 *>(pa + b * NCp + 2);
          const double2 b01 = *reinterpret_cast<const double2 *>(pb + b * NCp);
          const double2 b23 = *reinterpret_cast<const double2 *>(pb + b * NCp + 2);
          const double av[4] = {a01.x, a01.y, a23.x, a23.y};
          const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[a][i][j] += av[i] * bv[j];
        }
      }
    }
  }

  // cross-slice reduction (slices > 1 implies one tile per thread)
  if (slices > 1) {
    __syncthreads();
    double *red = sm;  // [slices][n_tiles][16]
    if (n_my) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          red[((size_t)my_slice * n_tiles + my_tile[0]) * 16 + i * 4 + j] = acc[0][i][j];
    }
    __syncthreads();
    if (n_my && my_slice == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double v = 0.0;
          for (int s = 0; s < slices; ++s)
            v += red[((size_t)s * n_tiles + my_tile[0]) * 16 + i * 4 + j];
          acc[0][i][j] = v;
        }
    } else {
      n_my = 0;
    }
  }
#pragma unroll
  for (int a = 0; a < GP_MAXT; ++a) {
    if (a < n_my) {
      const int ti = my_tile[a] / tiles_c, tc = my_tile[a] - ti * tiles_c;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = ti * 4 + i;
        if (row >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tc * 4 + j;
          if (c < K)
            G[((size_t)r * K + row) * K + c] = acc[a][i][j];
          else if (c < NC)
            H[((size_t)r * K + row) * L + (c - K)] = acc[a][i][j];
        }
      }
    }
  }
}

// ---- accum_u ----------------------------------------------------------------
constexpr int AU_BT = 64;  // columns of R per CTA
constexpr int AU_THREADS = 256;

template <int NLMAX>
__global__ void __launch_bounds__(AU_THREADS)
accum_u_kernel(const double *__restrict__ R, long long ldr, int count, int K, int B,
               const double *__restrict__ M, int L, int per_split, double *__restrict__ Psum,
               double *__restrict__ Psq) {
  extern __shared__ __align__(16) double sm[];
  double *Rs = sm;                // [K][AU_BT]
  double *Ms = sm + K * AU_BT;    // [K][L]
  const int tid = threadIdx.x, bl = tid & (AU_BT - 1), lg = tid >> 6;
  const int b0 = blockIdx.x * AU_BT, split = blockIdx.y;
  const int r_beg = split * per_split, r_end = min(count, r_beg + per_split);
  const int nl = (L - lg + 3) / 4;  // this thread's columns l = lg + 4 i, i < nl

  double us[NLMAX], uq[NLMAX];
#pragma unroll
  for (int i = 0; i < NLMAX; ++i) us[i] = uq[i] = 0.0;

  for (int r = r_beg; r < r_end; ++r) {
    const double *Rr = R + (size_t)r * K * ldr;
    const double *Mr = M + (size_t)r * K * L;
    __syncthreads();
    for (int e = tid; e < K * AU_BT; e += AU_THREADS) {
      const int k = e / AU_BT, b = e - k * AU_BT;
      Rs[e] = (b0 + b < B) ? Rr[(size_t)k * ldr + b0 + b] : 0.0;
    }
    for (int e = tid; e < K * L; e += AU_THREADS) Ms[e] = Mr[e];
    __syncthreads();
    double u[NLMAX];
#pragma unroll
    for (int i = 0; i < NLMAX; ++i) u[i] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double x = Rs[k * AU_BT + bl];
      const double *mrow = Ms + k * L + lg;
#pragma unroll
      for (int i = 0; i < NLMAX; ++i)
        if (i < nl) u[i] += x * mrow[4 * i];
    }
#pragma unroll
    for (int i = 0; i < NLMAX; ++i) {
      us[i] += u[i];
      uq[i] += u[i] * u[i];
    }
  }
  if (b0 + bl < B) {
    const size_t base = ((size_t)split * B + b0 + bl) * L;
#pragma unroll
    for (int i = 0; i < NLMAX; ++i)
      if (i < nl) {
        Psum[base + lg + 4 * i] = us[i];
        Psq[base + lg + 4 * i] = uq[i];
      }
  }
}

__global__ void reduce_partials_kernel(const double *__restrict__ P, int n_splits, size_t stride,
                                       size_t n, double *__restrict__ out) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < n_splits; ++s) v += P[(size_t)s * stride + e];
    out[e] += v;
  }
}

}  // namespace

int launch_reduce_partials(plsb_ctx *h, const double *P, int n_splits, size_t stride, size_t n,
                           double *out, cudaStream_t st) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 16);
  reduce_partials_kernel<<<blocks, 256, 0, st>>>(P, n_splits, stride, n, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_finish_rowsq(plsb_ctx *h, const double *rowsq, int n_splits, int M_pad, int n_rows,
                        double *out, cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_rows <= 0) return PLSB_OK;
  finish_rowsq_kernel<<<cdiv(n_rows, 256), 256, 0, st>>>(rowsq, n_splits, M_pad, n_rows, out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_colscale(plsb_ctx *h, double *S1, const double *S2, int n_rows, long long ld,
                    cudaStream_t st) {
  KernelTimer kt(h, KC_STATS, st);
  if (n_rows <= 0) return PLSB_OK;
  const size_t total = (size_t)n_rows * ld;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->sm_count * 32);
  colscale_kernel<<<blocks, 256, 0, st>>>(S1, S2, n_rows, ld, h->lay.B, h->lay.J, h->d_cell_n);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_gram_proj_fma(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                     const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  KernelTimer kt(h, KC_GRAM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= MAX_K, PLSB_ERR_ARG, "gram_proj: K=%d outside [1,%d]", K, MAX_K);
  const int NC = K + ((Uo && H) ? L : 0);
  const int NCp = round_up(NC, 4);
  const int tiles_i = cdiv(K, 4), tiles_c = NCp / 4;
  const int n_tiles = tiles_i * tiles_c;
  PLSB_CHECK(n_tiles <= GP_MAXT * GP_THREADS, PLSB_ERR_ARG, "gram_proj: too many output tiles");
  int slices = 1;
  if (n_tiles < GP_THREADS) slices = std::min(GP_THREADS / n_tiles, GP_BC);
  size_t smem = sizeof(double) * (size_t)GP_BC * NCp;
  if (slices > 1) smem = std::max(smem, sizeof(double) * (size_t)slices * n_tiles * 16);
  PLSB_CUDA(cudaFuncSetAttribute(gram_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  gram_proj_kernel<<<count, GP_THREADS, smem, st>>>(R, ldr, K, B, Uo, L, G, H, NC, NCp, tiles_c,
                                                    n_tiles, slices);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int launch_accum_u_fma(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                   const double *M, int L, double *usum, double *usq, cudaStream_t st) {
  KernelTimer kt(h, KC_ACCUM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(L >= 1 && L <= MAX_K && K <= MAX_K, PLSB_ERR_ARG, "accum_u: K=%d L=%d too large", K, L);
  const int n_bt = cdiv(B, AU_BT);
  int n_splits = std::max(1, std::min({cdiv(4 * h->sm_count, n_bt), count, 64}));
  const int per_split = cdiv(count, n_splits);
  n_splits = cdiv(count, per_split);
  const size_t stride = (size_t)B * L;
  PLSB_TRY(h->part.ensure(sizeof(double) * 2 * stride * n_splits));
  double *Psum = h->part.as<double>(), *Psq = Psum + stride * n_splits;
  const size_t smem = sizeof(double) * ((size_t)K * AU_BT + (size_t)K * L);
  dim3 grid(n_bt, n_splits);
#define PLSB_AU(NL)                                                                              \
  do {                                                                                           \
    PLSB_CUDA(cudaFuncSetAttribute(accum_u_kernel<NL>,                                           \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    accum_u_kernel<NL><<<grid, AU_THREADS, smem, st>>>(R, ldr, count, K, B, M, L, per_split,     \
                                                       Psum, Psq);                               \
  } while (0)
  const int nl = cdiv(L, 4);
  if (nl <= 4) PLSB_AU(4);
  else if (nl <= 8) PLSB_AU(8);
  else if (nl <= 12) PLSB_AU(12);
  else if (nl <= 16) PLSB_AU(16);
  else PLSB_AU(20);
#undef PLSB_AU
  PLSB_LAUNCHED(h);
  const int blocks = (int)std::min<size_t>((stride + 255) / 256, (size_t)h->sm_count * 16);
  reduce_partials_kernel<<<blocks, 256, 0, st>>>(Psum, n_splits, stride, stride, usum);
  PLSB_LAUNCHED(h);
  reduce_partials_kernel<<<blocks, 256, 0, st>>>(Psq, n_splits, stride, stride, usq);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace plsb
