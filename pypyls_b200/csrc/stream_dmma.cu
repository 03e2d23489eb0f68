// FP64 tensor-core (DMMA m8n8k4) versions of the two streaming contractions
// over the stored cross-covariance matrices R (count blocks of K rows x ldr).
//
//   gram_proj   [G | H] = R [R^T | U_orig]     (K x (K+L), contraction over B)
//               one CTA per resample; R and U_orig tiles of 64 columns are
//               staged with cp.async in a two-slot ring; the 16 warps split
//               the output columns (2 fragments each) x the contraction steps
//               and are reduced through shared memory at the end
//   accum_u     U_r = R_r^T M_r (B x L) for every resample r of a split,
//               u_sum += U_r, u_square += U_r^2 kept in registers; one CTA per
//               (64 columns of R, split of the resamples), two-slot cp.async
//               ring over the resamples, partial sums reduced afterwards
//
// Shared-memory leading dimensions are == 4 or 12 (mod 16) doubles so that
// every fragment load (8 x 4 doubles) is bank-conflict free.
#include "common.cuh"

namespace plsb {

int launch_reduce_partials(plsb_ctx *h, const double *P, int n_splits, size_t stride, size_t n,
                           double *out, cudaStream_t st);

namespace {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// 8-byte copy; src_bytes == 0 writes zeros without reading
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ---- gram_proj ----------------------------------------------------------------
constexpr int GD_THREADS = 512;
constexpr int GD_WARPS = GD_THREADS / 32;

// GD_BC = columns of R per stage (64 or 128)
template <int MF, int GD_BC>
__global__ void __launch_bounds__(GD_THREADS, 1)
gram_proj_dmma_kernel(const double *__restrict__ R, long long ldr, int K, int B, int n_chunks,
                      const double *__restrict__ Uo, int L, int LP, int ldu,
                      double *__restrict__ G, double *__restrict__ H, int NG, int KG) {
  extern __shared__ __align__(16) double sm[];
  constexpr int KP = MF * 8;
  constexpr int GD_LDR = GD_BC + 4;              // == 4 (mod 16)
  const int stage = KP * GD_LDR + GD_BC * ldu;   // doubles per ring slot
  const int NF = (KP + LP) / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int r = blockIdx.x;
  const double *Rr = R + (size_t)r * K * ldr;

  // zero the padding that the copies never touch: rows >= K of Rs, columns >= L of Us
  for (int s = 0; s < 2; ++s) {
    double *Rs = sm + s * stage, *Us = Rs + KP * GD_LDR;
    for (int e = tid; e < (KP - K) * GD_LDR; e += GD_THREADS) Rs[K * GD_LDR + e] = 0.0;
    if (LP > L)
      for (int e = tid; e < GD_BC * (LP - L); e += GD_THREADS) {
        const int b = e / (LP - L), l = L + e - b * (LP - L);
        Us[b * ldu + l] = 0.0;
      }
  }

  auto load = [&](int ch, int slot) {
    double *Rs = sm + slot * stage, *Us = Rs + KP * GD_LDR;
    const int b0 = ch * GD_BC;
    for (int e = tid; e < K * (GD_BC / 2); e += GD_THREADS) {
      const int c = e / (GD_BC / 2), seg = e - c * (GD_BC / 2);
      cp_async16(Rs + c * GD_LDR + seg * 2, Rr + (size_t)c * ldr + b0 + seg * 2);
    }
    if (LP > 0)
      for (int e = tid; e < GD_BC * L; e += GD_THREADS) {
        const int b = e / L, l = e - b * L;
        const bool ok = b0 + b < B;
        cp_async8(Us + b * ldu + l, ok ? Uo + (size_t)(b0 + b) * L + l : Uo, ok ? 8 : 0);
      }
    cp_async_commit();
  };

  const int ng = warp % NG, kg = warp / NG;
  const bool active = kg < KG;
  double acc[MF][2][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  load(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    if (ch + 1 < n_chunks) {
      load(ch + 1, (ch + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (active) {
      const double *Rs = sm + (ch & 1) * stage, *Us = Rs + KP * GD_LDR;
      for (int kk = kg; kk < GD_BC / 4; kk += KG) {
        double a[MF];
#pragma unroll
        for (int i = 0; i < MF; ++i) a[i] = Rs[(i * 8 + g) * GD_LDR + kk * 4 + q];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nf = ng * 2 + j;
          if (nf < NF) {
            const int col0 = nf * 8;
            const double bv = col0 < KP ? Rs[(col0 + g) * GD_LDR + kk * 4 + q]
                                        : Us[(kk * 4 + q) * ldu + (col0 - KP) + g];
#pragma unroll
            for (int i = 0; i < MF; ++i) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], bv);
          }
        }
      }
    }
    __syncthreads();
  }

  // reduce the contraction groups through shared memory: red[kg][row][col]
  const int NCP = NF * 8;
  double *red = sm;
  if (active) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int nf = ng * 2 + j;
      if (nf < NF) {
#pragma unroll
        for (int i = 0; i < MF; ++i) {
          double *p = red + ((size_t)kg * KP + i * 8 + g) * NCP + nf * 8 + 2 * q;
          p[0] = acc[i][j][0];
          p[1] = acc[i][j][1];
        }
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < K * NCP; e += GD_THREADS) {
    const int row = e / NCP, col = e - row * NCP;
    double v = 0.0;
    for (int s = 0; s < KG; ++s) v += red[((size_t)s * KP + row) * NCP + col];
    if (col < K)
      G[((size_t)r * K + row) * K + col] = v;
    else if (col >= KP && col - KP < L)
      H[((size_t)r * K + row) * L + (col - KP)] = v;
  }
}

// ---- accum_u ------------------------------------------------------------------
constexpr int AD_BT = 64;            // columns of R per CTA (8 per warp)
constexpr int AD_THREADS = 256;
constexpr int AD_LDR = AD_BT + 4;    // 68

template <int NFL>
__global__ void __launch_bounds__(AD_THREADS)
accum_u_dmma_kernel(const double *__restrict__ R, long long ldr, int count, int K, int KP4, int B,
                    const double *__restrict__ M, int L, int ldm, int per_split,
                    double *__restrict__ Psum, double *__restrict__ Psq) {
  extern __shared__ __align__(16) double sm[];
  const int stage = KP4 * AD_LDR + KP4 * ldm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * AD_BT, split = blockIdx.y;
  const int r_beg = split * per_split, r_end = min(count, r_beg + per_split);
  const int LP = NFL * 8;

  for (int s = 0; s < 2; ++s) {
    double *Rs = sm + s * stage, *Ms = Rs + KP4 * AD_LDR;
    for (int e = tid; e < (KP4 - K) * AD_LDR; e += AD_THREADS) Rs[K * AD_LDR + e] = 0.0;
    for (int e = tid; e < KP4 * ldm; e += AD_THREADS) {
      const int k = e / ldm, l = e - k * ldm;
      if (k >= K || l >= L) Ms[e] = 0.0;
    }
  }
  __syncthreads();

  auto load = [&](int r, int slot) {
    double *Rs = sm + slot * stage, *Ms = Rs + KP4 * AD_LDR;
    const double *Rr = R + (size_t)r * K * ldr + b0;
    const double *Mr = M + (size_t)r * K * L;
    for (int e = tid; e < K * (AD_BT / 2); e += AD_THREADS) {
      const int k = e / (AD_BT / 2), seg = e - k * (AD_BT / 2);
      cp_async16(Rs + k * AD_LDR + seg * 2, Rr + (size_t)k * ldr + seg * 2);
    }
    for (int e = tid; e < K * L; e += AD_THREADS) {
      const int k = e / L, l = e - k * L;
      cp_async8(Ms + k * ldm + l, Mr + e, 8);
    }
    cp_async_commit();
  };

  double us[NFL][2], uq[NFL][2];
#pragma unroll
  for (int j = 0; j < NFL; ++j) us[j][0] = us[j][1] = uq[j][0] = uq[j][1] = 0.0;

  if (r_beg < r_end) load(r_beg, 0);
  for (int r = r_beg; r < r_end; ++r) {
    const int slot = (r - r_beg) & 1;
    if (r + 1 < r_end) {
      load(r + 1, slot ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double *Rs = sm + slot * stage, *Ms = Rs + KP4 * AD_LDR;
    double acc[NFL][2];
#pragma unroll
    for (int j = 0; j < NFL; ++j) acc[j][0] = acc[j][1] = 0.0;
    for (int kk = 0; kk < KP4 / 4; ++kk) {
      const double a = Rs[(kk * 4 + q) * AD_LDR + warp * 8 + g];
#pragma unroll
      for (int j = 0; j < NFL; ++j) {
        const double bv = Ms[(kk * 4 + q) * ldm + j * 8 + g];
        dmma_8x8x4(acc[j][0], acc[j][1], a, bv);
      }
    }
#pragma unroll
    for (int j = 0; j < NFL; ++j) {
      us[j][0] += acc[j][0];
      us[j][1] += acc[j][1];
      uq[j][0] += acc[j][0] * acc[j][0];
      uq[j][1] += acc[j][1] * acc[j][1];
    }
    __syncthreads();
  }

  const int b = b0 + warp * 8 + g;
  if (b < B) {
    const size_t base = ((size_t)split * B + b) * L;
#pragma unroll
    for (int j = 0; j < NFL; ++j) {
      const int l = j * 8 + 2 * q;
      if (l < L) {
        Psum[base + l] = us[j][0];
        Psq[base + l] = uq[j][0];
      }
      if (l + 1 < L) {
        Psum[base + l + 1] = us[j][1];
        Psq[base + l + 1] = uq[j][1];
      }
    }
  }
  (void)LP;
}

template <int MF, int BC>
int launch_gp_bc(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                 const double *Uo, int L, int LP, double *G, double *H, size_t smem, int NG, int KG,
                 cudaStream_t st) {
  PLSB_CUDA(cudaFuncSetAttribute(gram_proj_dmma_kernel<MF, BC>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_chunks = (int)(ldr / BC);
  gram_proj_dmma_kernel<MF, BC><<<count, GD_THREADS, smem, st>>>(R, ldr, K, B, n_chunks, Uo, L, LP,
                                                                 LP + 4, G, H, NG, KG);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

template <int MF>
int launch_gp(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
              const double *Uo, int L, int LP, double *G, double *H, cudaStream_t st) {
  constexpr int KP = MF * 8;
  const int ldu = LP + 4;
  const int NF = (KP + LP) / 8;
  const int NG = (NF + 1) / 2;
  PLSB_CHECK(NG <= GD_WARPS, PLSB_ERR_ARG, "gram_proj: too many output fragments");
  auto smem_for = [&](int bc, int kg) {
    const size_t stage = (size_t)KP * (bc + 4) + (size_t)bc * ldu;
    const size_t red = (size_t)kg * KP * NF * 8;
    return sizeof(double) * std::max(2 * stage, red);
  };
  // 128-column stages (fewer barriers, better balance of the contraction
  // groups) when two of them fit, else 64
  const int KG128 = std::max(1, std::min(GD_WARPS / NG, 128 / 4));
  if (ldr % 128 == 0 && smem_for(128, KG128) <= 200 * 1024)
    return launch_gp_bc<MF, 128>(h, R, ldr, count, K, B, Uo, L, LP, G, H, smem_for(128, KG128), NG,
                                 KG128, st);
  const int KG64 = std::max(1, std::min(GD_WARPS / NG, 64 / 4));
  return launch_gp_bc<MF, 64>(h, R, ldr, count, K, B, Uo, L, LP, G, H, smem_for(64, KG64), NG, KG64,
                              st);
}

template <int NFL>
int launch_au(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
              const double *M, int L, int n_bt, int n_splits, int per_split, double *Psum,
              double *Psq, cudaStream_t st) {
  const int KP4 = round_up(K, 4);
  const int ldm = NFL * 8 + 4;
  const size_t smem = sizeof(double) * 2 * ((size_t)KP4 * AD_LDR + (size_t)KP4 * ldm);
  PLSB_CUDA(cudaFuncSetAttribute(accum_u_dmma_kernel<NFL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_bt, n_splits);
  accum_u_dmma_kernel<NFL><<<grid, AD_THREADS, smem, st>>>(R, ldr, count, K, KP4, B, M, L, ldm,
                                                           per_split, Psum, Psq);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

}  // namespace

// R must be a (count*K rows, ldr) buffer whose row pitch ldr is a multiple of 64
// and whose columns >= B are zero (the GEMM's padded output).
int launch_gram_proj(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                     const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  KernelTimer kt(h, KC_GRAM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= MAX_K, PLSB_ERR_ARG, "gram_proj: K=%d outside [1,%d]", K, MAX_K);
  PLSB_CHECK(ldr % 64 == 0, PLSB_ERR_ARG, "gram_proj: row pitch %lld not a multiple of 64", ldr);
  const bool proj = Uo && H;
  PLSB_CHECK(!proj || (L >= 1 && L <= MAX_K), PLSB_ERR_ARG, "gram_proj: L=%d", L);
  const int LP = proj ? round_up(L, 8) : 0;
  const int Lk = proj ? L : 0;
  switch (cdiv(K, 8)) {
    case 1: return launch_gp<1>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 2: return launch_gp<2>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 3: return launch_gp<3>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 4: return launch_gp<4>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 5: return launch_gp<5>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 6: return launch_gp<6>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 7: return launch_gp<7>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 8: return launch_gp<8>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    case 9: return launch_gp<9>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
    default: return launch_gp<10>(h, R, ldr, count, K, B, Uo, Lk, LP, G, H, st);
  }
}

int launch_accum_u(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                   const double *M, int L, double *usum, double *usq, cudaStream_t st) {
  KernelTimer kt(h, KC_ACCUM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(L >= 1 && L <= MAX_K && K >= 1 && K <= MAX_K, PLSB_ERR_ARG,
             "accum_u: K=%d L=%d outside [1,%d]", K, L, MAX_K);
  PLSB_CHECK(ldr % AD_BT == 0 && ldr >= B, PLSB_ERR_ARG,
             "accum_u: row pitch %lld not a multiple of %d", ldr, AD_BT);
  const int n_bt = cdiv(B, AD_BT);
  int n_splits = std::max(1, std::min({cdiv(6 * h->sm_count, n_bt), count, 256}));
  const int per_split = cdiv(count, n_splits);
  n_splits = cdiv(count, per_split);
  const size_t stride = (size_t)B * L;
  PLSB_TRY(h->part.ensure(sizeof(double) * 2 * stride * n_splits));
  double *Psum = h->part.as<double>(), *Psq = Psum + stride * n_splits;
  int rc;
  switch (cdiv(L, 8)) {
    case 1: rc = launch_au<1>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 2: rc = launch_au<2>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 3: rc = launch_au<3>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 4: rc = launch_au<4>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 5: rc = launch_au<5>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 6: rc = launch_au<6>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 7: rc = launch_au<7>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 8: rc = launch_au<8>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    case 9: rc = launch_au<9>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
    default: rc = launch_au<10>(h, R, ldr, count, K, B, M, L, n_bt, n_splits, per_split, Psum, Psq, st); break;
  }
  PLSB_TRY(rc);
  PLSB_TRY(launch_reduce_partials(h, Psum, n_splits, stride, stride, usum, st));
  PLSB_TRY(launch_reduce_partials(h, Psq, n_splits, stride, stride, usq, st));
  return PLSB_OK;
}

}  // namespace plsb
