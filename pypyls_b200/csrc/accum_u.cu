// accum_u: the FP64 tensor-core (DMMA m8n8k4) streaming contraction that turns
// the stored cross-covariance matrices R (count blocks of K rows x ldr) and the
// per-resample rotations M into the bootstrap sums (pyls/base.py:510-511 with
// compute.procrustes, pyls/compute.py:260-262, folded into M):
//
//   U_r = R_r^T M_r  (B x L)  for every resample r of a split,
//   u_sum += U_r,  u_square += U_r^2   kept in registers.
//
// One CTA per (column tile of R, split of the resamples): 8 consumer warps and 2
// producer warps.  A consumer warp owns MA * 8 columns x all L latent variables
// (MA = 2 while the three accumulator sets fit the register file, else 1).  A
// stage holds `rs` resamples: the K x BT tile of R_r (bank-conflict-free pitch)
// and the rotation M_r, which is stored in global memory with the padded pitch
// the fragment loads want (accum_ldm) and is therefore one contiguous block.
// The producers fill a ring of slots with cp.async (16 B per thread and copy;
// measured: per-row TMA bulk copies of 0.5-1 KB are issue-bound here) and
// signal a "full" mbarrier through cp.async.mbarrier.arrive; consumers release
// slots through an "empty" mbarrier.  There is no CTA-wide barrier in the loop.
// Partial sums of the splits are reduced afterwards in a fixed order
// (stream_kernels.cu).
#include "common.cuh"

namespace plsb {

int launch_reduce_partials(plsb_ctx *h, const double *P, int n_splits, size_t stride, size_t n,
                           double *out, cudaStream_t st);

namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
// the mbarrier receives one arrival when all earlier cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

constexpr int AU_WARPS = 8;                        // consumer warps
constexpr int AU_PRODUCERS = 64;                   // producer threads (2 warps)
constexpr int AU_THREADS = AU_WARPS * 32 + AU_PRODUCERS;
constexpr int AU_MAX_SLOTS = 8;
constexpr size_t AU_SMEM_MAX = 226 * 1024;

// NFL = fragments of 8 latent variables, MA = fragments of 8 columns per warp
template <int NFL, int MA>
__global__ void __launch_bounds__(AU_THREADS, 1)
accum_u_kernel(const double *__restrict__ R, long long ldr, int count, int K, int B,
               const double *__restrict__ M, int L, int per_split, int rs, int nslot,
               double *__restrict__ Psum, double *__restrict__ Psq) {
  constexpr int BT = AU_WARPS * MA * 8;   // columns per CTA
  constexpr int LDR = BT + 4;             // == 4 (mod 16)
  constexpr int LDM = NFL * 8 + 4;        // == 4 or 12 (mod 16); also the pitch of M in HBM
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t full_bar[AU_MAX_SLOTS], empty_bar[AU_MAX_SLOTS];
  const int KP4 = (K + 3) & ~3;
  const int per_res = KP4 * (LDR + LDM);  // doubles per resample: R tile, then M
  const int stage = rs * per_res;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int b0 = blockIdx.x * BT, split = blockIdx.y;
  const int r_beg = split * per_split, r_end = min(count, r_beg + per_split);
  const int n_stages = (r_end - r_beg + rs - 1) / rs;

  for (int e = tid; e < nslot * stage; e += AU_THREADS) sm[e] = 0.0;   // padding rows stay zero
  if (tid == 0) {
    for (int s = 0; s < nslot; ++s) {
      mbar_init(&full_bar[s], AU_PRODUCERS);
      mbar_init(&empty_bar[s], AU_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= AU_WARPS) {
    // ---- producers: thread p copies 16-byte segment (p % SEGS) of rows p / SEGS, ... ----
    constexpr int SEGS = BT / 2, ROWS_PER_PASS = AU_PRODUCERS / SEGS;
    static_assert(AU_PRODUCERS % SEGS == 0, "producer threads vs. tile width");
    const int p = tid - AU_WARPS * 32;
    const int row0 = p / SEGS, seg = p - row0 * SEGS;
    const int m_segs = K * LDM / 2;
    for (int st = 0; st < n_stages; ++st) {
      const int slot = st % nslot;
      if (st >= nslot) mbar_wait(&empty_bar[slot], ((st / nslot) - 1) & 1);
      const int n_here = min(rs, r_end - (r_beg + st * rs));
      for (int t = 0; t < n_here; ++t) {
        const int r = r_beg + st * rs + t;
        double *Rs = sm + (size_t)slot * stage + t * per_res, *Ms = Rs + KP4 * LDR;
        const double *src = R + ((size_t)r * K + row0) * ldr + b0 + seg * 2;
        double *dst = Rs + row0 * LDR + seg * 2;
        for (int row = row0; row < K; row += ROWS_PER_PASS) {
          cp_async16(dst, src);
          src += (size_t)ROWS_PER_PASS * ldr;
          dst += ROWS_PER_PASS * LDR;
        }
        const double *Mr = M + (size_t)r * K * LDM;
        for (int e = p; e < m_segs; e += AU_PRODUCERS) cp_async16(Ms + e * 2, Mr + e * 2);
      }
      cp_async_arrive(&full_bar[slot]);
    }
    return;
  }

  double us[MA][NFL][2], uq[MA][NFL][2];
#pragma unroll
  for (int i = 0; i < MA; ++i)
#pragma unroll
    for (int j = 0; j < NFL; ++j) us[i][j][0] = us[i][j][1] = uq[i][j][0] = uq[i][j][1] = 0.0;

  for (int st = 0; st < n_stages; ++st) {
    const int slot = st % nslot;
    mbar_wait(&full_bar[slot], (st / nslot) & 1);
    const int n_here = min(rs, r_end - (r_beg + st * rs));
    for (int t = 0; t < n_here; ++t) {
      const double *Rs = sm + (size_t)slot * stage + t * per_res, *Ms = Rs + KP4 * LDR;
      const double *ap = Rs + q * LDR + warp * (MA * 8) + g;
      const double *bp = Ms + q * LDM + g;
      double acc[MA][NFL][2];
#pragma unroll
      for (int i = 0; i < MA; ++i)
#pragma unroll
        for (int j = 0; j < NFL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 2
      for (int k0 = 0; k0 < KP4; k0 += 4) {
        double a[MA], b[NFL];
#pragma unroll
        for (int i = 0; i < MA; ++i) a[i] = ap[k0 * LDR + i * 8];
#pragma unroll
        for (int j = 0; j < NFL; ++j) b[j] = bp[k0 * LDM + j * 8];
#pragma unroll
        for (int i = 0; i < MA; ++i)
#pragma unroll
          for (int j = 0; j < NFL; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
#pragma unroll
      for (int i = 0; i < MA; ++i)
#pragma unroll
        for (int j = 0; j < NFL; ++j) {
          us[i][j][0] += acc[i][j][0];
          us[i][j][1] += acc[i][j][1];
          uq[i][j][0] += acc[i][j][0] * acc[i][j][0];
          uq[i][j][1] += acc[i][j][1] * acc[i][j][1];
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);
  }

#pragma unroll
  for (int i = 0; i < MA; ++i) {
    const int b = b0 + warp * (MA * 8) + i * 8 + g;
    if (b < B) {
      const size_t base = ((size_t)split * B + b) * L;
#pragma unroll
      for (int j = 0; j < NFL; ++j) {
        const int l = j * 8 + 2 * q;
        if (l < L) {
          Psum[base + l] = us[i][j][0];
          Psq[base + l] = uq[i][j][0];
        }
        if (l + 1 < L) {
          Psum[base + l + 1] = us[i][j][1];
          Psq[base + l + 1] = uq[i][j][1];
        }
      }
    }
  }
}

// splits of the resamples so that the grid fills whole waves of one CTA per SM
int pick_splits(int sm_count, int n_bt, int count) {
  int best = 1;
  double best_eff = 0.0;
  const int lo = std::max(1, cdiv(2 * sm_count, n_bt)), hi = std::max(lo, cdiv(10 * sm_count, n_bt));
  for (int s = lo; s <= std::min(hi, count); ++s) {
    const int per = cdiv(count, s), used = cdiv(count, per);
    const long long ctas = (long long)used * n_bt;
    // every CTA of a wave runs `per` resamples; idle slots of the last wave are the loss
    const long long waves = (ctas + sm_count - 1) / sm_count;
    const double eff = (double)count * n_bt / ((double)waves * sm_count * per);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = used;
    }
  }
  return std::min(best, std::max(count, 1));
}

template <int NFL, int MA>
int launch_au(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
              const double *M, int L, double *usum, double *usq, cudaStream_t st) {
  constexpr int BT = AU_WARPS * MA * 8, LDR = BT + 4, LDM = NFL * 8 + 4;
  PLSB_CHECK(ldr % BT == 0 && ldr >= B, PLSB_ERR_ARG,
             "accum_u: row pitch %lld not a multiple of %d", ldr, BT);
  const int KP4 = round_up(K, 4);
  const size_t per_res = sizeof(double) * (size_t)KP4 * (LDR + LDM);
  // resamples per stage: ~48 KB stages
  int rs = (int)std::max<size_t>(1, std::min<size_t>((48 * 1024) / per_res, 8));
  rs = std::max(1, tune_int("PLSB_AU_RS", rs));
  const int n_bt = cdiv(B, BT);
  int n_splits = tune_int("PLSB_AU_SPLITS", pick_splits(h->sm_count, n_bt, count));
  const int per_split = cdiv(count, n_splits);
  n_splits = cdiv(count, per_split);
  rs = std::min(rs, per_split);
  const size_t stage = per_res * rs;
  int nslot = (int)std::min<size_t>(AU_MAX_SLOTS, AU_SMEM_MAX / stage);
  nslot = std::max(2, std::min(nslot, cdiv(per_split, rs) + 1));
  nslot = tune_int("PLSB_AU_SLOTS", nslot);
  const size_t smem = stage * nslot;
  PLSB_CHECK(smem <= AU_SMEM_MAX, PLSB_ERR_ARG, "accum_u: %zu bytes of shared memory", smem);
  const size_t stride = (size_t)B * L;
  PLSB_TRY(h->part.ensure(sizeof(double) * 2 * stride * n_splits));
  double *Psum = h->part.as<double>(), *Psq = Psum + stride * n_splits;
  auto kern = accum_u_kernel<NFL, MA>;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_bt, n_splits);
  kern<<<grid, AU_THREADS, smem, st>>>(R, ldr, count, K, B, M, L, per_split, rs, nslot, Psum, Psq);
  PLSB_LAUNCHED(h);
  PLSB_TRY(launch_reduce_partials(h, Psum, n_splits, stride, stride, usum, st));
  PLSB_TRY(launch_reduce_partials(h, Psq, n_splits, stride, stride, usq, st));
  return PLSB_OK;
}

}  // namespace

// R: (count*K rows, ldr) with ldr a multiple of 128 and zero columns >= B;
// M: (count, K, accum_ldm(L)) -- rows padded to the pitch the kernel's fragment
// loads use, padding columns zero.
int launch_accum_u(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                   const double *M, int L, double *usum, double *usq, cudaStream_t st) {
  KernelTimer kt(h, KC_ACCUM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(L >= 1 && K >= 1, PLSB_ERR_ARG, "accum_u: K=%d L=%d", K, L);
  if (small_k_applies(K, L, true))
    return launch_accum_u_small(h, R, ldr, count, K, B, M, L, usum, usq, st);
  if (K > MAX_K || L > MAX_K)
    return launch_accum_u_generic(h, R, ldr, count, K, B, M, accum_ldm(L), L, usum, usq, st);
  switch (cdiv(L, 8)) {
    case 1: return launch_au<1, 2>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 2: return launch_au<2, 2>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 3: return launch_au<3, 2>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 4: return launch_au<4, 2>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 5: return launch_au<5, 2>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 6: return launch_au<6, 1>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 7: return launch_au<7, 1>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 8: return launch_au<8, 1>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    case 9: return launch_au<9, 1>(h, R, ldr, count, K, B, M, L, usum, usq, st);
    default: return launch_au<10, 1>(h, R, ldr, count, K, B, M, L, usum, usq, st);
  }
}

}  // namespace plsb
