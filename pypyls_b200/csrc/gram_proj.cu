// gram_proj:  [G | H] = R [R^T | U_orig]   (K x (K + L), contraction over B)
//
// FP64 tensor-core (DMMA m8n8k4) streaming contraction over one stored
// cross-covariance matrix R (K rows x ldr) per CTA -- all the small SVD +
// Procrustes step of a bootstrap needs (pyls/compute.py:36-49, 260: the
// randomized SVD of R and the U_orig^T U_boot product collapse onto G and H).
//
// Work decomposition (16 consumer warps + 2 producer warps):
//   * G is symmetric: only the MF (MF + 1) / 2 upper fragments (8 x 8) are
//     computed; with the MF * MFH fragments of H that is FT fragments, dealt
//     in contiguous runs to NG warp groups (compile-time tables, so every
//     accumulator stays in a register);
//   * inside a group the KGW = 16 / NG warps split the contraction steps of a
//     stage.  Warp w has (group w / KGW, slice w % KGW): KGW is a multiple of
//     4 (except for K > 56), so the four warps that share an SM sub-partition's
//     DMMA pipe are the same slice of each group and the sub-partitions are
//     balanced whatever the group sizes are;
//   * U_orig is kept transposed (L rows of ldr doubles), so a stage is one tile
//     of K + L rows x BC columns of [R; U_orig^T] and H = R (U_orig^T)^T has the
//     same fragment addressing as G = R R^T.  The producer warps stage the
//     tiles with cp.async (16 B per thread and copy, rows at a bank-conflict-free
//     pitch) through a ring of `nslot` slots and signals a "full" mbarrier per
//     slot through cp.async.mbarrier.arrive; the 16 consumer warps release slots
//     through an "empty" mbarrier.  No CTA-wide barrier in the loop: a warp only
//     waits for the stage it reads.  (Measured: per-row TMA bulk copies of
//     0.5 KB are issue-bound here, 72 % of the DMMA peak instead of ~80 %.);
//   * the slices are reduced pairwise through shared memory at the end (fixed
//     order: results are bit-reproducible).
//
// The shared-memory row pitch is == 4 (mod 16) doubles so that every fragment
// load (8 rows x 4 doubles) is bank-conflict free.
#include "common.cuh"

#include <type_traits>

namespace plsb {
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
// barrier among the consumer warps only (the producers have left by then)
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, 512;\n" ::: "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int F, int F1, typename Fn>
__device__ __forceinline__ void static_for(Fn &&fn) {
  if constexpr (F < F1) {
    fn(std::integral_constant<int, F>{});
    static_for<F + 1, F1>(fn);
  }
}

// ---- compile-time fragment tables ---------------------------------------------
// flat fragment index f: [0, FG) upper triangle of G by rows, [FG, FG + FH) H by columns
__host__ __device__ constexpr int tri_count(int MF) { return MF * (MF + 1) / 2; }
__host__ __device__ constexpr int frag_is_h(int MF, int f) { return f >= tri_count(MF); }
__host__ __device__ constexpr int frag_row(int MF, int f) {
  if (f >= tri_count(MF)) return (f - tri_count(MF)) % MF;
  int i = 0;
  while (f >= MF - i) {
    f -= MF - i;
    ++i;
  }
  return i;
}
__host__ __device__ constexpr int frag_col(int MF, int f) {   // column fragment inside G or H
  if (f >= tri_count(MF)) return (f - tri_count(MF)) / MF;
  int i = 0;
  while (f >= MF - i) {
    f -= MF - i;
    ++i;
  }
  return i + f;
}
// does the run [F0, F1) read row-fragment i of R (as A operand, or as B operand of G)?
__host__ __device__ constexpr bool need_a(int MF, int F0, int F1, int i) {
  for (int f = F0; f < F1; ++f) {
    if (frag_row(MF, f) == i) return true;
    if (!frag_is_h(MF, f) && frag_col(MF, f) == i) return true;
  }
  return false;
}
__host__ __device__ constexpr bool need_h(int MF, int F0, int F1, int j) {
  for (int f = F0; f < F1; ++f)
    if (frag_is_h(MF, f) && frag_col(MF, f) == j) return true;
  return false;
}

constexpr int GP_WARPS = 16;                         // consumer warps
constexpr int GP_CONSUMERS = GP_WARPS * 32;
constexpr int GP_PRODUCERS = 64;                     // producer threads (2 warps)
constexpr int GP_THREADS = GP_CONSUMERS + GP_PRODUCERS;

template <int MF, int MFH, int BC, int NG> struct GpCfg {
  static constexpr int KP = MF * 8, LP = MFH * 8;
  static constexpr int LDR = BC + 4;                 // == 4 (mod 16)
  static constexpr int STAGE = (KP + LP) * LDR;      // doubles per ring slot: [R; U_orig^T] tile
  static constexpr int FT = tri_count(MF) + MF * MFH;
  static constexpr int FPG = (FT + NG - 1) / NG;     // fragments per warp group
  static constexpr int KGW = GP_WARPS / NG;          // contraction slices per group
  static constexpr int KKW = BC / 4 / KGW;           // k steps per warp and stage
  // KGW % 4 == 0 balances the sub-partitions exactly; NG = 8 (K > 56: too many
  // accumulators per warp otherwise) balances them up to the last group's size
  static_assert(KGW % 4 == 0 || NG == 8, "slices must map onto the 4 SM sub-partitions");
  static_assert(KKW >= 1 && KKW * KGW * 4 == BC, "stage width vs. slices");
};

// one stage of DMMA work of warp group GID
template <int MF, int MFH, int BC, int NG, int GID, bool RAGGED>
__device__ __forceinline__ void gp_compute(double (&acc)[GpCfg<MF, MFH, BC, NG>::FPG][2],
                                           const double *Rs, int kg, int g, int q, int kvalid) {
  using C = GpCfg<MF, MFH, BC, NG>;
  constexpr int F0 = GID * C::FPG < C::FT ? GID * C::FPG : C::FT;
  constexpr int F1 = F0 + C::FPG < C::FT ? F0 + C::FPG : C::FT;
  if constexpr (F0 < F1) {
#pragma unroll
    for (int kk = 0; kk < C::KKW; ++kk) {
      const int k0 = (kk * C::KGW + kg) * 4;
      if (RAGGED && k0 >= kvalid) break;   // last stage of a row pitch that BC does not divide
      const double *ap = Rs + g * C::LDR + k0 + q;
      const double *bp = ap + C::KP * C::LDR;      // rows of U_orig^T
      double a[MF], bh[MFH > 0 ? MFH : 1];
      static_for<0, MF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if constexpr (need_a(MF, F0, F1, i)) a[i] = ap[i * 8 * C::LDR];
      });
      static_for<0, MFH>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (need_h(MF, F0, F1, j)) bh[j] = bp[j * 8 * C::LDR];
      });
      static_for<F0, F1>([&](auto fc) {
        constexpr int f = decltype(fc)::value;
        constexpr int i = frag_row(MF, f), j = frag_col(MF, f);
        if constexpr (frag_is_h(MF, f))
          dmma_8x8x4(acc[f - F0][0], acc[f - F0][1], a[i], bh[j]);
        else
          dmma_8x8x4(acc[f - F0][0], acc[f - F0][1], a[i], a[j]);
      });
    }
  }
}

// final store of warp group GID (its slice-0 warp holds the reduced sums)
template <int MF, int MFH, int BC, int NG, int GID>
__device__ __forceinline__ void gp_store(const double (&acc)[GpCfg<MF, MFH, BC, NG>::FPG][2],
                                         int K, int L, double *Gr, double *Hr, int g, int q) {
  using C = GpCfg<MF, MFH, BC, NG>;
  constexpr int F0 = GID * C::FPG < C::FT ? GID * C::FPG : C::FT;
  constexpr int F1 = F0 + C::FPG < C::FT ? F0 + C::FPG : C::FT;
  static_for<F0, F1>([&](auto fc) {
    constexpr int f = decltype(fc)::value;
    constexpr int i = frag_row(MF, f), j = frag_col(MF, f);
    const int row = i * 8 + g, col = j * 8 + 2 * q;
    const double c0 = acc[f - F0][0], c1 = acc[f - F0][1];
    if (row < K) {
      if constexpr (frag_is_h(MF, f)) {
        if (col < L) Hr[(size_t)row * L + col] = c0;
        if (col + 1 < L) Hr[(size_t)row * L + col + 1] = c1;
      } else {
        // upper triangle (diagonal fragments hold both) and its mirror image:
        // G comes out exactly symmetric
        if (col < K && col >= row) {
          Gr[(size_t)row * K + col] = c0;
          Gr[(size_t)col * K + row] = c0;
        }
        if (col + 1 < K && col + 1 >= row) {
          Gr[(size_t)row * K + col + 1] = c1;
          Gr[(size_t)(col + 1) * K + row] = c1;
        }
      }
    }
  });
}

constexpr int GP_MAX_SLOTS = 8;

template <int MF, int MFH, int BC, int NG>
__global__ void __launch_bounds__(GP_THREADS, 1)
gram_proj_kernel(const double *__restrict__ R, long long ldr, int K, int n_chunks, int nslot,
                 const double *__restrict__ UoT, long long uot_stride, int uot_div, int L,
                 double *__restrict__ G, double *__restrict__ H) {
  using C = GpCfg<MF, MFH, BC, NG>;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t full_bar[GP_MAX_SLOTS], empty_bar[GP_MAX_SLOTS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int ng = warp / C::KGW, kg = warp % C::KGW;
  const int r = blockIdx.x;

  // zero the ring once: the padding rows (>= K, >= L) are never written by the copies
  for (int e = tid; e < nslot * C::STAGE; e += GP_THREADS) sm[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < nslot; ++s) {
      mbar_init(&full_bar[s], GP_PRODUCERS);
      mbar_init(&empty_bar[s], GP_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= GP_WARPS) {
    // ---- producers: 16-byte segments of the rows of [R; U_orig^T], stage after stage ----
    constexpr int SEGS = BC / 2;                       // segments per tile row
    constexpr int RPP = SEGS <= GP_PRODUCERS ? GP_PRODUCERS / SEGS : 1;   // rows per pass
    const int p = tid - GP_CONSUMERS;
    const int row0 = SEGS <= GP_PRODUCERS ? p / SEGS : 0;
    const int seg0 = SEGS <= GP_PRODUCERS ? p - row0 * SEGS : p;
    const int nu = MFH > 0 ? L : 0;
    const double *Rr = R + (size_t)r * K * ldr;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int slot = ch % nslot;
      if (ch >= nslot) mbar_wait(&empty_bar[slot], ((ch / nslot) - 1) & 1);
      const long long b0 = (long long)ch * BC;
      const int kvalid = (int)min((long long)BC, ldr - b0);
      double *dst = sm + (size_t)slot * C::STAGE;
      for (int seg = seg0; seg < SEGS && seg * 2 < kvalid; seg += GP_PRODUCERS) {
        const double *src = Rr + (size_t)row0 * ldr + b0 + seg * 2;
        double *d = dst + row0 * C::LDR + seg * 2;
        for (int row = row0; row < K; row += RPP) {
          cp_async16(d, src);
          src += (size_t)RPP * ldr;
          d += RPP * C::LDR;
        }
        src = UoT + (size_t)(r / uot_div) * uot_stride + (size_t)row0 * ldr + b0 + seg * 2;
        d = dst + (C::KP + row0) * C::LDR + seg * 2;
        for (int row = row0; row < nu; row += RPP) {
          cp_async16(d, src);
          src += (size_t)RPP * ldr;
          d += RPP * C::LDR;
        }
      }
      cp_async_arrive(&full_bar[slot]);
    }
    return;
  }

  double acc[C::FPG][2];
#pragma unroll
  for (int f = 0; f < C::FPG; ++f) acc[f][0] = acc[f][1] = 0.0;

  for (int ch = 0; ch < n_chunks; ++ch) {
    const int slot = ch % nslot;
    mbar_wait(&full_bar[slot], (ch / nslot) & 1);
    const double *Rs = sm + (size_t)slot * C::STAGE;
    const int kvalid = (int)min((long long)BC, ldr - (long long)ch * BC);
    static_for<0, NG>([&](auto gc) {
      constexpr int GID = decltype(gc)::value;
      if (ng == GID) {
        if (kvalid == BC)
          gp_compute<MF, MFH, BC, NG, GID, false>(acc, Rs, kg, g, q, kvalid);
        else
          gp_compute<MF, MFH, BC, NG, GID, true>(acc, Rs, kg, g, q, kvalid);
      }
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);
  }
  consumer_sync();

  // pairwise reduction of the contraction slices: slices [s, 2s) hand their
  // sums to slices [0, s)
  double2 *red = reinterpret_cast<double2 *>(sm);
#pragma unroll
  for (int s = C::KGW / 2; s >= 1; s >>= 1) {
    if (kg >= s && kg < 2 * s) {
      double2 *p = red + ((size_t)(ng * s + (kg - s)) * C::FPG) * 32 + lane;
#pragma unroll
      for (int f = 0; f < C::FPG; ++f) p[f * 32] = make_double2(acc[f][0], acc[f][1]);
    }
    consumer_sync();
    if (kg < s) {
      const double2 *p = red + ((size_t)(ng * s + kg) * C::FPG) * 32 + lane;
#pragma unroll
      for (int f = 0; f < C::FPG; ++f) {
        const double2 v = p[f * 32];
        acc[f][0] += v.x;
        acc[f][1] += v.y;
      }
    }
    consumer_sync();
  }
  if (kg == 0) {
    double *Gr = G + (size_t)r * K * K;
    double *Hr = H ? H + (size_t)r * K * L : nullptr;
    static_for<0, NG>([&](auto gc) {
      constexpr int GID = decltype(gc)::value;
      if (ng == GID) gp_store<MF, MFH, BC, NG, GID>(acc, K, L, Gr, Hr, g, q);
    });
  }
}

constexpr size_t GP_SMEM_MAX = 226 * 1024;   // + the static mbarrier arrays

template <int MF, int MFH, int BC, int NG>
int launch_cfg(plsb_ctx *h, const double *R, long long ldr, int count, int K, const double *UoT,
               int L, double *G, double *H, cudaStream_t st, long long us, int ud) {
  using C = GpCfg<MF, MFH, BC, NG>;
  const size_t stage = sizeof(double) * C::STAGE;
  const size_t red = sizeof(double2) * 32 * C::FPG * (GP_WARPS / 2);
  PLSB_CHECK(ldr % 4 == 0, PLSB_ERR_ARG, "gram_proj: row pitch %lld not a multiple of 4", ldr);
  const int n_chunks = (int)((ldr + BC - 1) / BC);
  int nslot = (int)std::min<size_t>(GP_MAX_SLOTS, GP_SMEM_MAX / stage);
  nslot = std::max(2, std::min(nslot, n_chunks + 1));
  nslot = tune_int("PLSB_GP_SLOTS", nslot);
  const size_t smem = std::max(stage * nslot, red);
  PLSB_CHECK(smem <= GP_SMEM_MAX, PLSB_ERR_ARG, "gram_proj: %zu bytes of shared memory", smem);
  auto kern = gram_proj_kernel<MF, MFH, BC, NG>;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<count, GP_THREADS, smem, st>>>(R, ldr, K, n_chunks, nslot, UoT, us, ud, L, G, H);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

// stage width / group count per problem size: enough DMMA work between two
// barriers, as many bytes in flight as shared memory allows
template <int MF, int MFH>
int launch_mf(plsb_ctx *h, const double *R, long long ldr, int count, int K, const double *UoT,
              int L, double *G, double *H, cudaStream_t st, long long us, int ud) {
  // ~40-70 KB per stage: enough DMMA work per mbarrier round trip, >= 3 slots in flight
  constexpr int FT = tri_count(MF) + MF * MFH;
  constexpr int ROWS = (MF + MFH) * 8;
  if constexpr (FT <= 4)
    return launch_cfg<MF, MFH, (ROWS <= 16 ? 512 : 256), 1>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
  else if constexpr (FT <= 12)
    return launch_cfg<MF, MFH, 256, 2>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
  else if constexpr (ROWS <= 48)
    return launch_cfg<MF, MFH, 128, 4>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
  else if constexpr (ROWS <= 112)
    return launch_cfg<MF, MFH, 64, 4>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
  else
    return launch_cfg<MF, MFH, 32, 8>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
}

template <int MF>
int launch_proj(plsb_ctx *h, bool proj, const double *R, long long ldr, int count, int K,
                const double *UoT, int L, double *G, double *H, cudaStream_t st, long long us,
                int ud) {
  if (proj) return launch_mf<MF, MF>(h, R, ldr, count, K, UoT, L, G, H, st, us, ud);
  return launch_mf<MF, 0>(h, R, ldr, count, K, nullptr, 0, G, nullptr, st, 0, 1);
}

}  // namespace

// R must be a (count*K rows, ldr) buffer whose row pitch ldr is a multiple of 4
// and whose columns >= B are zero (the GEMM's padded output).  UoT is U_orig
// transposed, (L rows, ldr) with zero columns >= B; L and K must span the same
// number of 8-column fragments (the engine has L == K); null: G only.  Matrix r
// is projected on the block UoT + (r / uot_div) * uot_stride (split-half: one
// block per permutation, shared by its split halves).
int launch_gram_proj(plsb_ctx *h, const double *R, long long ldr, int count, int K,
                     const double *UoT, int L, double *G, double *H, cudaStream_t st,
                     long long uot_stride, int uot_div) {
  KernelTimer kt(h, KC_GRAM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1, PLSB_ERR_ARG, "gram_proj: K=%d", K);
  if (K > MAX_K)   // beyond the fragment tables: the generic tiled kernel (columns >= B are zero)
    return launch_gram_proj_generic(h, R, ldr, count, K, (int)ldr, UoT, L, G, H, st, uot_stride,
                                    uot_div);
  const bool proj = UoT && H;
  if (small_k_applies(K, L, proj))   // few latent variables: un-padded FMA work, TMA stream
    return launch_gram_proj_small(h, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
  PLSB_CHECK(!proj || cdiv(L, 8) == cdiv(K, 8), PLSB_ERR_ARG,
             "gram_proj: L=%d and K=%d must span the same number of fragments", L, K);
  switch (cdiv(K, 8)) {
    case 1: return launch_proj<1>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 2: return launch_proj<2>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 3: return launch_proj<3>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 4: return launch_proj<4>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 5: return launch_proj<5>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 6: return launch_proj<6>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 7: return launch_proj<7>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 8: return launch_proj<8>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    case 9: return launch_proj<9>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
    default: return launch_proj<10>(h, proj, R, ldr, count, K, UoT, L, G, H, st, uot_stride, uot_div);
  }
}

}  // namespace plsb
