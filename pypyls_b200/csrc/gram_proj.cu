// gram_proj:  [G | H] = R [R^T | U_orig]   (K x (K + L), contraction over B)
//
// FP64 tensor-core (DMMA m8n8k4) streaming contraction over one stored
// cross-covariance matrix R (K rows x ldr) per CTA -- all the small SVD +
// Procrustes step of a bootstrap needs (pyls/compute.py:36-49, 260: the
// randomized SVD of R and the U_orig^T U_boot product collapse onto G and H).
//
// Work decomposition (16 warps):
//   * G is symmetric: only the MF (MF + 1) / 2 upper fragments (8 x 8) are
//     computed; with the MF * MFH fragments of H that is FT fragments, dealt
//     in contiguous runs to NG warp groups (compile-time tables, so every
//     accumulator stays in a register);
//   * inside a group the KGW = 16 / NG warps split the contraction steps of a
//     stage.  Warp w has (group w / KGW, slice w % KGW): KGW is a multiple of
//     4 (except for K > 56), so the four warps that share an SM sub-partition's
//     DMMA pipe are the same slice of each group and the sub-partitions are
//     balanced whatever the group sizes are;
//   * R and U_orig tiles of BC columns are staged by cp.async through a ring
//     of `nslot` slots (as many as fit shared memory), one barrier per stage;
//   * the slices are reduced pairwise through shared memory at the end (fixed
//     order: results are bit-reproducible).
//
// Shared-memory leading dimensions are == 4 or 12 (mod 16) doubles so that
// every fragment load (8 x 4 doubles) is bank-conflict free.
#include "common.cuh"

#include <type_traits>

namespace plsb {
namespace {

__device__ __forceinline__ void cp_async16z(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8z(void *smem, const void *gmem, int src_bytes) {
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

constexpr int GP_THREADS = 512;
constexpr int GP_WARPS = 16;

template <int MF, int MFH, int BC, int NG> struct GpCfg {
  static constexpr int KP = MF * 8, LP = MFH * 8;
  static constexpr int LDR = BC + 4;                 // == 4 (mod 16)
  static constexpr int LDU = LP + 4;                 // == 4 or 12 (mod 16)
  static constexpr int STAGE = KP * LDR + BC * LDU;  // doubles per ring slot
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
template <int MF, int MFH, int BC, int NG, int GID>
__device__ __forceinline__ void gp_compute(double (&acc)[GpCfg<MF, MFH, BC, NG>::FPG][2],
                                           const double *Rs, const double *Us, int kg, int g,
                                           int q) {
  using C = GpCfg<MF, MFH, BC, NG>;
  constexpr int F0 = GID * C::FPG < C::FT ? GID * C::FPG : C::FT;
  constexpr int F1 = F0 + C::FPG < C::FT ? F0 + C::FPG : C::FT;
  if constexpr (F0 < F1) {
#pragma unroll
    for (int kk = 0; kk < C::KKW; ++kk) {
      const int k0 = (kk * C::KGW + kg) * 4;
      const double *ap = Rs + g * C::LDR + k0 + q;
      const double *bp = Us + (k0 + q) * C::LDU + g;
      double a[MF], bh[MFH > 0 ? MFH : 1];
      static_for<0, MF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if constexpr (need_a(MF, F0, F1, i)) a[i] = ap[i * 8 * C::LDR];
      });
      static_for<0, MFH>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (need_h(MF, F0, F1, j)) bh[j] = bp[j * 8];
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

template <int MF, int MFH, int BC, int NG>
__global__ void __launch_bounds__(GP_THREADS, 1)
gram_proj_kernel(const double *__restrict__ R, long long ldr, int K, int B, int n_chunks,
                 int nslot, const double *__restrict__ Uo, int L, double *__restrict__ G,
                 double *__restrict__ H) {
  using C = GpCfg<MF, MFH, BC, NG>;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int ng = warp / C::KGW, kg = warp % C::KGW;
  const int r = blockIdx.x;
  const double *Rr = R + (size_t)r * K * ldr;

  // zero the padding that the copies never touch: rows >= K of Rs, columns >= L of Us
  for (int s = 0; s < nslot; ++s) {
    double *Rs = sm + (size_t)s * C::STAGE, *Us = Rs + C::KP * C::LDR;
    for (int e = tid; e < (C::KP - K) * C::LDR; e += GP_THREADS) Rs[K * C::LDR + e] = 0.0;
    if (C::LP > L)
      for (int e = tid; e < BC * (C::LP - L); e += GP_THREADS) {
        const int b = e / (C::LP - L), l = L + e - b * (C::LP - L);
        Us[b * C::LDU + l] = 0.0;
      }
  }

  const bool l_even = (L & 1) == 0;
  auto load = [&](int ch) {
    if (ch < n_chunks) {
      double *Rs = sm + (size_t)(ch % nslot) * C::STAGE, *Us = Rs + C::KP * C::LDR;
      const int b0 = ch * BC;
      for (int e = tid; e < K * (BC / 2); e += GP_THREADS) {
        const int c = e / (BC / 2), seg = e - c * (BC / 2);
        const bool ok = b0 + seg * 2 < ldr;
        cp_async16z(Rs + c * C::LDR + seg * 2, ok ? Rr + (size_t)c * ldr + b0 + seg * 2 : Rr,
                    ok ? 16 : 0);
      }
      if (MFH > 0) {
        if (l_even) {
          const int hl = L >> 1;
          for (int e = tid; e < BC * hl; e += GP_THREADS) {
            const int b = e / hl, l = (e - b * hl) * 2;
            const bool ok = b0 + b < B;
            cp_async16z(Us + b * C::LDU + l, ok ? Uo + (size_t)(b0 + b) * L + l : Uo, ok ? 16 : 0);
          }
        } else {
          for (int e = tid; e < BC * L; e += GP_THREADS) {
            const int b = e / L, l = e - b * L;
            const bool ok = b0 + b < B;
            cp_async8z(Us + b * C::LDU + l, ok ? Uo + (size_t)(b0 + b) * L + l : Uo, ok ? 8 : 0);
          }
        }
      }
    }
    cp_async_commit();
  };

  double acc[C::FPG][2];
#pragma unroll
  for (int f = 0; f < C::FPG; ++f) acc[f][0] = acc[f][1] = 0.0;

  for (int s = 0; s < nslot - 1; ++s) load(s);
  for (int ch = 0; ch < n_chunks; ++ch) {
    // groups committed so far: ch + nslot - 1; stage ch has landed when at most
    // nslot - 2 of the most recent ones are pending
    switch (nslot) {
      case 2: cp_async_wait<0>(); break;
      case 3: cp_async_wait<1>(); break;
      case 4: cp_async_wait<2>(); break;
      case 5: cp_async_wait<3>(); break;
      default: cp_async_wait<4>(); break;
    }
    __syncthreads();
    load(ch + nslot - 1);   // refills the slot consumed in the previous iteration
    const double *Rs = sm + (size_t)(ch % nslot) * C::STAGE, *Us = Rs + C::KP * C::LDR;
    static_for<0, NG>([&](auto gc) {
      constexpr int GID = decltype(gc)::value;
      if (ng == GID) gp_compute<MF, MFH, BC, NG, GID>(acc, Rs, Us, kg, g, q);
    });
  }
  cp_async_wait<0>();
  __syncthreads();

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
    __syncthreads();
    if (kg < s) {
      const double2 *p = red + ((size_t)(ng * s + kg) * C::FPG) * 32 + lane;
#pragma unroll
      for (int f = 0; f < C::FPG; ++f) {
        const double2 v = p[f * 32];
        acc[f][0] += v.x;
        acc[f][1] += v.y;
      }
    }
    __syncthreads();
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

constexpr size_t GP_SMEM_MAX = 227 * 1024;

template <int MF, int MFH, int BC, int NG>
int launch_cfg(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
               const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  using C = GpCfg<MF, MFH, BC, NG>;
  const size_t stage = sizeof(double) * C::STAGE;
  const size_t red = sizeof(double2) * 32 * C::FPG * (GP_WARPS / 2);
  const int n_chunks = (int)((ldr + BC - 1) / BC);
  int nslot = (int)std::min<size_t>(6, GP_SMEM_MAX / stage);
  nslot = std::max(2, std::min(nslot, n_chunks + 1));
  const size_t smem = std::max(stage * nslot, red);
  PLSB_CHECK(smem <= GP_SMEM_MAX, PLSB_ERR_ARG, "gram_proj: %zu bytes of shared memory", smem);
  auto kern = gram_proj_kernel<MF, MFH, BC, NG>;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<count, GP_THREADS, smem, st>>>(R, ldr, K, B, n_chunks, nslot, Uo, L, G, H);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

// stage width / group count per problem size: enough DMMA work between two
// barriers, as many bytes in flight as shared memory allows
template <int MF, int MFH>
int launch_mf(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
              const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  constexpr int FT = tri_count(MF) + MF * MFH;
  if constexpr (FT <= 4)
    return launch_cfg<MF, MFH, 256, 1>(h, R, ldr, count, K, B, Uo, L, G, H, st);
  else if constexpr (FT <= 12)
    return launch_cfg<MF, MFH, 256, 2>(h, R, ldr, count, K, B, Uo, L, G, H, st);
  else if constexpr (MF <= 3)
    return launch_cfg<MF, MFH, 128, 4>(h, R, ldr, count, K, B, Uo, L, G, H, st);
  else if constexpr (MF <= 7)
    return launch_cfg<MF, MFH, 64, 4>(h, R, ldr, count, K, B, Uo, L, G, H, st);
  else
    return launch_cfg<MF, MFH, 32, 8>(h, R, ldr, count, K, B, Uo, L, G, H, st);
}

template <int MF>
int launch_proj(plsb_ctx *h, bool proj, const double *R, long long ldr, int count, int K, int B,
                const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  if (proj) return launch_mf<MF, MF>(h, R, ldr, count, K, B, Uo, L, G, H, st);
  return launch_mf<MF, 0>(h, R, ldr, count, K, B, nullptr, 0, G, nullptr, st);
}

}  // namespace

// R must be a (count*K rows, ldr) buffer whose row pitch ldr is even and whose
// columns >= B are zero (the GEMM's padded output).  Uo (B, L) needs L == K
// rounded to the same number of 8-column fragments (the engine has L == K).
int launch_gram_proj(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                     const double *Uo, int L, double *G, double *H, cudaStream_t st) {
  KernelTimer kt(h, KC_GRAM, st);
  if (count <= 0) return PLSB_OK;
  PLSB_CHECK(K >= 1 && K <= MAX_K, PLSB_ERR_ARG, "gram_proj: K=%d outside [1,%d]", K, MAX_K);
  PLSB_CHECK(ldr % 2 == 0 && ldr >= B, PLSB_ERR_ARG, "gram_proj: bad row pitch %lld", ldr);
  const bool proj = Uo && H;
  PLSB_CHECK(!proj || cdiv(L, 8) == cdiv(K, 8), PLSB_ERR_ARG,
             "gram_proj: L=%d and K=%d must span the same number of fragments", L, K);
  switch (cdiv(K, 8)) {
    case 1: return launch_proj<1>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 2: return launch_proj<2>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 3: return launch_proj<3>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 4: return launch_proj<4>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 5: return launch_proj<5>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 6: return launch_proj<6>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 7: return launch_proj<7>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 8: return launch_proj<8>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    case 9: return launch_proj<9>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
    default: return launch_proj<10>(h, proj, R, ldr, count, K, B, Uo, L, G, H, st);
  }
}

}  // namespace plsb
