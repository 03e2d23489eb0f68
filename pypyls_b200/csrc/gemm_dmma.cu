// FP64 tensor-core (DMMA) cross-covariance GEMM for sm_100a.
//
//   C (M, N) = A (M, Kd) @ X (Kd, N)
//
// A is the stack of per-resample left operands (a few rows per resample, built
// by operands.cu), X is the data matrix shared by every resample.  This is the
// contraction of compute.xcorr (pyls/compute.py:92, `Yn.T @ Xn`) for thousands
// of resamples at once.  tcgen05 has no f64 kind, so the tensor path for fp64
// on sm_100a is mma.sync.m8n8k4 (SASS DMMA.8x8x4).
//
// Tiling: CTA 128x128, 8 warps as 2 (M) x 4 (N), warp tile 64x32 = 8 x 4
// fragments, BK = 16 per pipeline stage, 4-stage cp.async ring.  The (N tile,
// k chunk) loops are flattened into one sequence so the ring never drains
// between N tiles.  Shared-memory leading dimensions are == 4 (mod 16) doubles,
// which makes every fragment load (8 rows x 4 k) bank-conflict free.
// A second instantiation with a 64x128 CTA tile (warp tile 32x32, two CTAs per
// SM) serves launches whose contraction per tile is so short that the store
// epilogue dominates: of two resident CTAs one computes while the other writes.
//
// Epilogues:
//   STORE     write C (optionally through a row map: operand row -> output row)
//   ROWSUMSQ  rowsq[split][m] = sum_n C[m,n]^2  (rotated permutation singular
//             values, pyls/base.py:699-700, without materialising C)
// Optional per-M-tile contraction ranges skip the structural zeros of
// block-diagonal operands (one block per group x condition cell).
#include "common.cuh"

namespace plsb {

namespace {

constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
constexpr int LDA_S = BK + 4;           // 20
constexpr int LDB_S = BN + 4;           // 132
constexpr int B_STAGE = BK * LDB_S;     // doubles
constexpr int NTHREADS = 256;
constexpr int NF = 4;                   // N fragments per warp tile (MF = 8 or 4 in M)
// ring depth and staged column-scale rows per tile, by tile height
__host__ __device__ constexpr int stages_of(int MF) { return MF == 8 ? 4 : 3; }
__host__ __device__ constexpr int max_slots_of(int MF) { return MF == 8 ? 16 : 8; }
__host__ __device__ constexpr int ring_doubles(int MF) { return stages_of(MF) * (2 * MF * 8 * LDA_S + B_STAGE); }
__host__ __device__ constexpr int scale_doubles(int MF) { return stages_of(MF) * max_slots_of(MF) * LDB_S; }
__host__ __device__ constexpr int smem_bytes(int MF, bool scaled) {
  return (ring_doubles(MF) + (scaled ? scale_doubles(MF) : 0)) * (int)sizeof(double);
}

enum { EPI_STORE = 0, EPI_ROWSUMSQ = 1 };

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
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

// MF = M fragments per warp: CTA tile (2 * MF * 8) x 128
template <int EPI, bool SQB, int MF>
__global__ void __launch_bounds__(NTHREADS, MF == 8 ? 1 : 2)
xcov_gemm_kernel(const double *__restrict__ A, int lda, const double *__restrict__ X, int ldx,
                 int n_mtiles, int n_ntiles, int nt_per_split, const int4 *__restrict__ kranges,
                 int Kd, double *__restrict__ C, long long ldc, const int *__restrict__ row_map,
                 const double *__restrict__ scale, int scale_div, long long lds,
                 double *__restrict__ rowsq, int M_pad) {
  constexpr int TM = 2 * MF * 8;          // rows of the CTA tile
  constexpr int A_STAGE = TM * LDA_S;     // doubles
  constexpr int STAGES = stages_of(MF);
  constexpr int MAXS = max_slots_of(MF);
  extern __shared__ __align__(16) double smem[];
  double *As = smem;
  double *Bs = smem + STAGES * A_STAGE;
  double *Sb = smem + ring_doubles(MF);   // [STAGES][MAXS][LDB_S] staged column scales
  __shared__ int s_slot[TM];              // tile row -> staged scale row
  __shared__ int s_srow[MAXS];            // staged scale row -> row of `scale`
  __shared__ int s_nslot;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;

  const int split = blockIdx.x / n_mtiles;
  const int mtile = blockIdx.x - split * n_mtiles;
  const int nt0 = split * nt_per_split;
  const int nt1 = min(nt0 + nt_per_split, n_ntiles);
  if (nt0 >= nt1) return;

  int kbeg = 0, kend = Kd, vbeg = 0, vend = Kd;
  if (kranges) {
    // the ranges are tabulated per GEMM_BM (128) rows
    int4 kr = kranges[(mtile * TM) / BM];
    kbeg = kr.x;
    kend = kr.y;
    vbeg = kr.z;   // k steps outside [vbeg, vend) multiply structural zeros of A: skipped
    vend = kr.w;
  }
  const int nkc = (kend - kbeg + BK - 1) / BK;     // k chunks per N tile
  const int total = (nt1 - nt0) * nkc;             // flattened pipeline steps

  const double *Ablk = A + (size_t)mtile * TM * lda;

  // Column scales (bootstrap z-scores of X) are shared by the scale_div operand
  // rows of one (resample, cell): the few distinct rows a tile needs are staged
  // in shared memory next to the tile's first k chunk, so that the epilogue
  // does not wait on global loads.  Falls back to direct loads when a tile
  // touches more than MAXS rows.
  int nslot = 0;
  int orow_i[MF], slot_i[MF];
  if (EPI == EPI_STORE) {
    if (scale) {
      if (tid < TM) {
        const int m = mtile * TM + tid;
        const int orow = row_map ? row_map[m] : m;
        s_slot[tid] = orow >= 0 ? orow / scale_div : -1;
      }
      __syncthreads();
      if (tid == 0) {
        int n = 0, prev = -1;
        for (int t = 0; t < TM; ++t) {
          const int sr = s_slot[t];
          if (sr < 0) {
            s_slot[t] = 0;
            continue;
          }
          if (sr != prev) {
            if (n < MAXS) s_srow[n] = sr;
            ++n;
            prev = sr;
          }
          s_slot[t] = n - 1;
        }
        s_nslot = n <= MAXS ? n : 0;
      }
      __syncthreads();
      nslot = s_nslot;
    }
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      const int ml = wm * MF * 8 + i * 8 + g;
      const int m = mtile * TM + ml;
      orow_i[i] = row_map ? row_map[m] : m;
      slot_i[i] = nslot > 0 ? s_slot[ml] : 0;
    }
  }

  // producer: issue the loads of flattened step `s` into ring slot s % STAGES
  auto issue = [&](int s) {
    if (s < total) {
      const int nt = nt0 + s / nkc;
      const int kc = s - (s / nkc) * nkc;
      const int k0 = kbeg + kc * BK;
      double *as = As + (s % STAGES) * A_STAGE;
      double *bs = Bs + (s % STAGES) * B_STAGE;
      // A: TM rows x 16 doubles = TM * 8 x 16 B, TM / 32 per thread
#pragma unroll
      for (int it = 0; it < TM / 32; ++it) {
        int c = tid + it * NTHREADS;
        int row = c >> 3, seg = c & 7;
        cp_async16(as + row * LDA_S + seg * 2, Ablk + (size_t)row * lda + k0 + seg * 2);
      }
      // X: 16 rows x 128 doubles = 1024 x 16 B, 4 per thread
      const double *xg = X + (size_t)k0 * ldx + (size_t)nt * BN;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        int c = tid + it * NTHREADS;
        int row = c >> 6, seg = c & 63;
        cp_async16(bs + row * LDB_S + seg * 2, xg + (size_t)row * ldx + seg * 2);
      }
      if (EPI == EPI_STORE && kc == 0 && nslot > 0) {
        // scale rows of this N tile; buffer (tile ordinal) % STAGES is free again:
        // the tile that used it finished its epilogue >= 1 iteration ago
        double *sb = Sb + ((s / nkc) % STAGES) * MAXS * LDB_S;
        for (int e = tid; e < nslot * (BN / 2); e += NTHREADS) {
          const int sl = e / (BN / 2), seg = e - sl * (BN / 2);
          cp_async16(sb + sl * LDB_S + seg * 2,
                     scale + (size_t)s_srow[sl] * lds + (size_t)nt * BN + seg * 2);
        }
      }
    }
    cp_async_commit();
  };

  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double rsq[MF];
#pragma unroll
  for (int i = 0; i < MF; ++i) rsq[i] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  int kc = 0, nt = nt0;
  for (int s = 0; s < total; ++s) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // refill the slot consumed in the previous iteration
    issue(s + STAGES - 1);

    const double *as = As + (s % STAGES) * A_STAGE + (wm * MF * 8 + g) * LDA_S + q;
    const double *bs = Bs + (s % STAGES) * B_STAGE + q * LDB_S + wn * 32 + g;
    const int kabs = kbeg + kc * BK;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      if (kabs + kk * 4 < vbeg || kabs + kk * 4 >= vend) continue;
      double af[MF], bf[NF];
#pragma unroll
      for (int i = 0; i < MF; ++i) af[i] = as[i * 8 * LDA_S + kk * 4];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        double v = bs[kk * 4 * LDB_S + j * 8];
        bf[j] = SQB ? v * v : v;
      }
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }

    if (++kc == nkc) {
      // ---- epilogue of N tile `nt` ----
      if (EPI == EPI_STORE) {
        const double *sb = Sb + (((s / nkc) % STAGES) * MAXS) * LDB_S + wn * 32 + 2 * q;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
          const int orow = orow_i[i];
          if (orow >= 0) {
            const size_t col = (size_t)nt * BN + wn * 32 + 2 * q;
            double *crow = C + (size_t)orow * ldc + col;
            if (scale) {
              const double *srow = nslot > 0 ? sb + slot_i[i] * LDB_S
                                             : scale + (size_t)(orow / scale_div) * lds + col;
#pragma unroll
              for (int j = 0; j < NF; ++j) {
                const double2 sc = *reinterpret_cast<const double2 *>(srow + j * 8);
                *reinterpret_cast<double2 *>(crow + j * 8) =
                    make_double2(acc[i][j][0] * sc.x, acc[i][j][1] * sc.y);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NF; ++j)
                *reinterpret_cast<double2 *>(crow + j * 8) =
                    make_double2(acc[i][j][0], acc[i][j][1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j)
            rsq[i] += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
      }
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      kc = 0;
      ++nt;
    }
  }

  if (EPI == EPI_ROWSUMSQ) {
    // reduce over the 4 lanes of a quad, then over the 4 N-warps through smem
    cp_async_wait<0>();
    __syncthreads();
    double *red = smem;  // [4 wn][TM rows]
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      double v = rsq[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (q == 0) red[wn * TM + wm * MF * 8 + i * 8 + g] = v;
    }
    __syncthreads();
    if (tid < TM) {
      double v = red[tid] + red[TM + tid] + red[2 * TM + tid] + red[3 * TM + tid];
      rowsq[(size_t)split * M_pad + (size_t)mtile * TM + tid] = v;
    }
  }
}

// ---- v2: the same tiling with the producer interleaved into the tensor stream ----------
// In the kernel above every warp stops at the CTA barrier, then issues its share of the
// next stage's cp.async (address arithmetic + 8 LDGSTS), then loads fragments, and only
// then feeds the DMMA pipe again: ~300 of the ~4100 cycles of a k chunk with the tensor
// pipe idle.  Here the copies of ring step s + STAGES - 1 are issued ONE AT A TIME between
// the DMMA groups of step s (source pointers advance incrementally, no divisions), so that
// after the barrier a warp only loads its first fragments before the pipe is busy again.
// Also skips the k steps beyond `k_valid` (rows of the data matrix that are zero padding).
template <int EPI, bool SQB, int MF>
__global__ void __launch_bounds__(NTHREADS, MF == 8 ? 1 : 2)
xcov_gemm2_kernel(const double *__restrict__ A, int lda, const double *__restrict__ X, int ldx,
                  int n_mtiles, int n_ntiles, int nt_per_split, const int4 *__restrict__ kranges,
                  int Kd, int k_valid, double *__restrict__ C, long long ldc,
                  const int *__restrict__ row_map, const double *__restrict__ scale,
                  int scale_div, long long lds, double *__restrict__ rowsq, int M_pad) {
  constexpr int TM = 2 * MF * 8;          // rows of the CTA tile
  constexpr int A_STAGE = TM * LDA_S;     // doubles
  constexpr int STAGES = stages_of(MF);
  constexpr int MAXS = max_slots_of(MF);
  constexpr int NA = TM / 32;             // A copies per thread and stage
  constexpr int NX = 4;                   // X copies per thread and stage
  constexpr int NP = NA + NX;
  extern __shared__ __align__(16) double smem[];
  double *As = smem;
  double *Bs = smem + STAGES * A_STAGE;
  double *Sb = smem + ring_doubles(MF);   // [STAGES][MAXS][LDB_S] staged column scales
  __shared__ int s_slot[TM];              // tile row -> staged scale row
  __shared__ int s_srow[MAXS];            // staged scale row -> row of `scale`
  __shared__ int s_nslot;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;

  const int split = blockIdx.x / n_mtiles;
  const int mtile = blockIdx.x - split * n_mtiles;
  const int nt0 = split * nt_per_split;
  const int nt1 = min(nt0 + nt_per_split, n_ntiles);
  if (nt0 >= nt1) return;
  int kbeg = 0, kend = Kd, vbeg = 0, vend = min(Kd, (k_valid + 3) & ~3);
  if (kranges) {
    int4 kr = kranges[(mtile * TM) / BM];
    kbeg = kr.x;
    kend = kr.y;
    vbeg = kr.z;
    vend = kr.w;
  } else {
    kend = min(Kd, (vend + BK - 1) / BK * BK);
  }
  const int nkc = (kend - kbeg + BK - 1) / BK;     // k chunks per N tile
  const int total = (nt1 - nt0) * nkc;             // flattened pipeline steps

  const double *Ablk = A + (size_t)mtile * TM * lda;

  int nslot = 0;
  int orow_i[MF], slot_i[MF];
  if (EPI == EPI_STORE) {
    if (scale) {
      if (tid < TM) {
        const int m = mtile * TM + tid;
        const int orow = row_map ? row_map[m] : m;
        s_slot[tid] = orow >= 0 ? orow / scale_div : -1;
      }
      __syncthreads();
      if (tid == 0) {
        int n = 0, prev = -1;
        for (int t = 0; t < TM; ++t) {
          const int sr = s_slot[t];
          if (sr < 0) {
            s_slot[t] = 0;
            continue;
          }
          if (sr != prev) {
            if (n < MAXS) s_srow[n] = sr;
            ++n;
            prev = sr;
          }
          s_slot[t] = n - 1;
        }
        s_nslot = n <= MAXS ? n : 0;
      }
      __syncthreads();
      nslot = s_nslot;
    }
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      const int ml = wm * MF * 8 + i * 8 + g;
      const int m = mtile * TM + ml;
      orow_i[i] = row_map ? row_map[m] : m;
      slot_i[i] = nslot > 0 ? s_slot[ml] : 0;
    }
  }

  // ---- producer state: one source pointer per copy, advanced per step ----
  const double *a_src[NA];
  unsigned a_dst[NA];
#pragma unroll
  for (int it = 0; it < NA; ++it) {
    const int c = tid + it * NTHREADS;
    const int row = c >> 3, seg = c & 7;
    a_src[it] = Ablk + (size_t)row * lda + kbeg + seg * 2;
    a_dst[it] = (unsigned)__cvta_generic_to_shared(As + row * LDA_S + seg * 2);
  }
  const double *x_src[NX];
  unsigned x_dst[NX];
#pragma unroll
  for (int it = 0; it < NX; ++it) {
    const int c = tid + it * NTHREADS;
    const int row = c >> 6, seg = c & 63;
    x_src[it] = X + (size_t)(kbeg + row) * ldx + (size_t)nt0 * BN + seg * 2;
    x_dst[it] = (unsigned)__cvta_generic_to_shared(Bs + row * LDB_S + seg * 2);
  }
  const size_t x_step = (size_t)BK * ldx;                       // next k chunk
  const long long x_wrap = (long long)BN - (long long)nkc * BK * ldx;   // next N tile, first chunk
  int p_step = 0, p_kc = 0, p_slot = 0, p_tile = 0;   // step being produced, its chunk, ring slot, tile ordinal

  auto copy16 = [](unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
  };
  // copy number `pi` of the step being produced (A first, then X)
  auto piece = [&](int pi) {
    if (p_step >= total) return;
    if (pi < NA) {
#pragma unroll
      for (int it = 0; it < NA; ++it)
        if (it == pi) copy16(a_dst[it] + p_slot * (A_STAGE * 8), a_src[it] + p_kc * BK);
    } else {
#pragma unroll
      for (int it = 0; it < NX; ++it)
        if (it == pi - NA) copy16(x_dst[it] + p_slot * (B_STAGE * 8), x_src[it]);
    }
  };
  // all copies of the step issued: commit, advance the producer to the next step
  auto produced = [&]() {
    if (p_step < total) {
      if (EPI == EPI_STORE && p_kc == 0 && nslot > 0) {
        // scale rows of this N tile; buffer (tile ordinal) % STAGES is free again:
        // the tile that used it finished its epilogue >= 1 iteration ago
        double *sb = Sb + (p_tile % STAGES) * MAXS * LDB_S;
        for (int e = tid; e < nslot * (BN / 2); e += NTHREADS) {
          const int sl = e / (BN / 2), seg = e - sl * (BN / 2);
          cp_async16(sb + sl * LDB_S + seg * 2,
                     scale + (size_t)s_srow[sl] * lds + (size_t)(nt0 + p_tile) * BN + seg * 2);
        }
      }
      ++p_step;
      p_slot = p_slot + 1 == STAGES ? 0 : p_slot + 1;
      if (++p_kc == nkc) {
        p_kc = 0;
        ++p_tile;
#pragma unroll
        for (int it = 0; it < NX; ++it) x_src[it] += x_wrap + (long long)x_step;
      } else {
#pragma unroll
        for (int it = 0; it < NX; ++it) x_src[it] += x_step;
      }
    }
    cp_async_commit();
  };

  double acc[MF][NF][2];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double rsq[MF];
#pragma unroll
  for (int i = 0; i < MF; ++i) rsq[i] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
#pragma unroll
    for (int pi = 0; pi < NP; ++pi) piece(pi);
    produced();
  }

  int kc = 0, nt = nt0, slot = 0;
  for (int s = 0; s < total; ++s) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();

    const double *as = As + slot * A_STAGE + (wm * MF * 8 + g) * LDA_S + q;
    const double *bs = Bs + slot * B_STAGE + q * LDB_S + wn * 32 + g;
    const int kabs = kbeg + kc * BK;
    if (kabs >= vbeg && kabs + BK <= vend) {
      // whole chunk: one copy of the step being produced after every group of DMMAs
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double af[MF], bf[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) af[i] = as[i * 8 * LDA_S + kk * 4];
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          double v = bs[kk * 4 * LDB_S + j * 8];
          bf[j] = SQB ? v * v : v;
        }
#pragma unroll
        for (int i = 0; i < MF; ++i) {
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
          if (kk * MF + i < NP) piece(kk * MF + i);
        }
      }
    } else {
      // chunk with k steps that multiply structural zeros (skipped): copies first
#pragma unroll
      for (int pi = 0; pi < NP; ++pi) piece(pi);
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        if (kabs + kk * 4 < vbeg || kabs + kk * 4 >= vend) continue;
        double af[MF], bf[NF];
#pragma unroll
        for (int i = 0; i < MF; ++i) af[i] = as[i * 8 * LDA_S + kk * 4];
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          double v = bs[kk * 4 * LDB_S + j * 8];
          bf[j] = SQB ? v * v : v;
        }
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    produced();
    slot = slot + 1 == STAGES ? 0 : slot + 1;

    if (++kc == nkc) {
      // ---- epilogue of N tile `nt` ----
      if (EPI == EPI_STORE) {
        const double *sb = Sb + (((nt - nt0) % STAGES) * MAXS) * LDB_S + wn * 32 + 2 * q;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
          const int orow = orow_i[i];
          if (orow >= 0) {
            const size_t col = (size_t)nt * BN + wn * 32 + 2 * q;
            double *crow = C + (size_t)orow * ldc + col;
            if (scale) {
              const double *srow = nslot > 0 ? sb + slot_i[i] * LDB_S
                                             : scale + (size_t)(orow / scale_div) * lds + col;
#pragma unroll
              for (int j = 0; j < NF; ++j) {
                const double2 sc = *reinterpret_cast<const double2 *>(srow + j * 8);
                *reinterpret_cast<double2 *>(crow + j * 8) =
                    make_double2(acc[i][j][0] * sc.x, acc[i][j][1] * sc.y);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NF; ++j)
                *reinterpret_cast<double2 *>(crow + j * 8) =
                    make_double2(acc[i][j][0], acc[i][j][1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
          for (int j = 0; j < NF; ++j)
            rsq[i] += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
      }
#pragma unroll
      for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      kc = 0;
      ++nt;
    }
  }

  if (EPI == EPI_ROWSUMSQ) {
    cp_async_wait<0>();
    __syncthreads();
    double *red = smem;  // [4 wn][TM rows]
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      double v = rsq[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (q == 0) red[wn * TM + wm * MF * 8 + i * 8 + g] = v;
    }
    __syncthreads();
    if (tid < TM) {
      double v = red[tid] + red[TM + tid] + red[2 * TM + tid] + red[3 * TM + tid];
      rowsq[(size_t)split * M_pad + (size_t)mtile * TM + tid] = v;
    }
  }
}

template <int EPI, bool SQB, int MF>
int launch_variant(plsb_ctx *h, const GemmArgs &a, int n_ntiles, int n_splits, cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  constexpr int TM = 2 * MF * 8;
  const int smem = smem_bytes(MF, EPI == EPI_STORE && a.scale != nullptr);
  const int n_mtiles = a.M_pad / TM;
  const int nt_per_split = (n_ntiles + n_splits - 1) / n_splits;
  dim3 grid((unsigned)(n_mtiles * n_splits));
  if (tune_int("PLSB_GEMM_V", 2) == 1) {
    auto kern = xcov_gemm_kernel<EPI, SQB, MF>;
    PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, NTHREADS, smem, st>>>(a.A, a.lda, a.X, a.ldx, n_mtiles, n_ntiles, nt_per_split,
                                       a.kranges, a.Kd, a.C, a.ldc, a.row_map, a.scale,
                                       a.scale_div, a.lds, a.rowsq, a.M_pad);
  } else {
    auto kern = xcov_gemm2_kernel<EPI, SQB, MF>;
    PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, NTHREADS, smem, st>>>(a.A, a.lda, a.X, a.ldx, n_mtiles, n_ntiles, nt_per_split,
                                       a.kranges, a.Kd, a.k_valid > 0 ? a.k_valid : a.Kd, a.C,
                                       a.ldc, a.row_map, a.scale, a.scale_div, a.lds, a.rowsq,
                                       a.M_pad);
  }
  PLSB_LAUNCHED(h);
  h->dmma_flops += 2.0 * a.M_pad * a.N_pad * (a.k_len > 0 ? a.k_len : (a.k_valid > 0 ? a.k_valid : a.Kd));
  return PLSB_OK;
}

}  // namespace

// 64-row tiles with two CTAs per SM (one CTA's barrier / refill bubbles are covered by the
// other's DMMAs) unless the environment says otherwise
bool gemm_small_tile(int klen) {
  (void)klen;
  return tune_int("PLSB_GEMM_SMALL_TILE", 1) != 0;
}

// Splits of the N range per M tile.  A launch is `n_mtiles x n_splits` equal CTAs that run
// in waves of (SMs x CTAs per SM): the split count is chosen to minimise the modelled
// makespan  waves x (N tiles per CTA + pipeline fill)  -- a last wave that is nearly empty
// costs as much as a full one (measured: 6.2 waves ran as 7, -12 % on the bootstrap launch
// of config 5), so many short CTAs beat few long ones.
int gemm_pick_splits(const plsb_ctx *h, int M_pad, int n_ntiles, bool small_tile) {
  const int forced = tune_int("PLSB_GEMM_SPLITS", 0);
  const int n_mtiles = M_pad / (small_tile ? 64 : BM);
  const long long slots = (long long)h->sm_count * (small_tile ? 2 : 1);
  int best = 1;
  double best_cost = 1e300;
  const int max_splits = std::min(n_ntiles, 256);
  for (int want = 1; want <= max_splits; ++want) {
    const int per = (n_ntiles + want - 1) / want;       // N tiles per CTA
    const int ns = (n_ntiles + per - 1) / per;          // every split owns >= 1 tile
    if (ns != want) continue;
    const long long ctas = (long long)n_mtiles * ns;
    const long long waves = (ctas + slots - 1) / slots;
    // + 0.25 tile per CTA for the pipeline fill / drain and the launch of its first loads
    const double cost = (double)waves * (per + 0.25);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ns;
    }
  }
  if (forced > 0) {
    const int per = (n_ntiles + std::min(forced, n_ntiles) - 1) / std::min(forced, n_ntiles);
    return (n_ntiles + per - 1) / per;
  }
  return best;
}

int gemm_rowsq_splits(const plsb_ctx *h, const GemmArgs &a) {
  const int n_ntiles = a.N_pad / BN;
  if (gemm_i8_applies(h, a)) return gemm_i8_pick_splits(h, a.M_pad, n_ntiles);
  return gemm_pick_splits(h, a.M_pad, n_ntiles, gemm_small_tile(a.k_len > 0 ? a.k_len : a.Kd));
}

int launch_gemm(plsb_ctx *h, const GemmArgs &a, cudaStream_t st) {
  PLSB_CHECK(a.M_pad % BM == 0 && a.N_pad % BN == 0 && a.Kd % BK == 0, PLSB_ERR_ARG,
             "gemm: M_pad=%d N_pad=%d Kd=%d must be multiples of %d/%d/%d", a.M_pad, a.N_pad, a.Kd,
             BM, BN, BK);
  PLSB_CHECK(a.lda % 2 == 0 && a.ldx % 2 == 0, PLSB_ERR_ARG, "gemm: odd leading dimension");
  if (a.M_pad == 0 || a.N_pad == 0) return PLSB_OK;
  const int n_ntiles = a.N_pad / BN;
  const bool rowsq = a.rowsq != nullptr;
  if (rowsq) PLSB_CHECK(a.n_splits >= 1, PLSB_ERR_ARG, "gemm: n_splits");
  else PLSB_CHECK(a.C != nullptr && a.ldc % 2 == 0, PLSB_ERR_ARG, "gemm: bad output");
  if (gemm_i8_applies(h, a)) return launch_gemm_i8(h, a, st);
  if (rowsq) {
    if (a.square_b) return launch_variant<EPI_ROWSUMSQ, true, 8>(h, a, n_ntiles, a.n_splits, st);
    if (gemm_small_tile(a.k_len > 0 ? a.k_len : a.Kd))
      return launch_variant<EPI_ROWSUMSQ, false, 4>(h, a, n_ntiles, a.n_splits, st);
    return launch_variant<EPI_ROWSUMSQ, false, 8>(h, a, n_ntiles, a.n_splits, st);
  }
  PLSB_CHECK(a.C != nullptr && a.ldc % 2 == 0, PLSB_ERR_ARG, "gemm: bad output");
  // short contractions (block-diagonal operands: one cell's rows) are bound by the
  // store epilogue: 64-row tiles, two CTAs per SM
  const int klen = a.k_len > 0 ? a.k_len : a.Kd;
  const bool small_tile = gemm_small_tile(klen);
  if (small_tile) {
    const int n_splits = gemm_pick_splits(h, a.M_pad, n_ntiles, true);
    if (a.square_b) return launch_variant<EPI_STORE, true, 4>(h, a, n_ntiles, n_splits, st);
    return launch_variant<EPI_STORE, false, 4>(h, a, n_ntiles, n_splits, st);
  }
  const int n_splits = gemm_pick_splits(h, a.M_pad, n_ntiles, false);
  if (a.square_b) return launch_variant<EPI_STORE, true, 8>(h, a, n_ntiles, n_splits, st);
  return launch_variant<EPI_STORE, false, 8>(h, a, n_ntiles, n_splits, st);
}

}  // namespace plsb
