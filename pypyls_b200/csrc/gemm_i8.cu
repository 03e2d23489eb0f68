// FP64 cross-covariance GEMM on the int8 tcgen05 tensor cores of sm_100a.
//
//   C (M, N) = A (M, Kd) @ X (Kd, N)          (compute.xcorr, pyls/compute.py:92, `Yn.T @ Xn`)
//
// FP64 has no tcgen05 kind and DMMA tops out at ~35 TFLOP/s.  The int8 kind runs at
// 8192 MAC / clock / SM (measured: scripts/probes/tc_rate_probe.cu), so the product is
// evaluated in integer slices (Ozaki splitting): every row of A and every column of X is
// scaled by a power of two to a fixed-point integer of 8*S - 2 bits and cut into S signed
// base-256 digits (int8).  Digit planes i of A and j of X are multiplied exactly by
// tcgen05.mma.kind::i8 (int32 accumulators in tensor memory, |sum| < 2^25); all pairs with
// i + j = d share the weight 256^(2S-2-d) and accumulate into the same tensor-memory
// accumulator, pairs with i + j >= S are below the fixed-point resolution and dropped.
// The epilogue warps read the S diagonal accumulators, combine them in FP64 (exact up to
// 2^53), apply the row / column powers of two and finish the tile (store with optional
// column scales, or row sums of squares).  With S = 6 (21 integer products) the result
// differs from the FP64 product by ~1e-13 of |a||x| per entry; S = 7 gives ~1e-15.
//
// Kernel layout (one persistent CTA per SM, 18 warps):
//   warp 0      producer: cp.async.bulk (TMA bulk copies) of pre-tiled digit planes; the planes
//               of the CTA's 128 rows of A stay resident in shared memory (S x 28 KB), the
//               planes of X stream through a ring of 16 KB stages
//   warp 1      one thread issues tcgen05.mma (128 x 128 x 32 per instruction); owns TMEM
//   warps 2-17  epilogue: tcgen05.ld, FP64 combination, stores
// Tensor memory holds four 128 x 128 int32 accumulators, fewer than the S diagonals, so a
// tile is evaluated in two passes (diagonals S-4 .. S-1, then 0 .. S-5) over a ring of
// accumulator slots; the planes of X are streamed from the lowest digit up so that a pass
// touches its accumulators one after the other and the next pass starts while the
// epilogue still drains the previous one.
// Operand planes are stored in global memory exactly as the tensor core reads them from
// shared memory (no-swizzle K-major core matrices: [k step of 32][8-row group][2][8 rows]
// [16 B]), so a stage is one contiguous bulk copy and no tensor map is needed.
#include "common.cuh"

namespace plsb {

namespace {

constexpr int TM = 128, TN = 128;
constexpr int MAX_KS = 7;                 // k steps of 32 -> contraction length <= 224
constexpr int KSTEP_BYTES = 128 * 32;     // one k step of one 128-row plane
constexpr int NSLOT = 4;                  // 128-column accumulators in tensor memory
constexpr int NSTAGE = 3;
constexpr int N_EPI_WARPS = 16;
constexpr int THREADS = (2 + N_EPI_WARPS) * 32;

enum { EPI_STORE = 0, EPI_ROWSUMSQ = 1 };

template <int S> struct Plan {
  static constexpr int NPASS = S > NSLOT ? 2 : 1;
  static constexpr int STAGE_KS = S >= 7 ? 2 : 4;   // k steps per ring stage
  static constexpr int STAGE_BYTES = STAGE_KS * KSTEP_BYTES;
  static constexpr int A_BYTES = S * MAX_KS * KSTEP_BYTES;
  static constexpr int RING_OFF = A_BYTES;
  static constexpr int BAR_OFF = RING_OFF + NSTAGE * STAGE_BYTES;
  static constexpr int RED_OFF = BAR_OFF + 128;
  static constexpr int SMEM = RED_OFF + N_EPI_WARPS * 32 * 8;
  __host__ __device__ static constexpr int dlo(int p) { return p == 0 ? (S > NSLOT ? S - NSLOT : 0) : 0; }
  __host__ __device__ static constexpr int dhi(int p) { return p == 0 ? S - 1 : S - NSLOT - 1; }
};

struct TcParams {
  const int8_t *Aimg, *Ximg;
  const double *rscale, *cscale;
  int KS, n_mtiles, n_ntiles, n_splits, nt_per_split;
  double *C;
  long long ldc;
  const int *row_map;
  const double *scale;
  int scale_div;
  long long lds;
  double *rowsq;
  int M_pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// shared-memory matrix descriptor: no swizzle, K-major; core matrices adjacent in K 128 B
// apart, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) |
         ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc,
                                       uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                   "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- digit planes ---------------------------------------------------------------------
// value v with |v| < 2^e  ->  q = rint(v * 2^(Q - e)), Q = 8S - 2, as S balanced base-256
// digits (q = sum_i dig[i] * 256^(S-1-i), every digit in [-128, 127], |dig[0]| <= 65)
template <int S>
__device__ __forceinline__ void digits_into(double v, int sh, int t, uint32_t (&pk)[S][4]) {
  long long q = __double2ll_rn(scalbn(v, sh));
#pragma unroll
  for (int i = S - 1; i >= 1; --i) {
    const long long low = (long long)(int8_t)(q & 0xFF);
    pk[i][t >> 2] |= (uint32_t)(uint8_t)low << (8 * (t & 3));
    q = (q - low) >> 8;
  }
  pk[0][t >> 2] |= (uint32_t)(uint8_t)q << (8 * (t & 3));
}

// exponent e with max < 2^e (max > 0, finite)
__device__ __forceinline__ int exp_above(double mx) { return ilogb(mx) + 1; }

// One warp per row of A (row-major, k contiguous): lane c < 2*KS owns k = 16c .. 16c+15.
// Writes the S planes of the row into the tile images and the row's power of two.
template <int S>
__global__ void __launch_bounds__(256) quant_rows_kernel(const double *__restrict__ A, int lda,
                                                         long long rows_valid, long long rows_pad,
                                                         int k_valid, int KS, int8_t *__restrict__ img,
                                                         double *__restrict__ rscale) {
  constexpr int Q = 8 * S - 2;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows_pad) return;
  double v[16];
  double mx = 0.0;
  bool bad = false;
  const bool live = lane < 2 * KS && row < rows_valid;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int k = lane * 16 + t;
    v[t] = (live && k < k_valid) ? A[(size_t)row * lda + k] : 0.0;
    const double a = fabs(v[t]);
    bad |= !(a <= 1.79e308);
    mx = fmax(mx, a);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  bad = __any_sync(0xffffffffu, bad);
  int e = 0;
  if (mx > 0.0 && !bad) e = exp_above(mx);
  if (lane == 0)
    rscale[row] = bad ? __longlong_as_double(0x7ff8000000000000ll) : (mx > 0.0 ? scalbn(1.0, e - Q) : 0.0);
  if (lane >= 2 * KS) return;
  uint32_t pk[S][4];
#pragma unroll
  for (int i = 0; i < S; ++i) pk[i][0] = pk[i][1] = pk[i][2] = pk[i][3] = 0u;
#pragma unroll
  for (int t = 0; t < 16; ++t) digits_into<S>((mx > 0.0 && !bad) ? v[t] : 0.0, Q - e, t, pk);
  const long long tile = row / TM;
  const int r = (int)(row % TM);
  const size_t slice_bytes = (size_t)KS * KSTEP_BYTES;
  int8_t *dst = img + (size_t)tile * S * slice_bytes + (size_t)(lane >> 1) * KSTEP_BYTES +
                (size_t)(r >> 3) * 256 + (size_t)(lane & 1) * 128 + (size_t)(r & 7) * 16;
#pragma unroll
  for (int i = 0; i < S; ++i)
    *reinterpret_cast<uint4 *>(dst + (size_t)i * slice_bytes) =
        make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
}

// One CTA per 128 columns of X (row-major (k, ldx)): thread (x, y) owns column x of the tile
// and k step y (32 values).  cscale carries the column's power of two and the constant
// 256^(S-1) of the kept diagonals.
template <int S>
__global__ void __launch_bounds__(128 * MAX_KS) quant_cols_kernel(const double *__restrict__ X, int ldx,
                                                                  int k_valid, int KS, int square,
                                                                  int8_t *__restrict__ img,
                                                                  double *__restrict__ cscale) {
  constexpr int Q = 8 * S - 2;
  __shared__ double s_mx[MAX_KS][128];
  __shared__ int s_bad[128];
  const int x = threadIdx.x, y = threadIdx.y;
  const size_t col = (size_t)blockIdx.x * TN + x;
  if (y == 0) s_bad[x] = 0;
  __syncthreads();
  double mx = 0.0;
  bool bad = false;
  for (int t = 0; t < 32; ++t) {
    const int k = y * 32 + t;
    double a = k < k_valid ? fabs(X[(size_t)k * ldx + col]) : 0.0;
    if (square) a *= a;
    bad |= !(a <= 1.79e308);
    mx = fmax(mx, a);
  }
  s_mx[y][x] = mx;
  if (bad) s_bad[x] = 1;
  __syncthreads();
  mx = 0.0;
  for (int t = 0; t < KS; ++t) mx = fmax(mx, s_mx[t][x]);
  bad = s_bad[x] != 0;
  int e = 0;
  if (mx > 0.0 && !bad) e = exp_above(mx);
  if (y == 0)
    cscale[col] = bad ? __longlong_as_double(0x7ff8000000000000ll)
                      : (mx > 0.0 ? scalbn(1.0, e - Q + 8 * (S - 1)) : 0.0);
  const size_t slice_bytes = (size_t)KS * KSTEP_BYTES;
  int8_t *dst = img + (size_t)blockIdx.x * S * slice_bytes + (size_t)y * KSTEP_BYTES +
                (size_t)(x >> 3) * 256 + (size_t)(x & 7) * 16;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t pk[S][4];
#pragma unroll
    for (int i = 0; i < S; ++i) pk[i][0] = pk[i][1] = pk[i][2] = pk[i][3] = 0u;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int k = y * 32 + half * 16 + t;
      double v = (k < k_valid && mx > 0.0 && !bad) ? X[(size_t)k * ldx + col] : 0.0;
      if (square) v *= v;
      digits_into<S>(v, Q - e, t, pk);
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
      *reinterpret_cast<uint4 *>(dst + (size_t)i * slice_bytes + half * 128) =
          make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
  }
}

// ---- the GEMM ---------------------------------------------------------------------------
template <int S, int EPI>
__global__ void __launch_bounds__(THREADS, 1) xcov_gemm_i8_kernel(const TcParams p) {
  using P = Plan<S>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *As = smem;
  uint8_t *Bs = smem + P::RING_OFF;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + P::BAR_OFF);
  uint64_t *full_b = bars, *empty_b = bars + NSTAGE;
  uint64_t *a_full = bars + 2 * NSTAGE, *a_empty = a_full + 1;
  uint64_t *acc_full = a_empty + 1;          // [2]
  uint64_t *slot_empty = acc_full + 2;       // [NSLOT]
  double *red = reinterpret_cast<double *>(smem + P::RED_OFF);
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS = p.KS;
  const uint32_t slice_bytes = (uint32_t)KS * KSTEP_BYTES;
  const int n_units = p.n_mtiles * p.n_splits;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    for (int s = 0; s < NSLOT; ++s) mbar_init(&slot_empty[s], N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                     "r"(smem_u32(&s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t st = 0, st_phase = 0, a_phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int split = u / p.n_mtiles, mt = u - split * p.n_mtiles;
        const int nt0 = split * p.nt_per_split, nt1 = min(nt0 + p.nt_per_split, p.n_ntiles);
        mbar_wait(a_empty, a_phase ^ 1);
        mbar_expect_tx(a_full, S * slice_bytes);
        const int8_t *a_src = p.Aimg + (size_t)mt * S * slice_bytes;
#pragma unroll
        for (int i = 0; i < S; ++i)
          bulk_g2s(As + (size_t)i * slice_bytes, a_src + (size_t)i * slice_bytes, slice_bytes, a_full);
        a_phase ^= 1;
        for (int nt = nt0; nt < nt1; ++nt) {
          const int8_t *x_tile = p.Ximg + (size_t)nt * S * slice_bytes;
#pragma unroll
          for (int ps = 0; ps < P::NPASS; ++ps) {
            for (int j = P::dhi(ps); j >= 0; --j) {
              for (int k0 = 0; k0 < KS; k0 += P::STAGE_KS) {
                const uint32_t bytes = (uint32_t)min(P::STAGE_KS, KS - k0) * KSTEP_BYTES;
                mbar_wait(&empty_b[st], st_phase ^ 1);
                mbar_expect_tx(&full_b[st], bytes);
                bulk_g2s(Bs + (size_t)st * P::STAGE_BYTES,
                         x_tile + (size_t)j * slice_bytes + (size_t)k0 * KSTEP_BYTES, bytes,
                         &full_b[st]);
                if (++st == NSTAGE) {
                  st = 0;
                  st_phase ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3, M >> 4
      constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) |
                                 ((uint32_t)(TM >> 4) << 24);
      const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
      uint32_t st = 0, st_phase = 0, a_phase = 0, slot_ctr = 0, pass_ctr = 0, units_done = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int split = u / p.n_mtiles;
        const int nt0 = split * p.nt_per_split, nt1 = min(nt0 + p.nt_per_split, p.n_ntiles);
        mbar_wait(a_full, a_phase);
        tc_fence_after();
        a_phase ^= 1;
        for (int nt = nt0; nt < nt1; ++nt) {
#pragma unroll
          for (int ps = 0; ps < P::NPASS; ++ps) {
            const int dlo = P::dlo(ps), dhi = P::dhi(ps);
            for (int j = dhi; j >= 0; --j) {
              for (int k0 = 0; k0 < KS; k0 += P::STAGE_KS) {
                const int ks = min(P::STAGE_KS, KS - k0);
                mbar_wait(&full_b[st], st_phase);
                tc_fence_after();
                const uint32_t b_st = b_base + st * P::STAGE_BYTES;
                for (int d = max(dlo, j); d <= dhi; ++d) {
                  const int i = d - j;
                  const uint32_t c = slot_ctr + (uint32_t)(dhi - d);
                  const uint32_t slot = c % NSLOT;
                  const bool first = (i == 0 && k0 == 0);
                  if (first) {
                    // the accumulator slot must have been drained by every epilogue warp
                    mbar_wait(&slot_empty[slot], ((c / NSLOT) & 1) ^ 1);
                    tc_fence_after();
                  }
                  const uint32_t a_sl = a_base + (uint32_t)i * slice_bytes + (uint32_t)k0 * KSTEP_BYTES;
                  for (int k = 0; k < ks; ++k)
                    mma_i8(tmem + slot * TN, make_desc(a_sl + k * KSTEP_BYTES),
                           make_desc(b_st + k * KSTEP_BYTES), idesc, (first && k == 0) ? 0u : 1u);
                }
                mma_commit(&empty_b[st]);     // stage free once these MMAs have read it
                if (++st == NSTAGE) {
                  st = 0;
                  st_phase ^= 1;
                }
              }
            }
            mma_commit(&acc_full[pass_ctr & 1]);
            ++pass_ctr;
            slot_ctr += (uint32_t)(dhi - dlo + 1);
          }
        }
        mma_commit(a_empty);                  // resident planes of A may be replaced
        ++units_done;
      }
      // every asynchronous arrive has landed before the CTA retires
      if (units_done) mbar_wait(a_empty, (units_done - 1) & 1);
    }
  } else {
    // ===== epilogue: thread = one row of the tile x 32 columns =====
    const int quarter = warp & 3;             // TMEM lanes a warp may read: 32 * (warp id % 4)
    const int colq = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colq * 32);
    uint32_t slot_ctr = 0, pass_ctr = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int split = u / p.n_mtiles, mt = u - split * p.n_mtiles;
      const int nt0 = split * p.nt_per_split, nt1 = min(nt0 + p.nt_per_split, p.n_ntiles);
      const int m = mt * TM + row;
      const double rs = p.rscale[m];
      int orow = m;
      if (EPI == EPI_STORE && p.row_map) orow = p.row_map[m];
      double rsq = 0.0;
      for (int nt = nt0; nt < nt1; ++nt) {
        double T[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) T[e] = 0.0;
#pragma unroll
        for (int ps = 0; ps < P::NPASS; ++ps) {
          mbar_wait(&acc_full[pass_ctr & 1], (pass_ctr >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int d = P::dhi(ps); d >= P::dlo(ps); --d) {
            const uint32_t c = slot_ctr + (uint32_t)(P::dhi(ps) - d);
            const uint32_t slot = c % NSLOT;
            // int32 -> double through the exponent trick, already weighted by 256^(S-1-d):
            // bits (0x433 + sh) << 52 | (v ^ 2^31)  ==  (2^52 + 2^31 + v) * 2^sh
            const int sh = 8 * (S - 1 - d);
            const int hi = 0x43300000 + (sh << 20);
            const double magic = __hiloint2double(hi, (int)0x80000000);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16];
              tmem_ld16(t_lane + slot * TN + hh * 16, v);
              if (hh == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&slot_empty[slot]);
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const double xv = __hiloint2double(hi, (int)(v[e] ^ 0x80000000u));
                T[hh * 16 + e] += (xv - magic);
              }
            }
          }
          slot_ctr += (uint32_t)(P::dhi(ps) - P::dlo(ps) + 1);
          ++pass_ctr;
        }
        // ---- finish the tile ----
        const size_t col0 = (size_t)nt * TN + colq * 32;
        const double *cs = p.cscale + col0;
        if (EPI == EPI_ROWSUMSQ) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const double val = T[e] * __ldg(cs + e);
            rsq = fma(val, val, rsq);
          }
        } else if (orow >= 0) {
          double *crow = p.C + (size_t)orow * p.ldc + col0;
          if (p.scale) {
            const double *srow = p.scale + (size_t)(orow / p.scale_div) * p.lds + col0;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const double2 sc = __ldg(reinterpret_cast<const double2 *>(srow + e));
              const double2 cc = __ldg(reinterpret_cast<const double2 *>(cs + e));
              __stcs(reinterpret_cast<double2 *>(crow + e),
                     make_double2(T[e] * rs * cc.x * sc.x, T[e + 1] * rs * cc.y * sc.y));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const double2 cc = __ldg(reinterpret_cast<const double2 *>(cs + e));
              __stcs(reinterpret_cast<double2 *>(crow + e),
                     make_double2(T[e] * rs * cc.x, T[e + 1] * rs * cc.y));
            }
          }
        }
      }
      if (EPI == EPI_ROWSUMSQ) {
        // the four column quarters of a row meet in shared memory
        red[colq * TM + row] = rsq;
        asm volatile("bar.sync 1, %0;\n" ::"n"(N_EPI_WARPS * 32) : "memory");
        if (colq == 0) {
          const double v = red[row] + red[TM + row] + red[2 * TM + row] + red[3 * TM + row];
          p.rowsq[(size_t)split * p.M_pad + m] = v * rs * rs;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(N_EPI_WARPS * 32) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int S> int quantize_rows(plsb_ctx *h, const double *A, int lda, long long rows_valid,
                                   long long rows_pad, int k_valid, int KS, int8_t *img,
                                   double *rscale, cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  quant_rows_kernel<S><<<(unsigned)((rows_pad + 7) / 8), 256, 0, st>>>(A, lda, rows_valid, rows_pad,
                                                                        k_valid, KS, img, rscale);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

template <int S> int quantize_cols(plsb_ctx *h, const double *X, int ldx, int N_pad, int k_valid,
                                   int KS, bool square, int8_t *img, double *cscale,
                                   cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  quant_cols_kernel<S><<<N_pad / TN, dim3(128, KS), 0, st>>>(X, ldx, k_valid, KS, square ? 1 : 0, img,
                                                             cscale);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

template <int S, int EPI> int launch_kernel(plsb_ctx *h, const TcParams &p, cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  auto kern = xcov_gemm_i8_kernel<S, EPI>;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<S>::SMEM));
  const int units = p.n_mtiles * p.n_splits;
  kern<<<std::min(units, h->sm_count), THREADS, Plan<S>::SMEM, st>>>(p);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

template <int S> int run_i8(plsb_ctx *h, const GemmArgs &a, int KS, int k_valid, cudaStream_t st) {
  const size_t slice_bytes = (size_t)KS * KSTEP_BYTES;
  const int n_mtiles = a.M_pad / TM, n_ntiles = a.N_pad / TN;
  // planes of X: cached while the data matrix of the handle is unchanged
  plsb_ctx::PlaneCache *xc = nullptr;
  if (a.x_persistent) {
    for (auto &c : h->xplanes)
      if (c.src == a.X && c.epoch == h->data_epoch && c.S == S && c.ldx == a.ldx &&
          c.N_pad == a.N_pad && c.k_valid == k_valid && c.square == a.square_b)
        xc = &c;
    if (!xc) {
      // least recently filled entry
      xc = &h->xplanes[0];
      for (auto &c : h->xplanes)
        if (c.stamp < xc->stamp) xc = &c;
      xc->src = nullptr;
    }
  } else {
    xc = &h->xplanes_tmp;
    xc->src = nullptr;
  }
  if (xc->src == nullptr) {
    PLSB_TRY(xc->img.ensure((size_t)n_ntiles * S * slice_bytes));
    PLSB_TRY(xc->scale.ensure(sizeof(double) * (size_t)a.N_pad));
    PLSB_TRY(quantize_cols<S>(h, a.X, a.ldx, a.N_pad, k_valid, KS, a.square_b,
                              xc->img.as<int8_t>(), xc->scale.as<double>(), st));
    xc->src = a.X;
    xc->epoch = h->data_epoch;
    xc->S = S;
    xc->ldx = a.ldx;
    xc->N_pad = a.N_pad;
    xc->k_valid = k_valid;
    xc->square = a.square_b;
    xc->stamp = ++h->plane_stamp;
  }
  PLSB_TRY(h->aplanes.ensure((size_t)n_mtiles * S * slice_bytes));
  PLSB_TRY(h->ascale.ensure(sizeof(double) * (size_t)a.M_pad));
  PLSB_TRY(quantize_rows<S>(h, a.A, a.lda, a.M_pad, a.M_pad, k_valid, KS, h->aplanes.as<int8_t>(),
                            h->ascale.as<double>(), st));
  TcParams p;
  p.Aimg = h->aplanes.as<int8_t>();
  p.Ximg = xc->img.as<int8_t>();
  p.rscale = h->ascale.as<double>();
  p.cscale = xc->scale.as<double>();
  p.KS = KS;
  p.n_mtiles = n_mtiles;
  p.n_ntiles = n_ntiles;
  p.C = a.C;
  p.ldc = a.ldc;
  p.row_map = a.row_map;
  p.scale = a.scale;
  p.scale_div = a.scale_div;
  p.lds = a.lds;
  p.rowsq = a.rowsq;
  p.M_pad = a.M_pad;
  if (a.rowsq) {
    p.n_splits = a.n_splits;
  } else {
    p.n_splits = gemm_i8_pick_splits(h, a.M_pad, n_ntiles);
  }
  p.nt_per_split = (n_ntiles + p.n_splits - 1) / p.n_splits;
  if (a.rowsq) return launch_kernel<S, EPI_ROWSUMSQ>(h, p, st);
  return launch_kernel<S, EPI_STORE>(h, p, st);
}

}  // namespace

// Splits of the N range per M tile for the persistent kernel: units (M tile, split) are
// dealt round-robin to one CTA per SM; a unit costs its N tiles plus ~0.6 tile for loading
// the resident planes of A and filling / draining the pipeline.
int gemm_i8_pick_splits(const plsb_ctx *h, int M_pad, int n_ntiles) {
  const int n_mtiles = M_pad / TM;
  int best = 1;
  double best_cost = 1e300;
  for (int want = 1; want <= std::min(n_ntiles, 128); ++want) {
    const int per = (n_ntiles + want - 1) / want;
    const int ns = (n_ntiles + per - 1) / per;
    if (ns != want) continue;
    const long long units = (long long)n_mtiles * ns;
    const long long waves = (units + h->sm_count - 1) / h->sm_count;
    const double cost = (double)waves * (per + 0.6);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

// The slice GEMM takes dense operands (no per-tile contraction ranges) and contractions of
// at most 224 rows.
bool gemm_i8_applies(const plsb_ctx *h, const GemmArgs &a) {
  if (h->gemm_backend == PLSB_GEMM_DMMA) return false;
  if (tune_int("PLSB_GEMM_BACKEND", 1) == 0) return false;
  if (a.kranges) return false;
  const int k_valid = a.k_valid > 0 ? a.k_valid : a.Kd;
  if ((k_valid + 31) / 32 > MAX_KS) return false;
  if (a.M_pad % TM || a.N_pad % TN) return false;
  return true;
}

int gemm_i8_slices(const plsb_ctx *h) {
  int s = tune_int("PLSB_GEMM_SLICES", h->gemm_slices);
  return s >= 7 ? 7 : (s <= 5 ? 5 : 6);
}

int launch_gemm_i8(plsb_ctx *h, const GemmArgs &a, cudaStream_t st) {
  const int k_valid = a.k_valid > 0 ? a.k_valid : a.Kd;
  const int KS = (k_valid + 31) / 32;
  switch (gemm_i8_slices(h)) {
    case 5: return run_i8<5>(h, a, KS, k_valid, st);
    case 7: return run_i8<7>(h, a, KS, k_valid, st);
    default: return run_i8<6>(h, a, KS, k_valid, st);
  }
}

}  // namespace plsb
