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
// Kernel layout (one persistent CTA per SM, 20 warps):
//   warp 0       producer: cp.async.bulk (TMA bulk copies) of pre-tiled digit planes; the planes
//                of 128 rows of one operand stay resident in shared memory (S x 28 KB), the
//                planes of the other stream through a ring of 16 KB stages
//   warp 1       one elected lane issues tcgen05.mma (128 x 128 x 32 per instruction); owns TMEM
//   warps 4-19   epilogue: tcgen05.ld, FP64 combination, stores / row sums of squares
// Tensor memory holds four 128 x 128 int32 accumulators, fewer than the S diagonals, so a
// tile is evaluated in two passes (diagonals S-4 .. S-1, then 0 .. S-5) over a ring of
// accumulator slots; the streamed planes come lowest digit first so that a pass touches its
// accumulators one after the other and the next pass starts while the epilogue still
// drains the previous one.
// Operand planes are stored in global memory exactly as the tensor core reads them from
// shared memory (no-swizzle K-major core matrices: [k step of 32][8-row group][2][8 rows]
// [16 B]), so a stage is one contiguous bulk copy and no tensor map is needed.
#include "common.cuh"

#include <type_traits>

namespace plsb {

namespace {

constexpr int TM = 128, TN = 128;
constexpr int MAX_KS = 7;                 // k steps of 32 -> contraction length <= 224
constexpr int KSTEP_BYTES = 128 * 32;     // one k step of one 128-row plane
constexpr int NSLOT = 4;                  // 128-column accumulators in tensor memory
constexpr int NSTAGE = 3;
constexpr int NACC = 4;                   // "pass complete" barriers in flight
constexpr int N_EPI_WARPS = 16;
// warps 0-3: control warpgroup (producer, MMA issuer, two idle), warps 4-19: epilogue
constexpr int THREADS = (4 + N_EPI_WARPS) * 32;

enum { EPI_STORE = 0, EPI_ROWSUMSQ = 1 };

// NP0: diagonals of the first pass (the second pass takes the remaining S - NP0 <= NSLOT)
template <int S, int NP0> struct Plan {
  static constexpr int NPASS = S > NP0 ? 2 : 1;
  static constexpr int STAGE_KS = S >= 7 ? 2 : 4;   // k steps per ring stage
  static constexpr int STAGE_BYTES = STAGE_KS * KSTEP_BYTES;
  static constexpr int A_BYTES = S * MAX_KS * KSTEP_BYTES;
  static constexpr int RING_OFF = A_BYTES;
  static constexpr int BAR_OFF = RING_OFF + NSTAGE * STAGE_BYTES;
  static constexpr int RED_OFF = BAR_OFF + 128;
  static constexpr int SMEM = RED_OFF + 4 * 128 * 8;
  __host__ __device__ static constexpr int dhi(int p) { return p == 0 ? S - 1 : S - NP0 - 1; }
  __host__ __device__ static constexpr int dlo(int p) { return p == 0 ? (S > NP0 ? S - NP0 : 0) : 0; }
};

struct TcParams {
  const int8_t *res_img, *str_img;   // planes of the resident / the streamed operand
  const int *str_nplanes;            // optional: leading planes of the streamed operand that are not all
                                     // zero (multiplicity operands are exact in one): the rest is skipped
  const double *rscale, *cscale;     // powers of two of the rows of A / the columns of X
  int KS, n_rtiles, n_stiles, n_splits, st_per_split;
  double *C;
  long long ldc;
  const double *scale;
  int scale_div, scale_rows;
  uint32_t scale_div_magic;          // ceil(2^32 / scale_div): x / scale_div == umulhi(x, magic) for small x
  long long lds;
  double *rowsq;
  int M_pad;
  long long *prof;   // optional per-CTA wait-cycle counters (PLSB_I8_PROF), 8 per CTA
};

// mbarrier wait that adds its duration to `acc` when profiling
#define TIMED_WAIT(bar, par, acc)            \
  do {                                       \
    if (p.prof) {                            \
      const long long _t = clock64();        \
      mbar_wait(bar, par);                   \
      acc += clock64() - _t;                 \
    } else {                                 \
      mbar_wait(bar, par);                   \
    }                                        \
  } while (0)

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
// Shared-memory matrix descriptors: no swizzle, K-major; core matrices adjacent in K 128 B
// apart (LBO), 8-row groups 256 B apart (SBO).
// a_lo / b_lo: low words of the descriptors; the high word (SBO 256 B, version 1) is constant
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                       uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x4010u)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(pred));
  return pred != 0;
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
}
__device__ __forceinline__ void tmem_wait_ld() {
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
                                                         double *__restrict__ rscale,
                                                         int *__restrict__ n_planes) {
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
  int top = 0;   // planes up to the last one with a non-zero digit (small integers need one)
#pragma unroll
  for (int i = 0; i < S; ++i) {
    *reinterpret_cast<uint4 *>(dst + (size_t)i * slice_bytes) =
        make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
    if (pk[i][0] | pk[i][1] | pk[i][2] | pk[i][3]) top = i + 1;
  }
  top = __reduce_max_sync(__activemask(), top);
  if (lane == 0 && top > 0) atomicMax(n_planes, top);
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
// Per unit of work one operand's planes stay resident in shared memory (128 of its rows /
// columns: the M side of the MMAs = the 128 TMEM lanes, so an epilogue thread owns one
// resident row and 32 streamed ones per tile), the other operand's tiles stream past:
//   ROWSUMSQ: resident = 128 rows of A, streamed = column tiles of X; a thread owns one row of
//             C and sums its squares over the whole streamed column range in one register;
//   STORE:    resident = 128 columns of X, streamed = row tiles of A (C^T = X^T A^T); a thread
//             owns one COLUMN of C, so the 32 lanes of a warp write 32 consecutive doubles of
//             a row of C (and read the column scales the same way).
// The MMAs of a product (its k steps) are issued back to back into the same accumulator.
template <int S, int EPI, int NP0>
__global__ void __launch_bounds__(THREADS, 1) xcov_gemm_i8_kernel(const TcParams p) {
  using P = Plan<S, NP0>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *Rs = smem;                        // resident planes [S][KS][128 x 32 B]
  uint8_t *Ss = smem + P::RING_OFF;          // ring of streamed stages
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + P::BAR_OFF);
  uint64_t *full_b = bars, *empty_b = bars + NSTAGE;
  uint64_t *r_full = bars + 2 * NSTAGE, *r_empty = r_full + 1;
  uint64_t *acc_full = r_empty + 1;          // [NACC]
  uint64_t *slot_empty = acc_full + NACC;    // [NSLOT]
  double *red = reinterpret_cast<double *>(smem + P::RED_OFF);   // [4][128]
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS = p.KS;
  const uint32_t slice_bytes = (uint32_t)KS * KSTEP_BYTES;
  const int n_units = p.n_rtiles * p.n_splits;
  constexpr int NCHUNK = (MAX_KS + P::STAGE_KS - 1) / P::STAGE_KS;
  // streamed planes jn .. S-1 hold zeros only: never loaded, never multiplied
  const int jn = p.str_nplanes ? min(max(__ldg(p.str_nplanes), 1), S) : S;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int s = 0; s < NACC; ++s) mbar_init(&acc_full[s], 1);
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

  // Producer and issuer warps run their loops with all 32 lanes (every value is
  // warp-uniform, so it lives in uniform registers) and one elected lane issues.
  if (warp == 0) {
    // ===== producer =====
    const bool leader = elect_one();
    uint32_t st = 0, st_phase = 0, r_phase = 0;
    long long w_empty = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int split = u / p.n_rtiles, rt = u - split * p.n_rtiles;
      const int t0 = split * p.st_per_split, t1 = min(t0 + p.st_per_split, p.n_stiles);
      mbar_wait(r_empty, r_phase ^ 1);
      if (leader) {
        mbar_expect_tx(r_full, S * slice_bytes);
        const int8_t *r_src = p.res_img + (size_t)rt * S * slice_bytes;
#pragma unroll
        for (int i = 0; i < S; ++i)
          bulk_g2s(Rs + (size_t)i * slice_bytes, r_src + (size_t)i * slice_bytes, slice_bytes, r_full);
      }
      r_phase ^= 1;
      for (int t = t0; t < t1; ++t) {
        const int8_t *s_tile = p.str_img + (size_t)t * S * slice_bytes;
#pragma unroll
        for (int ps = 0; ps < P::NPASS; ++ps) {
#pragma unroll
          for (int j = P::dhi(ps); j >= 0; --j) {
            if (j >= jn) continue;
#pragma unroll
            for (int kc = 0; kc < NCHUNK; ++kc) {
              const int k0 = kc * P::STAGE_KS;
              if (k0 < KS) {
                const uint32_t bytes = (uint32_t)min(P::STAGE_KS, KS - k0) * KSTEP_BYTES;
                TIMED_WAIT(&empty_b[st], st_phase ^ 1, w_empty);
                if (leader) {
                  mbar_expect_tx(&full_b[st], bytes);
                  bulk_g2s(Ss + (size_t)st * P::STAGE_BYTES,
                           s_tile + (size_t)j * slice_bytes + (size_t)k0 * KSTEP_BYTES, bytes,
                           &full_b[st]);
                }
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
    if (p.prof && leader) p.prof[blockIdx.x * 8 + 4] = w_empty;
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = elect_one();
    // instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3, M >> 4
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) |
                               ((uint32_t)(TM >> 4) << 24);
    // low words of the shared-memory descriptors (address >> 4 | LBO 128 B); a k step is 256 units on
    const uint32_t r_lo0 = ((smem_u32(Rs) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
    const uint32_t s_lo0 = ((smem_u32(Ss) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
    const uint32_t slice16 = slice_bytes >> 4;
    uint32_t st = 0, st_phase = 0, r_phase = 0, slot_ctr = 0, pass_ctr = 0, units_done = 0;
    long long w_full = 0, w_slot = 0, w_a = 0;
    const long long t_begin = clock64();
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int split = u / p.n_rtiles;
      const int t0 = split * p.st_per_split, t1 = min(t0 + p.st_per_split, p.n_stiles);
      TIMED_WAIT(r_full, r_phase, w_a);
      tc_fence_after();
      r_phase ^= 1;
      for (int t = t0; t < t1; ++t) {
#pragma unroll
        for (int ps = 0; ps < P::NPASS; ++ps) {
#pragma unroll
          for (int j = P::dhi(ps); j >= 0; --j) {
            if (j >= jn) continue;
#pragma unroll
            for (int kc = 0; kc < NCHUNK; ++kc) {
              const int k0 = kc * P::STAGE_KS;
              if (k0 < KS) {
                TIMED_WAIT(&full_b[st], st_phase, w_full);
                tc_fence_after();
                const uint32_t s_lo = s_lo0 + st * (P::STAGE_BYTES >> 4);
#pragma unroll
                for (int d = (P::dlo(ps) > j ? P::dlo(ps) : j); d <= P::dhi(ps); ++d) {
                  const int i = d - j;
                  const uint32_t c = slot_ctr + (uint32_t)(P::dhi(ps) - d);
                  // diagonal d is touched for the first time by its highest streamed plane,
                  // min(d, jn - 1): its slot must have been drained by every epilogue warp
                  const bool first = kc == 0 && (i == 0 || j == jn - 1);
                  if (first) {
                    TIMED_WAIT(&slot_empty[c % NSLOT], ((c / NSLOT) & 1) ^ 1, w_slot);
                    tc_fence_after();
                  }
                  const uint32_t taddr = tmem + (c % NSLOT) * TN;
                  const uint32_t r_lo = r_lo0 + (uint32_t)i * slice16 + (uint32_t)k0 * (KSTEP_BYTES >> 4);
#pragma unroll
                  for (int k = 0; k < P::STAGE_KS; ++k) {
                    if (k0 + k < KS && leader)
                      mma_i8(taddr, r_lo + k * (KSTEP_BYTES >> 4), s_lo + k * (KSTEP_BYTES >> 4), idesc,
                             (first && k == 0) ? 0u : 1u);
                  }
                }
                if (leader) mma_commit(&empty_b[st]);   // stage free once these MMAs have read it
                if (++st == NSTAGE) {
                  st = 0;
                  st_phase ^= 1;
                }
              }
            }
          }
          if (leader) mma_commit(&acc_full[pass_ctr % NACC]);
          ++pass_ctr;
          slot_ctr += (uint32_t)(P::dhi(ps) - P::dlo(ps) + 1);
        }
      }
      if (leader) mma_commit(r_empty);                  // resident planes may be replaced
      ++units_done;
    }
    // every asynchronous arrive has landed before the CTA retires
    if (units_done) mbar_wait(r_empty, (units_done - 1) & 1);
    if (p.prof && leader) {
      p.prof[blockIdx.x * 8 + 0] = clock64() - t_begin;
      p.prof[blockIdx.x * 8 + 1] = w_full;
      p.prof[blockIdx.x * 8 + 2] = w_slot;
      p.prof[blockIdx.x * 8 + 3] = w_a;
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = one TMEM lane (resident row) x 32 accumulator columns =====
    const int quarter = warp & 3;             // TMEM lanes a warp may read: 32 * (warp id % 4)
    const int colq = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colq * 32);
    uint32_t slot_ctr = 0, pass_ctr = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int split = u / p.n_rtiles, rt = u - split * p.n_rtiles;
      const int t0 = split * p.st_per_split, t1 = min(t0 + p.st_per_split, p.n_stiles);
      const size_t r_idx = (size_t)rt * TM + row;      // resident index of this thread
      // ROWSUMSQ: resident = rows of A; STORE: resident = columns of X
      const double r_scale = __ldg((EPI == EPI_ROWSUMSQ ? p.rscale : p.cscale) + r_idx);
      double rsq = 0.0;
      for (int t = t0; t < t1; ++t) {
        double T[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) T[e] = 0.0;
        // the 32 streamed indices of this thread: s0 + e; lane e fetches the power of two of
        // index s0 + e, handed round by shuffles when the tile is finished
        const size_t s0 = (size_t)t * TN + colq * 32;
        // (a power of two, zero or NaN: the low word is zero, one 32-bit shuffle moves it)
        const int s_hi = __double2hiint(__ldg((EPI == EPI_ROWSUMSQ ? p.cscale : p.rscale) + s0 + lane));
        auto drain_pass = [&](auto ps_c) {
          constexpr int ps = decltype(ps_c)::value;
          mbar_wait(&acc_full[pass_ctr % NACC], (pass_ctr / NACC) & 1);
          tc_fence_after();
          constexpr int ND = P::dhi(ps) - P::dlo(ps) + 1;
          // (not unrolled over the diagonals: the instruction footprint of the three roles
          // has to stay in the instruction cache)
#pragma unroll 1
          for (int dd = 0; dd < ND; ++dd) {
            const uint32_t slot = (slot_ctr + (uint32_t)dd) % NSLOT;
            // int32 -> double through the exponent trick, already weighted by 256^(S-1-d):
            // bits (0x433 + sh) << 52 | (v ^ 2^31)  ==  (2^52 + 2^31 + v) * 2^sh
            const int hi = 0x43300000 + ((8 * (S - 1 - P::dhi(ps)) + 8 * dd) << 20);
            const double magic = __hiloint2double(hi, (int)0x80000000);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16];
              tmem_ld16(t_lane + slot * TN + hh * 16, v);
              tmem_wait_ld();
              if (hh == 1) {
                // the slot is in registers: hand it back to the MMA issuer
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
          slot_ctr += (uint32_t)ND;
          ++pass_ctr;
        };
        drain_pass(std::integral_constant<int, 0>{});
        if constexpr (P::NPASS > 1) drain_pass(std::integral_constant<int, 1>{});
        // ---- finish the tile ----  (branch-free: the shuffles need a converged warp)
        if (EPI == EPI_ROWSUMSQ) {
          double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const double val = T[e] * __hiloint2double(__shfl_sync(0xffffffffu, s_hi, e), 0);
            part[e & 3] = fma(val, val, part[e & 3]);
          }
          rsq += (part[0] + part[1]) + (part[2] + part[3]);
        } else {
          // thread = column r_idx of C, T[e] = row s0 + e: a warp writes 32 consecutive doubles
#pragma unroll
          for (int e = 0; e < 32; ++e)
            T[e] *= r_scale * __hiloint2double(__shfl_sync(0xffffffffu, s_hi, e), 0);
          double *cptr = p.C + s0 * (size_t)p.ldc + r_idx;
          if (p.scale) {
            // row s0 + e takes scale row (s0 + e) / scale_div; the quotient by multiply-shift,
            // exact for the values it sees.  Rows beyond the table (padding) reuse its last
            // row.  (Loading only the <= 5 distinct rows and selecting, a one-ahead software
            // prefetch and an L2 prefetch at the start of the tile were all measured: not faster.)
            const uint32_t q0 = (uint32_t)(s0 / (size_t)p.scale_div);
            const uint32_t rem0 = (uint32_t)(s0 - (size_t)q0 * p.scale_div);
            const uint32_t kmax = (uint32_t)max(p.scale_rows - 1 - (int)q0, 0);
            const double *sptr = p.scale + (size_t)min(q0, (uint32_t)(p.scale_rows - 1)) * p.lds + r_idx;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const uint32_t x = rem0 + (uint32_t)e;
              const uint32_t k = min(p.scale_div == 1 ? x : __umulhi(x, p.scale_div_magic), kmax);
              __stcs(cptr + (size_t)e * p.ldc, T[e] * __ldg(sptr + (size_t)k * p.lds));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) __stcs(cptr + (size_t)e * p.ldc, T[e]);
          }
        }
      }
      if (EPI == EPI_ROWSUMSQ) {
        // the four column quarters of a row meet in shared memory
        red[colq * TM + row] = rsq;
        asm volatile("bar.sync 1, %0;\n" ::"n"(N_EPI_WARPS * 32) : "memory");
        if (colq == 0) {
          const double v = red[row] + red[TM + row] + red[2 * TM + row] + red[3 * TM + row];
          p.rowsq[(size_t)split * p.M_pad + r_idx] = v * r_scale * r_scale;
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
                                   double *rscale, int *n_planes, cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  PLSB_CUDA(cudaMemsetAsync(n_planes, 0, sizeof(int), st));
  quant_rows_kernel<S><<<(unsigned)((rows_pad + 7) / 8), 256, 0, st>>>(
      A, lda, rows_valid, rows_pad, k_valid, KS, img, rscale, n_planes);
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

template <int S, int EPI, int NP0> int launch_kernel(plsb_ctx *h, const TcParams &p, cudaStream_t st) {
  KernelTimer kt(h, KC_GEMM, st);
  auto kern = xcov_gemm_i8_kernel<S, EPI, NP0>;
  constexpr int SMEM = Plan<S, NP0>::SMEM;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const int units = p.n_rtiles * p.n_splits;
  const int grid = std::min(units, h->sm_count);
  if (!tune_int("PLSB_I8_PROF", 0)) {
    kern<<<grid, THREADS, SMEM, st>>>(p);
    PLSB_LAUNCHED(h);
    return PLSB_OK;
  }
  // profiling aid: where the roles of the kernel wait (cycles, mean over the CTAs)
  TcParams q = p;
  PLSB_CUDA(cudaMalloc(&q.prof, sizeof(long long) * 8 * grid));
  PLSB_CUDA(cudaMemsetAsync(q.prof, 0, sizeof(long long) * 8 * grid, st));
  kern<<<grid, THREADS, SMEM, st>>>(q);
  PLSB_LAUNCHED(h);
  std::vector<long long> hp(8 * (size_t)grid);
  PLSB_CUDA(cudaMemcpyAsync(hp.data(), q.prof, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost, st));
  PLSB_CUDA(cudaStreamSynchronize(st));
  cudaFree(q.prof);
  double m[8] = {0};
  for (int c = 0; c < grid; ++c)
    for (int k = 0; k < 8; ++k) m[k] += (double)hp[8 * c + k] / grid;
  const double tiles = (double)p.n_rtiles * p.n_stiles / grid;
  fprintf(stderr,
          "[i8 prof] S=%d epi=%d grid=%d tiles/CTA=%.0f: cycles/tile %.0f | mma waits: stage %.0f "
          "slot %.0f resident %.0f | producer waits empty %.0f\n",
          S, EPI, grid, tiles, m[0] / tiles, m[1] / tiles, m[2] / tiles, m[3] / tiles, m[4] / tiles);
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
  PLSB_TRY(h->ascale.ensure(sizeof(double) * ((size_t)a.M_pad + 1)));   // + the plane count
  int *a_nplanes = reinterpret_cast<int *>(h->ascale.as<double>() + a.M_pad);
  PLSB_TRY(quantize_rows<S>(h, a.A, a.lda, a.M_pad, a.M_pad, k_valid, KS, h->aplanes.as<int8_t>(),
                            h->ascale.as<double>(), a_nplanes, st));
  TcParams p;
  p.str_nplanes = nullptr;
  p.rscale = h->ascale.as<double>();
  p.cscale = xc->scale.as<double>();
  p.KS = KS;
  if (a.rowsq) {
    // rows of A resident, column tiles of X streamed
    p.res_img = h->aplanes.as<int8_t>();
    p.str_img = xc->img.as<int8_t>();
    p.n_rtiles = n_mtiles;
    p.n_stiles = n_ntiles;
    p.n_splits = a.n_splits;
  } else {
    // columns of X resident, row tiles of A streamed
    p.res_img = xc->img.as<int8_t>();
    p.str_img = h->aplanes.as<int8_t>();
    p.str_nplanes = a_nplanes;
    p.n_rtiles = n_ntiles;
    p.n_stiles = n_mtiles;
    p.n_splits = gemm_i8_pick_splits(h, a.N_pad, n_mtiles);
  }
  p.st_per_split = (p.n_stiles + p.n_splits - 1) / p.n_splits;
  p.C = a.C;
  p.ldc = a.ldc;
  p.scale = a.scale;
  p.scale_div = a.scale_div;
  p.scale_rows = a.scale_rows > 0 ? a.scale_rows : (a.M_pad + a.scale_div - 1) / a.scale_div;
  p.scale_div_magic = a.scale_div > 1 ? (uint32_t)(((1ull << 32) + a.scale_div - 1) / a.scale_div) : 0u;
  p.lds = a.lds;
  p.rowsq = a.rowsq;
  p.M_pad = a.M_pad;
  p.prof = nullptr;
  // plane products executed: all pairs i + j < S -- or, for the multiplicity operands of
  // the column statistics (one non-empty plane, the kernel skips the rest), S of them
  h->i8_macs += (double)a.M_pad * a.N_pad * (KS * 32.0) *
                (a.a_counts && !a.rowsq ? S : S * (S + 1) / 2);
  // (a 3 + 3 split of the diagonals, Plan<S, 3>, and passes of two were measured: not faster)
  if (a.rowsq) return launch_kernel<S, EPI_ROWSUMSQ, 4>(h, p, st);
  return launch_kernel<S, EPI_STORE, 4>(h, p, st);
}

}  // namespace

// Splits of the streamed range per resident tile for the persistent kernel: units (resident
// tile, split) are dealt round-robin to one CTA per SM; a unit costs its streamed tiles plus
// ~0.6 tile for loading the resident planes and filling / draining the pipeline.
// (res_pad: padded extent of the resident operand, n_ntiles: tiles of the streamed one.)
int gemm_i8_pick_splits(const plsb_ctx *h, int res_pad, int n_ntiles) {
  const int n_mtiles = res_pad / TM;
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
  if (a.kranges || a.row_map) return false;
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
