// Streaming passes over the stored cross-covariances for analyses with FEW latent
// variables (K = L <= 16; BASELINE config 5 has K = 10).
//
// The tensor-core kernels of gram_proj.cu / accum_u.cu pad K to whole 8-row DMMA
// fragments: at K = 10 they execute 896 and 512 flop per column and resample where the
// algorithm needs 310 and 220, which made both passes tensor bound at 2.6 - 3.4 TB/s of R
// traffic (31 + 24 ms per step of config 5).  FP64 FMA and FP64 DMMA have the same peak on
// sm_100a, so for small K the plain FMA formulation does the un-padded work and the passes
// become what they should be: HBM streams of R.
//
//   gram_proj_small<K>   G[r] = R[r] R[r]^T,  H[r] = R[r] U_orig      (one CTA per resample)
//   accum_u_small<K,L>   u_sum += sum_r R[r]^T M[r],  u_square += sum_r (R[r]^T M[r])^2
//
// Data path: a producer thread moves whole row segments (2 KB) of R -- and the matching
// rows of U_orig^T, the rotations M[r] -- into a ring of shared-memory stages with TMA bulk
// copies (cp.async.bulk.shared::cluster.global, completion counted in bytes on the stage's
// "full" mbarrier); the compute warps release a stage through its "empty" mbarrier.  There
// is no CTA-wide barrier in the loops and the copy engine, not the warps, generates the
// addresses.  Every thread owns columns (features) and keeps its outputs in registers.
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
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
// TMA bulk copy global -> shared (1-D); bytes and both addresses multiples of 16
__device__ __forceinline__ void tma_load(void *smem, const void *gmem, unsigned bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(smem)),
      "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void group_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// gram_proj_small
// ---------------------------------------------------------------------------------------
constexpr int GP_GROUP = 96;          // threads per compute group (2 columns each per stage)
constexpr int GP_GROUPS = 3;          // G | H[:, :L/2] | H[:, L/2:]
constexpr int GP_MAX_SLOTS = 8;
constexpr size_t GP_SMEM_BUDGET = 218 * 1024;   // dynamic shared memory; the rest: barriers, reductions
constexpr int GP_THREADS = GP_GROUPS * GP_GROUP + 32;

// sum over the threads of one compute group of n per-thread values; thread i < n of the
// group returns with the total of value i in *out (others: untouched)
template <int N>
__device__ __forceinline__ void group_reduce(double (&acc)[N], double *red, int gtid, int group,
                                             double *out) {
  const int gw = gtid >> 5, lane = gtid & 31;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) red[gw * N + i] = v;
  }
  group_sync(1 + group, GP_GROUP);
  if (gtid < N) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < GP_GROUP / 32; ++w) t += red[w * N + gtid];
    *out = t;
  }
}

template <int K, bool PROJ, int GP_CH>
__global__ void __launch_bounds__(GP_THREADS, 1)
gram_proj_small_kernel(const double *__restrict__ R, long long ldr, const double *__restrict__ UoT,
                       long long uot_stride, int uot_div, double *__restrict__ G,
                       double *__restrict__ H, int n_slots) {
  constexpr int L = K;
  constexpr int ROWS = PROJ ? K + L : K;           // rows of a stage: R, then U_orig^T
  constexpr int STAGE = ROWS * GP_CH;              // doubles
  constexpr int NG = K * (K + 1) / 2;              // upper triangle of G
  constexpr int LH0 = L / 2, LH1 = L - LH0;        // columns of H per group
  constexpr int NRED = NG > K * LH1 ? NG : K * LH1;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t full_bar[GP_MAX_SLOTS], empty_bar[GP_MAX_SLOTS];
  __shared__ double red[GP_GROUPS][(GP_GROUP / 32) * NRED];
  const int tid = threadIdx.x, r = blockIdx.x;
  const int n_stages = (int)((ldr + GP_CH - 1) / GP_CH);
  // PROJ: every stage is consumed by all three groups; G only: by group (stage % 3)
  const int consumers = PROJ ? GP_GROUPS * (GP_GROUP / 32) : GP_GROUP / 32;
  if (tid == 0) {
    for (int s = 0; s < n_slots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], consumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const double *Rr = R + (size_t)r * K * ldr;

  if (tid >= GP_GROUPS * GP_GROUP) {
    // ---- producer warp: lane 0 owns the barriers, every lane issues the bulk copy of its
    // own row (a single thread issuing all ~20 copies of a stage is issue bound) ----
    const int lane = tid & 31;
    const double *Pr = PROJ ? UoT + (size_t)(r / uot_div) * uot_stride : nullptr;
    for (int st = 0; st < n_stages; ++st) {
      const int slot = st % n_slots;
      const long long c0 = (long long)st * GP_CH;
      const unsigned bytes = (unsigned)(min((long long)GP_CH, ldr - c0) * sizeof(double));
      if (lane == 0) {
        if (st >= n_slots) mbar_wait(&empty_bar[slot], ((st / n_slots) - 1) & 1);
        mbar_expect_tx(&full_bar[slot], bytes * ROWS);
      }
      __syncwarp();
      double *dst = sm + (size_t)slot * STAGE;
      for (int row = lane; row < ROWS; row += 32) {
        const double *src = row < K ? Rr + (size_t)row * ldr + c0
                                    : Pr + (size_t)(row - K) * ldr + c0;
        tma_load(dst + row * GP_CH, src, bytes, &full_bar[slot]);
      }
    }
    return;
  }

  // ---- compute groups ----
  const int group = tid / GP_GROUP, gtid = tid - group * GP_GROUP;
  const int lane = tid & 31;
  if (group == 0 || !PROJ) {
    double acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0.0;
    for (int st = PROJ ? 0 : group; st < n_stages; st += PROJ ? 1 : GP_GROUPS) {
      const int slot = st % n_slots;
      mbar_wait(&full_bar[slot], (st / n_slots) & 1);
      const int cols = (int)min((long long)GP_CH, ldr - (long long)st * GP_CH);
      const double *src = sm + (size_t)slot * STAGE;
#pragma unroll
      for (int j = 0; j < GP_CH / GP_GROUP; ++j) {
        const int c = gtid + j * GP_GROUP;
        if (c < cols) {
          double x[K];
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = src[k * GP_CH + c];
          int o = 0;
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = a; b < K; ++b) acc[o] = fma(x[a], x[b], acc[o]), ++o;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
    if (PROJ) {
      double total = 0.0;
      group_reduce<NG>(acc, red[0], gtid, 0, &total);
      if (gtid < NG) {
        // entry gtid of the upper triangle -> (a, b)
        int a = 0, o = gtid;
        while (o >= K - a) o -= K - a, ++a;
        const int b = a + o;
        double *Gr = G + (size_t)r * K * K;
        Gr[a * K + b] = total;
        Gr[b * K + a] = total;
      }
    } else {
      // the three groups hold partial sums over disjoint stages: group 0 adds them up
      double total = 0.0;
      group_reduce<NG>(acc, red[group], gtid, group, &total);
      __shared__ double part[GP_GROUPS][NG];
      if (gtid < NG) part[group][gtid] = total;
      group_sync(4, GP_GROUPS * GP_GROUP);
      if (group == 0 && gtid < NG) {
        total = part[0][gtid] + part[1][gtid] + part[2][gtid];
        int a = 0, o = gtid;
        while (o >= K - a) o -= K - a, ++a;
        const int b = a + o;
        double *Gr = G + (size_t)r * K * K;
        Gr[a * K + b] = total;
        Gr[b * K + a] = total;
      }
    }
  } else {
    // H[:, l0 : l0 + nl] = sum_b R[:, b] * UoT[l, b]
    const int l0 = group == 1 ? 0 : LH0;
    constexpr int NLMAX = LH1;
    const int nl = group == 1 ? LH0 : LH1;
    double acc[K * NLMAX];
#pragma unroll
    for (int i = 0; i < K * NLMAX; ++i) acc[i] = 0.0;
    for (int st = 0; st < n_stages; ++st) {
      const int slot = st % n_slots;
      mbar_wait(&full_bar[slot], (st / n_slots) & 1);
      const int cols = (int)min((long long)GP_CH, ldr - (long long)st * GP_CH);
      const double *src = sm + (size_t)slot * STAGE;
#pragma unroll
      for (int j = 0; j < GP_CH / GP_GROUP; ++j) {
        const int c = gtid + j * GP_GROUP;
        if (c < cols) {
          double x[K], u[NLMAX];
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = src[k * GP_CH + c];
#pragma unroll
          for (int l = 0; l < NLMAX; ++l)
            u[l] = l < nl ? src[(K + l0 + l) * GP_CH + c] : 0.0;
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int l = 0; l < NLMAX; ++l) acc[k * NLMAX + l] = fma(x[k], u[l], acc[k * NLMAX + l]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
    double total = 0.0;
    group_reduce<K * NLMAX>(acc, red[group], gtid, group, &total);
    if (gtid < K * NLMAX) {
      const int k = gtid / NLMAX, l = gtid - k * NLMAX;
      if (l < nl) H[(size_t)r * K * L + (size_t)k * L + l0 + l] = total;
    }
  }
}

// ---------------------------------------------------------------------------------------
// accum_u_small
// ---------------------------------------------------------------------------------------
constexpr int AS_COLS = 256;          // columns per CTA == compute threads
constexpr int AS_RS = 2;              // resamples per stage
constexpr int AS_SLOTS = 2;
constexpr int AS_THREADS = AS_COLS + 32;

template <int K, int L>
__global__ void __launch_bounds__(AS_THREADS, 2)
accum_u_small_kernel(const double *__restrict__ R, long long ldr, int count, int B,
                     const double *__restrict__ M, int per_split, double *__restrict__ Psum,
                     double *__restrict__ Psq) {
  constexpr int LDM = ((L + 7) / 8) * 8 + 4;            // accum_ldm(L): pitch of M in HBM
  constexpr int PER_RES = K * AS_COLS + K * LDM;        // doubles per resample in a stage
  constexpr int STAGE = AS_RS * PER_RES;
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t full_bar[AS_SLOTS], empty_bar[AS_SLOTS];
  const int tid = threadIdx.x;
  const long long b0 = (long long)blockIdx.x * AS_COLS;
  const int split = blockIdx.y;
  const int r_beg = split * per_split, r_end = min(count, r_beg + per_split);
  const int n_stages = (r_end - r_beg + AS_RS - 1) / AS_RS;
  const int cols = (int)min((long long)AS_COLS, ldr - b0);
  if (tid == 0) {
    for (int s = 0; s < AS_SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], AS_COLS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (tid >= AS_COLS) {
    // producer warp: lane 0 owns the barriers, lanes issue one row copy each
    const int lane = tid & 31;
    const unsigned row_bytes = (unsigned)(cols * sizeof(double));
    constexpr unsigned m_bytes = K * LDM * sizeof(double);
    for (int st = 0; st < n_stages; ++st) {
      const int slot = st % AS_SLOTS;
      const int r0 = r_beg + st * AS_RS, n_here = min(AS_RS, r_end - r0);
      if (lane == 0) {
        if (st >= AS_SLOTS) mbar_wait(&empty_bar[slot], ((st / AS_SLOTS) - 1) & 1);
        mbar_expect_tx(&full_bar[slot], (unsigned)n_here * (K * row_bytes + m_bytes));
      }
      __syncwarp();
      // copies of the stage: n_here x (K rows of R, then the rotation M)
      for (int c = lane; c < n_here * (K + 1); c += 32) {
        const int t = c / (K + 1), k = c - t * (K + 1);
        double *dst = sm + (size_t)slot * STAGE + (size_t)t * PER_RES;
        if (k < K)
          tma_load(dst + k * AS_COLS, R + ((size_t)(r0 + t) * K + k) * ldr + b0, row_bytes,
                   &full_bar[slot]);
        else
          tma_load(dst + K * AS_COLS, M + (size_t)(r0 + t) * K * LDM, m_bytes, &full_bar[slot]);
      }
    }
    return;
  }

  double s1[L], s2[L];
#pragma unroll
  for (int l = 0; l < L; ++l) s1[l] = s2[l] = 0.0;
  for (int st = 0; st < n_stages; ++st) {
    const int slot = st % AS_SLOTS;
    mbar_wait(&full_bar[slot], (st / AS_SLOTS) & 1);
    const int n_here = min(AS_RS, r_end - (r_beg + st * AS_RS));
    for (int t = 0; t < n_here; ++t) {
      const double *Rs = sm + (size_t)slot * STAGE + (size_t)t * PER_RES;
      const double *Ms = Rs + K * AS_COLS;
      double x[K], u[L];
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = Rs[k * AS_COLS + tid];
#pragma unroll
      for (int l = 0; l < L; ++l) u[l] = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int l = 0; l < L; ++l) u[l] = fma(x[k], Ms[k * LDM + l], u[l]);   // broadcast loads
#pragma unroll
      for (int l = 0; l < L; ++l) {
        s1[l] += u[l];
        s2[l] = fma(u[l], u[l], s2[l]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty_bar[slot]);
  }
  const long long b = b0 + tid;
  if (tid < cols && b < B) {
    double *ps = Psum + ((size_t)split * B + b) * L, *pq = Psq + ((size_t)split * B + b) * L;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      ps[l] = s1[l];
      pq[l] = s2[l];
    }
  }
}

template <int K, int CH>
int launch_gp_small(plsb_ctx *h, bool proj, const double *R, long long ldr, int count,
                    const double *UoT, double *G, double *H, cudaStream_t st,
                    long long uot_stride, int uot_div) {
  constexpr int GP_CH = CH;
  const size_t stage = sizeof(double) * (size_t)(proj ? 2 * K : K) * GP_CH;
  // G only: stage st is consumed by group st % 3 alone, so the slot count must be a multiple
  // of 3 -- every slot then belongs to one group and that group sees each of its phases
  int n_slots = (int)std::min<size_t>(GP_MAX_SLOTS, GP_SMEM_BUDGET / stage);
  if (!proj) n_slots = n_slots / 3 * 3;
  PLSB_CHECK(n_slots >= 3, PLSB_ERR_ARG, "gram_proj_small: stage of %zu bytes", stage);
  const size_t smem = stage * n_slots;
  if (proj) {
    auto kern = gram_proj_small_kernel<K, true, CH>;
    PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<count, GP_THREADS, smem, st>>>(R, ldr, UoT, uot_stride, std::max(uot_div, 1), G, H,
                                          n_slots);
  } else {
    auto kern = gram_proj_small_kernel<K, false, CH>;
    PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<count, GP_THREADS, smem, st>>>(R, ldr, nullptr, 0, 1, G, nullptr, n_slots);
  }
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

template <int K>
int launch_au_small(plsb_ctx *h, const double *R, long long ldr, int count, int B, const double *M,
                    double *usum, double *usq, cudaStream_t st) {
  constexpr int L = K;
  constexpr int LDM = ((L + 7) / 8) * 8 + 4;
  const size_t smem = sizeof(double) * AS_SLOTS * AS_RS * (size_t)(K * AS_COLS + K * LDM);
  const int n_ct = (int)((ldr + AS_COLS - 1) / AS_COLS);
  // splits of the resamples: enough CTAs for ~4 waves of two CTAs per SM
  int n_splits = std::max(1, std::min(count / AS_RS, cdiv(8 * h->sm_count, n_ct)));
  const int per_split = round_up(cdiv(count, n_splits), AS_RS);
  n_splits = cdiv(count, per_split);
  const size_t stride = (size_t)B * L;
  PLSB_TRY(h->part.ensure(sizeof(double) * 2 * stride * n_splits));
  double *Psum = h->part.as<double>(), *Psq = Psum + stride * n_splits;
  auto kern = accum_u_small_kernel<K, L>;
  PLSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_ct, n_splits);
  kern<<<grid, AS_THREADS, smem, st>>>(R, ldr, count, B, M, per_split, Psum, Psq);
  PLSB_LAUNCHED(h);
  PLSB_TRY(launch_reduce_partials(h, Psum, n_splits, stride, stride, usum, st));
  PLSB_TRY(launch_reduce_partials(h, Psq, n_splits, stride, stride, usq, st));
  return PLSB_OK;
}

}  // namespace

#define PLSB_SMALL_K_SWITCH(K, CALL)                                                         \
  switch (K) {                                                                               \
    case 1: return CALL(1);  case 2: return CALL(2);  case 3: return CALL(3);                \
    case 4: return CALL(4);  case 5: return CALL(5);  case 6: return CALL(6);                \
    case 7: return CALL(7);  case 8: return CALL(8);  case 9: return CALL(9);                \
    case 10: return CALL(10); case 11: return CALL(11); case 12: return CALL(12);            \
    default: break;                                                                          \
  }

bool small_k_applies(int K, int L, bool proj) {
  return K >= 1 && K <= SMALL_K_MAX && (!proj || L == K) && tune_int("PLSB_SMALL_K", 1) != 0;
}

int launch_gram_proj_small(plsb_ctx *h, const double *R, long long ldr, int count, int K,
                           const double *UoT, int L, double *G, double *H, cudaStream_t st,
                           long long uot_stride, int uot_div) {
  const bool proj = UoT && H;
  PLSB_CHECK(small_k_applies(K, L, proj) && ldr % 2 == 0, PLSB_ERR_ARG,
             "gram_proj_small: K=%d L=%d not supported", K, L);
  // 3 KB row segments (384 columns) keep the copy engine ahead of the FMA groups; 1.5 KB
  // ones do not (measured at K = 10: 20.5 vs 30.5 ms per step of config 5); the narrow
  // stage is for K where three wide ones do not fit
  const bool wide = 3 * sizeof(double) * (size_t)(proj ? 2 * K : K) * 384 <= GP_SMEM_BUDGET;
  if (tune_int("PLSB_GP_CH", wide ? 384 : 192) == 384) {
#define PLSB_GP_CALL(KK) \
  launch_gp_small<KK, 384>(h, proj, R, ldr, count, UoT, G, H, st, uot_stride, uot_div)
    PLSB_SMALL_K_SWITCH(K, PLSB_GP_CALL)
#undef PLSB_GP_CALL
  }
#define PLSB_GP_CALL(KK) \
  launch_gp_small<KK, 192>(h, proj, R, ldr, count, UoT, G, H, st, uot_stride, uot_div)
  PLSB_SMALL_K_SWITCH(K, PLSB_GP_CALL)
#undef PLSB_GP_CALL
  return PLSB_ERR_ARG;
}

int launch_accum_u_small(plsb_ctx *h, const double *R, long long ldr, int count, int K, int B,
                         const double *M, int L, double *usum, double *usq, cudaStream_t st) {
  PLSB_CHECK(small_k_applies(K, L, true) && ldr % 2 == 0, PLSB_ERR_ARG,
             "accum_u_small: K=%d L=%d not supported", K, L);
#define PLSB_AU_CALL(KK) launch_au_small<KK>(h, R, ldr, count, B, M, usum, usq, st)
  PLSB_SMALL_K_SWITCH(K, PLSB_AU_CALL)
#undef PLSB_AU_CALL
  return PLSB_ERR_ARG;
}

}  // namespace plsb
