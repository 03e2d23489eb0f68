// accum_u: the FP64 tensor-core (DMMA m8n8k4) streaming contraction that turns
// the stored cross-covariance matrices R (count blocks of K rows x ldr) and the
// per-resample rotations M into the bootstrap sums (pyls/base.py:510-511 with
// compute.procrustes, pyls/compute.py:260-262, folded into M):
//
//   U_r = R_r^T M_r (B x L) for every resample r of a split,
//   u_sum += U_r, u_square += U_r^2 kept in registers; one CTA per
//   (column tile of R, split of the resamples), cp.async ring over the
//   resamples, partial sums reduced afterwards (stream_kernels.cu).
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
