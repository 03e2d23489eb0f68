// Stand-alone feasibility probe (not part of the library): one tcgen05.mma.kind::i8 tile
//   D (128 x N, int32, TMEM) = A (128 x K, int8, K-major) * B (N x K, int8, K-major)^T
// with operands staged by plain stores into the canonical no-swizzle K-major layout
// (8 x 16 B core matrices), checked against the CPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tcgen05_i8_probe tcgen05_i8_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 32, K = 224;          // K bytes per row, multiple of 32
constexpr int KCH = K / 16;                      // 16-byte chunks along K
constexpr int LBO = 128;                         // bytes between core matrices adjacent in K
constexpr int SBO = KCH * 128;                   // bytes between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);              // start address
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;          // leading byte offset
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;          // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  return d;                                            // layout type 0: no swizzle
}

__global__ void __launch_bounds__(128) probe(const int8_t *A, const int8_t *B, int32_t *D) {
  __shared__ __align__(128) uint8_t As[M * K];
  __shared__ __align__(128) uint8_t Bs[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage operands: 16-byte chunk (r, c) -> (r % 8) * 16 + (r / 8) * SBO + c * LBO
  for (int e = tid; e < M * KCH; e += 128) {
    const int r = e / KCH, c = e % KCH;
    *reinterpret_cast<int4 *>(As + (r % 8) * 16 + (r / 8) * SBO + c * LBO) =
        *reinterpret_cast<const int4 *>(A + (size_t)r * K + c * 16);
  }
  for (int e = tid; e < N * KCH; e += 128) {
    const int r = e / KCH, c = e % KCH;
    *reinterpret_cast<int4 *>(Bs + (r % 8) * 16 + (r / 8) * SBO + c * LBO) =
        *reinterpret_cast<const int4 *>(B + (size_t)r * K + c * 16);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                     "r"(smem_u32(&tmem_base)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // generic-proxy stores -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // instruction descriptor: c = S32 (2), a = b = S8 (1), K-major both, N >> 3, M >> 4
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    for (int k = 0; k < K / 32; ++k) {
      const uint64_t da = make_desc(smem_u32(As) + k * 2 * LBO);
      const uint64_t db = make_desc(smem_u32(Bs) + k * 2 * LBO);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                     "r"(smem_u32(&bar))
                 : "memory");
  }
  // wait for the MMAs
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::
          "r"(smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // warp w reads TMEM lanes 32w .. 32w+31: thread = row, 8 columns per load
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(32));
}

int main() {
  std::vector<int8_t> A(M * K), B(N * K);
  srand(1);
  for (auto &v : A) v = (int8_t)(rand() % 201 - 100);
  for (auto &v : B) v = (int8_t)(rand() % 129 - 64);
  int8_t *dA, *dB;
  int32_t *dD;
  cudaMalloc(&dA, A.size());
  cudaMalloc(&dB, B.size());
  cudaMalloc(&dD, sizeof(int32_t) * M * N);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, sizeof(int32_t) * M * N);
  probe<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<int32_t> D(M * N);
  cudaMemcpy(D.data(), dD, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)A[m * K + k] * (int32_t)B[n * K + k];
      if (ref != D[m * N + n]) {
        if (bad < 8) printf("mismatch (%d,%d): got %d want %d\n", m, n, D[m * N + n], ref);
        ++bad;
      }
    }
  printf("mismatches: %ld of %d\n", bad, M * N);
  return bad != 0;
}
