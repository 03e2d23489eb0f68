// Stand-alone micro-benchmarks (not part of the library) that size the int8 slice GEMM:
//   1. tcgen05.mma.kind::i8 issue rate, A and B from shared memory, M = 128, N = 64 .. 256,
//      operands in the no-swizzle K-major layout [k step][8-row group][2][8 rows][16 B]
//      (LBO = 128, SBO = 256), with a correctness check of that layout;
//   2. cp.async.bulk streaming rate global -> shared memory, all SMs reading the same /
//      disjoint data.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tc_rate_probe tc_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, K = 224, KSTEPS = K / 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128 >> 4) & 0x3FFF) << 16;   // LBO: core matrices adjacent in K
  d |= (uint64_t)((256 >> 4) & 0x3FFF) << 32;   // SBO: 8-row groups
  d |= (uint64_t)1 << 46;
  return d;
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

// image offset of byte k of row r in a [k step][row group][2][8][16] image of `rows` rows
__host__ __device__ inline size_t img_off(int rows, int r, int k) {
  return (size_t)(k / 32) * rows * 32 + (size_t)(r / 8) * 256 + (size_t)((k % 32) / 16) * 128 +
         (size_t)(r % 8) * 16 + (k % 16);
}

template <int N>
__global__ void __launch_bounds__(128) mma_rate(const int8_t *Aimg, const int8_t *Bimg, int iters,
                                                long long *cycles, int32_t *D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *As = smem, *Bs = smem + M * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < M * K / 16; e += 128)
    reinterpret_cast<int4 *>(As)[e] = reinterpret_cast<const int4 *>(Aimg)[e];
  for (int e = tid; e < N * K / 16; e += 128)
    reinterpret_cast<int4 *>(Bs)[e] = reinterpret_cast<const int4 *>(Bimg)[e];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::
                     "r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;
  constexpr int NSLOT = 512 / N;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t acc = tmem + (uint32_t)((it % NSLOT) * N);
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k)
        mma_i8(acc, make_desc(a0 + k * M * 32), make_desc(b0 + k * N * 32), idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (blockIdx.x == 0 && D) {
    // slot 0 holds the full product (the last pass over it started with accumulate = 0)
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
          "%12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, "
          "%30, %31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
            "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
            "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

// one thread per CTA streams `chunks` pieces of `bytes` through a ring of STAGES buffers
template <int STAGES>
__global__ void __launch_bounds__(32) bulk_rate(const uint8_t *src, size_t span, int bytes,
                                                int chunks, int disjoint, long long *cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  size_t off = disjoint ? ((size_t)blockIdx.x * (span / gridDim.x)) / 128 * 128 : 0;
  const size_t lim = disjoint ? off + (span / gridDim.x) / 128 * 128 : span;
  const size_t base = off;
  auto issue = [&](int c) {
    const int s = c % STAGES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full[s])),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(smem + (size_t)s * bytes)),
        "l"(src + off), "r"(bytes), "r"(smem_u32(&full[s]))
        : "memory");
    off += bytes;
    if (off + bytes > lim) off = base;
  };
  const long long t0 = clock64();
  for (int c = 0; c < STAGES && c < chunks; ++c) issue(c);
  for (int c = 0; c < chunks; ++c) {
    mbar_wait(&full[c % STAGES], (c / STAGES) & 1);
    if (c + STAGES < chunks) issue(c + STAGES);
  }
  cycles[blockIdx.x] = clock64() - t0;
}

template <int N>
static void run_mma(const std::vector<int8_t> &A, const std::vector<int8_t> &B, int grid, int iters,
                    bool check) {
  std::vector<int8_t> Ai(M * K), Bi(N * K);
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < K; ++k) Ai[img_off(M, r, k)] = A[r * K + k];
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < K; ++k) Bi[img_off(N, r, k)] = B[r * K + k];
  int8_t *dA, *dB;
  int32_t *dD;
  long long *dC;
  cudaMalloc(&dA, Ai.size());
  cudaMalloc(&dB, Bi.size());
  cudaMalloc(&dD, sizeof(int32_t) * M * N);
  cudaMalloc(&dC, sizeof(long long) * grid);
  cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice);
  const int smem = (M + N) * K;
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mma_rate<N><<<grid, 128, smem>>>(dA, dB, 64, dC, dD);   // warm-up
  cudaEventRecord(e0);
  mma_rate<N><<<grid, 128, smem>>>(dA, dB, iters, dC, dD);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), dC, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long cmax = 0;
  for (auto c : cyc) cmax = std::max(cmax, c);
  const double macs = (double)iters * M * N * K;
  printf("mma N=%3d grid=%3d: %s  %.1f cycles per %dx%dx%d product (floor %d), %.0f MAC/clk/SM, "
         "%.1f ms -> %.2f POP/s chip, %.0f MHz\n",
         N, grid, cudaGetErrorString(e), (double)cmax / iters, M, N, K, KSTEPS * N / 2,
         macs / cmax, ms, 2.0 * macs * grid / (ms * 1e-3) / 1e15, cmax / (ms * 1e-3) / 1e6);
  if (check) {
    std::vector<int32_t> D(M * N);
    cudaMemcpy(D.data(), dD, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int32_t ref = 0;
        for (int k = 0; k < K; ++k) ref += (int32_t)A[m * K + k] * (int32_t)B[n * K + k];
        if (ref != D[m * N + n]) {
          if (bad < 4) printf("  mismatch (%d,%d): got %d want %d\n", m, n, D[m * N + n], ref);
          ++bad;
        }
      }
    printf("  layout check N=%d: %ld mismatches of %d\n", N, bad, M * N);
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dC);
}

template <int STAGES>
static void run_bulk(const uint8_t *src, size_t span, int bytes, int grid, int disjoint) {
  long long *dC;
  cudaMalloc(&dC, sizeof(long long) * grid);
  const int smem = STAGES * bytes;
  cudaFuncSetAttribute(bulk_rate<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int chunks = 4000;
  bulk_rate<STAGES><<<grid, 32, smem>>>(src, span, bytes, 200, disjoint, dC);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bulk_rate<STAGES><<<grid, 32, smem>>>(src, span, bytes, chunks, disjoint, dC);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), dC, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long cmax = 0;
  for (auto c : cyc) cmax = std::max(cmax, c);
  printf("bulk %5d B x %d stages grid=%3d %s span %4zu MB: %s  %.1f B/clk/SM, %.2f TB/s chip\n", bytes,
         STAGES, grid, disjoint ? "disjoint" : "shared  ", span >> 20, cudaGetErrorString(e),
         (double)chunks * bytes / cmax, (double)chunks * bytes * grid / (ms * 1e-3) / 1e12);
  cudaFree(dC);
}

int main() {
  srand(7);
  std::vector<int8_t> A(M * K), B(256 * K);
  for (auto &v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto &v : B) v = (int8_t)(rand() % 256 - 128);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run_mma<64>(A, B, 1, 2000, true);
  run_mma<128>(A, B, 1, 2000, true);
  run_mma<256>(A, B, 1, 2000, true);
  run_mma<64>(A, B, sms, 20000, false);
  run_mma<96>(A, B, sms, 20000, false);
  run_mma<128>(A, B, sms, 20000, false);
  run_mma<192>(A, B, sms, 20000, false);
  run_mma<256>(A, B, sms, 20000, false);
  uint8_t *src;
  const size_t span = 160ull << 20;
  cudaMalloc(&src, span);
  cudaMemset(src, 1, span);
  run_bulk<4>(src, span, 14336, sms, 0);
  run_bulk<4>(src, span, 14336, sms, 1);
  run_bulk<2>(src, span, 28672, sms, 0);
  run_bulk<8>(src, span, 7168, sms, 0);
  run_bulk<4>(src, 32ull << 20, 14336, sms, 0);
  run_bulk<4>(src, span, 14336, 1, 0);
  return 0;
}
