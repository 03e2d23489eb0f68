// Stand-alone micro-benchmark (not part of the library): per-SM throughput of the instructions
// the epilogue of the int8 slice GEMM is made of (DADD, DFMA, IMAD.WIDE, LOP3, SHFL), 16 warps.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rate_probe pipe_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) rate(long long *cycles, double *sink, int iters, double seed) {
  double a[8];
  long long w[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x + i;
    w[i] = threadIdx.x + i;
  }
  int x = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = a[i] + seed;                     // DADD
      if (OP == 1) a[i] = fma(a[i], seed, seed);           // DFMA
      if (OP == 2) w[i] = (long long)(int)w[(i + 1) & 7] * 256 + w[i];   // IMAD.WIDE
      if (OP == 6) w[i] = w[i] + (long long)(int)w[(i + 1) & 7];         // 64-bit add of a sign-extended int
      if (OP == 3) a[i] = __hiloint2double(0x43300000, __double2loint(a[i]) ^ 0x80000000);  // LOP3
      if (OP == 4) a[i] = __shfl_sync(0xffffffffu, a[i], (i + it) & 31);                    // 2 SHFL
      if (OP == 5) w[i] = __shfl_sync(0xffffffffu, (int)w[i], (i + it) & 31);               // 1 SHFL
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + (double)w[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int grid) {
  long long *dc;
  double *ds;
  cudaMalloc(&dc, sizeof(long long) * grid);
  cudaMalloc(&ds, sizeof(double) * grid * 512);
  const int iters = 2000;
  rate<OP><<<grid, 512>>>(dc, ds, iters, 1.000001);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, dc, sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%-10s grid %3d: %.1f thread-ops / clock / SM\n", name, grid, 512.0 * 8 * iters / c);
  cudaFree(dc);
  cudaFree(ds);
}

int main() {
  run<0>("DADD", 1);
  run<1>("DFMA", 1);
  run<1>("DFMA", 148);
  run<2>("IMAD.WIDE", 1);
  run<6>("IADD64", 1);
  run<3>("LOP3", 1);
  run<4>("SHFL x2", 1);
  run<5>("SHFL", 1);
  return 0;
}
