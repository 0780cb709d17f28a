// Microbenchmark: MUFU.EX2 throughput per SM for f32, f16x2 and bf16x2 operands (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, long long* cyc) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0xbc00bc00u + threadIdx.x + i;   // small negative halves / arbitrary floats
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));          // F2FP.BF16.F32.PACK_AB
      if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i])); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(y[i]) : "r"(x[i])); }
      if (MODE == 5) asm volatile("max.f32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  uint32_t* d; long long* c;
  cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 6; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, c);
        if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, c);
        if (mode == 2) k<2><<<148, warps * 32>>>(d, iters, c);
        if (mode == 3) k<3><<<148, warps * 32>>>(d, iters, c);
        if (mode == 4) k<4><<<148, warps * 32>>>(d, iters, c);
        if (mode == 5) k<5><<<148, warps * 32>>>(d, iters, c);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double ops = (double)iters * 8 * warps * 32;   // MUFU thread-ops per SM
      printf("mode %d (%s) warps/SM %2d: %.2f thread-ops/clk/SM  (%.2f results/clk/SM)\n", mode, mode == 0 ? "f32" : mode == 1 ? "f16x2" : mode == 2 ? "bf16x2" : mode == 3 ? "F2FP pack" : mode == 4 ? "ex2+F2FP" : "FMNMX",
             warps, ops / h, ops / h * ((mode == 1 || mode == 2) ? 2 : 1));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
