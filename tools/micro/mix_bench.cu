// Microbenchmark: the attention exp-phase instruction mix per key pair (2 MUFU.EX2 + FFMA2 + FADD2 + F2FP) on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_bench mix_bench.cu && ./mix_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE bit0: MUFU, bit1: FFMA2, bit2: FADD2 (ACCS independent accumulators), bit3: F2FP
template <int MODE, int ACCS>
__global__ void k(uint32_t* out, int iters, long long* cyc) {
  uint32_t x[16];
  uint64_t acc[ACCS];
  uint32_t pk = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0xbf800000u + threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < ACCS; ++i) acc[i] = 0;
  const uint64_t m = 0x3f8000003f800000ull, c = 0xbf000000bf000000ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 2)
        asm volatile("{ .reg .b64 t; mov.b64 t, {%0, %1}; fma.rn.f32x2 t, t, %2, %3; mov.b64 {%0, %1}, t; }" : "+r"(x[2 * i]), "+r"(x[2 * i + 1]) : "l"(m), "l"(c));
      if (MODE & 1) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[2 * i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[2 * i + 1]));
      }
      if (MODE & 4)
        asm volatile("{ .reg .b64 t; mov.b64 t, {%1, %2}; add.rn.f32x2 %0, %0, t; }" : "+l"(acc[i % ACCS]) : "r"(x[2 * i]), "r"(x[2 * i + 1]));
      if (MODE & 8) {
        uint32_t d;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "r"(x[2 * i + 1]), "r"(x[2 * i]));
        pk ^= d;
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = pk;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
#pragma unroll
  for (int i = 0; i < ACCS; ++i) s ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ACCS>
void run(const char* name, uint32_t* d, long long* c) {
  const int iters = 1000;
  for (int warps = 4; warps <= 8; warps *= 2) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { k<MODE, ACCS><<<148, warps * 32>>>(d, iters, c); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM %d: %.1f clk per key pair per SMSP-warp-slot, %.2f keys/clk/SM\n", name, warps, (double)h / (iters * 8.0) / (warps / 4),
           (double)iters * 16 * warps * 32 / h);
  }
}

int main() {
  uint32_t* d; long long* c;
  cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
  run<1, 1>("MUFU only", d, c);
  run<2, 1>("FFMA2 only", d, c);
  run<4, 1>("FADD2 only (1 acc chain)", d, c);
  run<4, 4>("FADD2 only (4 acc chains)", d, c);
  run<8, 1>("F2FP only", d, c);
  run<14, 1>("FFMA2+FADD2(1)+F2FP", d, c);
  run<14, 4>("FFMA2+FADD2(4)+F2FP", d, c);
  run<15, 1>("all, 1 acc chain", d, c);
  run<15, 4>("all, 4 acc chains", d, c);
  run<9, 1>("MUFU+F2FP", d, c);
  run<3, 1>("MUFU+FFMA2", d, c);
  run<5, 4>("MUFU+FADD2(4)", d, c);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
