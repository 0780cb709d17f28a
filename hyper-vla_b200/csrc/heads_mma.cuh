// Generation heads on bf16 tensor cores: out[T, NGP] = E[T,128] * W[128, NGP] + b   (hypernetwork.py:205-217).
// The 73 Dense heads are ONE skinny GEMM whose cost is streaming W (51.6 MB bf16) from HBM exactly once and
// writing the per-task weight rows; each CTA owns a 128-column slab of W in shared memory and loops over the
// tasks 64 at a time (warp-level mma.sync m16n8k16; the problem is K=128 deep and HBM-bound, not a tcgen05 shape).
#pragma once
#include "common.cuh"
#include "attn_mma.cuh"

namespace hvla {
namespace heads {

using attn::ldsm_x4;
using attn::ldsm_x4_t;
using attn::mma_bf16;
using attn::pack2;

constexpr int NC = 128, TM = 64, LDS_ = 136;                 // slab columns, tasks per pass, padded smem row (bf16)
constexpr int SMEM = (128 * LDS_ + TM * LDS_ + TM * LDS_) * 2 + NC * 4;   // W slab | E chunk | out tile | bias

__global__ void __launch_bounds__(256, 2)
heads_mma_kernel(const float* __restrict__ E, const bf16* __restrict__ W, const float* __restrict__ bias, bf16* __restrict__ out, int T) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* Ws = reinterpret_cast<bf16*>(smem);
  bf16* Es = Ws + 128 * LDS_;
  bf16* Os = Es + TM * LDS_;
  float* bs = reinterpret_cast<float*>(Os + TM * LDS_);
  const int64_t col0 = (int64_t)blockIdx.x * NC;
  const int ncols = (int)((NGP - col0) < NC ? (NGP - col0) : NC);     // NGP % 32 == 0: the last slab is 32 wide
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // stage the W slab (128 rows x up to 128 columns) and the bias
  for (int i = threadIdx.x; i < 128 * 16; i += 256) {
    const int r = i >> 4, c = (i & 15) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c < ncols) v = __ldg(reinterpret_cast<const uint4*>(W + (int64_t)r * NGP + col0 + c));
    *reinterpret_cast<uint4*>(Ws + r * LDS_ + c) = v;
  }
  if (threadIdx.x < NC) bs[threadIdx.x] = threadIdx.x < ncols ? bias[col0 + threadIdx.x] : 0.f;
  const uint32_t sW = (uint32_t)__cvta_generic_to_shared(Ws), sE = (uint32_t)__cvta_generic_to_shared(Es);
  const int mt = warp & 3, nh = warp >> 2;                      // m-tile (16 tasks) and 64-column half of this warp
  for (int t0 = 0; t0 < T; t0 += TM) {
    __syncthreads();                                            // previous pass done with Es / Os (and W staged on pass 0)
    for (int i = threadIdx.x; i < TM * 32; i += 256) {          // E chunk: fp32 -> bf16, 4 values per thread
      const int r = i >> 5, c = (i & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + r < T) v = __ldg(reinterpret_cast<const float4*>(E + (int64_t)(t0 + r) * CD + c));
      *reinterpret_cast<uint2*>(Es + r * LDS_ + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t a[4];
      // A fragment (16 tasks x 16 k): matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
      ldsm_x4(sE + (uint32_t)(((mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS_ + ks * 16 + (lane >> 4) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int i = lane >> 3;
        ldsm_x4_t(sW + (uint32_t)(((ks * 16 + (i & 1) * 8 + (lane & 7)) * LDS_ + nh * 64 + (np * 2 + (i >> 1)) * 8) * 2), b0, b1, b2, b3);
        mma_bf16(acc[2 * np], a, b0, b1);
        mma_bf16(acc[2 * np + 1], a, b2, b3);
      }
    }
    // + bias -> bf16 -> smem tile
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = nh * 64 + j * 8 + (lane & 3) * 2;
      const int r = mt * 16 + (lane >> 2);
      *reinterpret_cast<uint32_t*>(Os + r * LDS_ + c) = pack2(acc[j][0] + bs[c], acc[j][1] + bs[c + 1]);
      *reinterpret_cast<uint32_t*>(Os + (r + 8) * LDS_ + c) = pack2(acc[j][2] + bs[c], acc[j][3] + bs[c + 1]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TM * 16; i += 256) {          // coalesced 16-byte stores, 256 B per task row
      const int r = i >> 4, c = (i & 15) * 8;
      if (t0 + r < T && c < ncols)
        *reinterpret_cast<uint4*>(out + (int64_t)(t0 + r) * NGP + col0 + c) = *reinterpret_cast<const uint4*>(Os + r * LDS_ + c);
    }
  }
}

inline int heads_gemm_bf16(cudaStream_t st, const float* E, const bf16* W, const float* bias, bf16* out, int T) {
  static bool attr = false;
  if (!attr) {
    HVLA_CUDA(cudaFuncSetAttribute(heads_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    HVLA_CUDA(cudaFuncSetAttribute(heads_mma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  ProfScope ps(st, "heads_gemm");
  heads_mma_kernel<<<cdiv(NGP, NC), 256, SMEM, st>>>(E, W, bias, out, T);
  HVLA_LAUNCH_CHECK("heads_mma");
  return HVLA_OK;
}

}  // namespace heads
}  // namespace hvla
