// Generation heads on bf16 tensor cores: out[T, NGP] = E[T,128] * W[128, NGP] + b   (hypernetwork.py:205-217).
// The 73 Dense heads are ONE skinny GEMM whose cost is streaming W (51.6 MB bf16) from HBM exactly once and
// writing the per-task weight rows; persistent CTAs (two per SM) walk the 128-column slabs of W through a double-buffered
// cp.async ring and loop over the tasks 64 at a time (warp-level mma.sync m16n8k16; the problem is K=128 deep and
// HBM-bound, not a tcgen05 shape).
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
constexpr int WSLAB = 128 * LDS_ * 2;                        // one staged W slab (bytes)
constexpr int SMEM = 2 * WSLAB + (TM * LDS_ + TM * LDS_) * 2 + 2 * NC * 4;   // 2 x W slab | E chunk | out tile | 2 x bias

using attn::cp_async16;
using attn::cp_async_commit;
using attn::cp_async_wait;

// Persistent: a CTA walks the 128-column slabs of W with a double-buffered cp.async ring (the next slab streams
// from HBM while the current one is multiplied and its output tile is written), and keeps the tasks' context
// embeddings in shared memory across slabs when they fit one pass (T <= 64).
__global__ void __launch_bounds__(256, 2)
heads_mma_kernel(const float* __restrict__ E, const bf16* __restrict__ W, const float* __restrict__ bias, bf16* __restrict__ out, int T,
                 int nslabs, const int32_t* __restrict__ rows, int T_max, int64_t ngp) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* Es = reinterpret_cast<bf16*>(smem + 2 * WSLAB);
  bf16* Os = Es + TM * LDS_;
  float* bs = reinterpret_cast<float*>(Os + TM * LDS_);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sE = (uint32_t)__cvta_generic_to_shared(Es), sB = (uint32_t)__cvta_generic_to_shared(bs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, nh = warp >> 2;                      // m-tile (16 tasks) and 64-column half of this warp
  auto stage_w = [&](int slab, int buf) {
    if (slab < nslabs) {
      const int64_t col0 = (int64_t)slab * NC;
      const int ncols = (int)((ngp - col0) < NC ? (ngp - col0) : NC);   // ngp % 32 == 0: the last slab may be narrower
      for (int i = threadIdx.x; i < 128 * 16; i += 256) {
        const int r = i >> 4, c = (i & 15) * 8;
        if (c < ncols) cp_async16(sbase + (uint32_t)(buf * WSLAB + (r * LDS_ + c) * 2), W + (int64_t)r * ngp + col0 + c);
      }
      if (threadIdx.x < ncols / 4) cp_async16(sB + (uint32_t)((buf * NC + threadIdx.x * 4) * 4), bias + col0 + threadIdx.x * 4);
    }
    cp_async_commit();
  };
  auto stage_e = [&](int t0) {                                  // E chunk: fp32 -> bf16, 4 values per thread
    for (int i = threadIdx.x; i < TM * 32; i += 256) {
      const int r = i >> 5, c = (i & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + r < T) v = __ldg(reinterpret_cast<const float4*>(E + (int64_t)(t0 + r) * CD + c));
      *reinterpret_cast<uint2*>(Es + r * LDS_ + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
  };
  stage_w(blockIdx.x, 0);
  if (T <= TM) stage_e(0);
  int it = 0;
  for (int slab = blockIdx.x; slab < nslabs; slab += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t col0 = (int64_t)slab * NC;
    const int ncols = (int)((ngp - col0) < NC ? (ngp - col0) : NC);
    stage_w(slab + gridDim.x, buf ^ 1);                         // the other buffer was released by the barrier that ended the previous slab
    cp_async_wait<1>();
    __syncthreads();                                            // this slab (and E on the first one) is visible to every warp
    const uint32_t sW = sbase + (uint32_t)(buf * WSLAB);
    const float* bsl = bs + buf * NC;
    for (int t0 = 0; t0 < T; t0 += TM) {
      if (T > TM) {
        __syncthreads();                                        // previous pass done with Es
        stage_e(t0);
        __syncthreads();
      }
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
        const float b0 = c < ncols ? bsl[c] : 0.f, b1 = c < ncols ? bsl[c + 1] : 0.f;
        *reinterpret_cast<uint32_t*>(Os + r * LDS_ + c) = pack2(acc[j][0] + b0, acc[j][1] + b1);
        *reinterpret_cast<uint32_t*>(Os + (r + 8) * LDS_ + c) = pack2(acc[j][2] + b0, acc[j][3] + b1);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < TM * 16; i += 256) {        // coalesced 16-byte stores, 256 B per task row
        const int r = i >> 4, c = (i & 15) * 8;
        if (t0 + r < T && c < ncols) {
          const int orow = rows ? __ldg(rows + t0 + r) : t0 + r;          // task-switch scheduler: scattered rows of a persistent buffer
          if (orow >= 0 && orow < T_max)
            *reinterpret_cast<uint4*>(out + (int64_t)orow * ngp + col0 + c) = *reinterpret_cast<const uint4*>(Os + r * LDS_ + c);
        }
      }
      __syncthreads();                                          // Os (and this W buffer, after the last pass) may be overwritten
    }
  }
  cp_async_wait<0>();
}

inline int heads_gemm_bf16(cudaStream_t st, const float* E, const bf16* W, const float* bias, bf16* out, int T,
                           const int32_t* rows = nullptr, int T_max = 0, int64_t ngp = NGP) {
  if (ngp % 32 != 0) return fail(HVLA_ERR_ARG, "heads_gemm_bf16: the row stride must be a multiple of 32 elements");
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(heads_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    HVLA_CUDA(cudaFuncSetAttribute(heads_mma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const int grid = 2 * num_sms();                               // two resident CTAs per SM
  const int nslabs = cdiv(ngp, NC);
  ProfScope ps(st, "heads_gemm");
  heads_mma_kernel<<<grid < nslabs ? grid : nslabs, 256, SMEM, st>>>(E, W, bias, out, T, nslabs, rows, rows ? T_max : T, ngp);
  HVLA_LAUNCH_CHECK("heads_mma");
  return HVLA_OK;
}

}  // namespace heads
}  // namespace hvla
