// DINOv2 self-attention on tcgen05 tensor cores (257 tokens, 12 heads x 64, q pre-divided by sqrt(64)).
//
// Per (image, head) item the 257x257 problem is split so that the tensor-core part is a clean 256x256:
//   * query rows 0..255 = two M=128 tiles; keys 0..255 = one N=256 MMA  (S = Q K^T, fp32 in TMEM);
//   * key 256 (the last patch token) is a rank-1 correction on CUDA cores, read from the staged smem tiles;
//   * query row 256 is a single row done on CUDA cores by the softmax warps (32 keys each, merged).
// Softmax is a full-row (not online) two-pass softmax straight out of TMEM: pass 1 row max, pass 2
// P = exp2(..) written back IN PLACE over S as bf16 (tcgen05.st), then O = P V runs with A from TMEM
// (tcgen05.mma TS form) and V as an MN-major shared-memory operand; O lands in the free half of the same
// TMEM region.  The two query tiles of an item are processed by two softmax warp groups in ping-pong (each
// owns 256 TMEM columns), so one group's TMEM/MUFU work overlaps the other's MMA waits.
//
// Warps: 0-3 / 4-7 softmax + epilogue of tile 0 / 1 (TMEM lane quarter = warp & 3), 8 TMA producer,
// 9 MMA issuer.  Query row 256 is shared out over the 8 softmax warps (32 keys each) and merged by warp 0.
// Persistent: one CTA per SM loops over items; Q/K/V of the next item are prefetched (2 smem stages).
#pragma once
#include "gemm_tc.cuh"
#include <stdlib.h>

namespace hvla {
namespace attn5 {

using namespace tc;

constexpr int S_ = DTOK;                       // 257
constexpr int W_TMA = 8, W_MMA = 9;           // warps 0-7: two softmax groups (+ query row 256 cooperatively); 8: TMA; 9: MMA
constexpr int NTHREADS = 10 * 32;
constexpr int TILE_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int KV_BYTES = 272 * 128;            // keys 0..271 (256 = last real key, 257.. = padding rows, P is 0 there)
constexpr int OFF_K = 2 * TILE_BYTES, OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_QT = OFF_V + KV_BYTES;        // query rows 256..271 (only row 256 is used)
constexpr int STAGE_BYTES = OFF_QT + 16 * 128; // Q0 Q1 | K[272] | V[272] | Qtail[16] = 102 KB
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 1024 + 256 + 2 * 8 * 66 * 4;   // + row-256 partials (double-buffered)
constexpr int TM_OREL = 160;                   // TMEM: group g owns columns [256g, 256g+256): S fp32 -> P bf16 in [0,128), O in [160,224)
static_assert(STAGE_BYTES % 1024 == 0 && OFF_V % 1024 == 0, "UMMA / TMA 128B-swizzle tiles need 1024-byte alignment");
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major B operand (V[key][d], 128-byte rows, SWIZZLE_128B): 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                    // leading byte offset: unused (N = 64 is one swizzle atom wide)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-key groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
constexpr uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }   // B is MN-major

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, float acc) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __bfloat1622float2(pa[i]), y = __bfloat1622float2(pb[i]);
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTail, const bf16* __restrict__ qkv,
               bf16* __restrict__ out, int n_items, int dbg, long long* __restrict__ tstamp) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + 2 * STAGE_BYTES;
  const uint32_t in_full = bars, in_empty = bars + 16, s_full = bars + 32, p_full = bars + 48;
  const uint32_t o_full = bars + 64, o_free = bars + 80, tmem_slot = bars + 96;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  float* part = reinterpret_cast<float*>(smem_raw + (bars + 256 - smem_u32(smem_raw)));   // [2][8][66]
  int* cnt = reinterpret_cast<int*>(smem_raw + (bars + 128 - smem_u32(smem_raw)));           // [2] arrival counters
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTail);
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full + 8 * s, 1);
      mbar_init(in_empty + 8 * s, 9);     // MMA commit + 8 softmax warps
      mbar_init(s_full + 8 * s, 1);
      mbar_init(p_full + 8 * s, 4);
      mbar_init(o_full + 8 * s, 1);
      mbar_init(o_free + 8 * s, 4);
    }
    cnt[0] = 0;
    cnt[1] = 0;
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == W_TMA) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int b = item / DH, h = item % DH;
        const uint32_t st = smem_base + s * STAGE_BYTES;
        mbar_wait(in_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(in_full + 8 * s, STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          tma_load_2d(st + j * TILE_BYTES, &tmQKV, in_full + 8 * s, h * DHD, b * S_ + 128 * j);
          tma_load_2d(st + OFF_K + j * TILE_BYTES, &tmQKV, in_full + 8 * s, DD + h * DHD, b * S_ + 128 * j);
          tma_load_2d(st + OFF_V + j * TILE_BYTES, &tmQKV, in_full + 8 * s, 2 * DD + h * DHD, b * S_ + 128 * j);
        }
        // rows 256..271: token 256 of this image (read by the CUDA-core rank-1 paths) + 15 unused rows
        tma_load_2d(st + OFF_K + 2 * TILE_BYTES, &tmTail, in_full + 8 * s, DD + h * DHD, b * S_ + 256);
        tma_load_2d(st + OFF_V + 2 * TILE_BYTES, &tmTail, in_full + 8 * s, 2 * DD + h * DHD, b * S_ + 256);
        tma_load_2d(st + OFF_QT, &tmTail, in_full + 8 * s, h * DHD, b * S_ + 256);
      }
    }
  } else if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, 256);
      constexpr uint32_t idesc_o = make_idesc_bmn(128, 64);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t st = smem_base + s * STAGE_BYTES;
        const uint32_t par = it & 1;        // every per-group barrier completes once per item
        mbar_wait(in_full + 8 * s, (it >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = make_smem_desc(st + OFF_K);
        const uint64_t dv = make_smem_desc_mn(st + OFF_V);
        // Group g owns TMEM columns [256g, 256g+256): S_g = Q_g K^T for both query tiles first ...
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(o_free + 8 * g, par ^ 1);           // group g has read the previous item's O out of this region
          tc_fence_after();
          const uint64_t dq = make_smem_desc(st + g * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 256 * g, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(s_full + 8 * g);
        }
        // ... then O_g = P_g V as soon as group g has written P_g (bf16, in place over S_g); O_g lives in the same region
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(p_full + 8 * g, par);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_bf16_ts(tmem_base + 256 * g + TM_OREL, tmem_base + 256 * g + 8 * k, dv + (uint64_t)(k * (2048 >> 4)), idesc_o,
                         k != 0 ? 1u : 0u);
          umma_commit(o_full + 8 * g);
        }
        umma_commit(in_empty + 8 * s);                  // all MMAs reading this stage have retired
      }
    }
  } else if (warp < 8) {
    // ============================ softmax + epilogue: group g = warp / 4 owns query tile g ============================
    const int g = warp >> 2, quarter = warp & 3;
#define HVLA_TS(k) do { if (tstamp && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4) && it < 6) tstamp[((warp >> 2) * 6 + it) * 8 + (k)] = clock64(); } while (0)
    const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256 * g;
    const int rl = quarter * 32 + lane;                 // row inside the tile
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t par = it & 1;
      const int b = item / DH, h = item % DH;
      const uint8_t* st = smem_al + s * STAGE_BYTES;
      // score against key 256 on CUDA cores from the staged tiles (128B swizzle: chunk ^= row & 7)
      HVLA_TS(0);
      mbar_wait(in_full + 8 * s, (it >> 1) & 1);
      HVLA_TS(1);
      float sx = 0.f;
      {
        const uint8_t* qr = st + g * TILE_BYTES + rl * 128;
        const uint8_t* kr = st + OFF_K + 256 * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          sx = dot8(*reinterpret_cast<const uint4*>(qr + ((u ^ (rl & 7)) << 4)), *reinterpret_cast<const uint4*>(kr + (u << 4)), sx);
      }
      mbar_wait(s_full + 8 * g, par);
      tc_fence_after();
      HVLA_TS(2);
      uint32_t r[2][32];
      // pass 1: row max (TMEM loads double-buffered against the max reduction)
      float mx = sx;
      tmem_ld32(tm, r[0]);
      tmem_wait_ld();
#pragma unroll 1
      for (int c = (dbg & 1) ? 8 : 0; c < 8; c += 2) {
        tmem_ld32(tm + (c + 1) * 32, r[1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[0][i]));
        tmem_wait_ld();
        tmem_ld32(tm + ((c + 2) & 7) * 32, r[0]);        // wraps to chunk 0 = first chunk of pass 2
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[1][i]));
        tmem_wait_ld();
      }
      const float nm = -mx * LOG2E;
      float sum = 0.f;
      HVLA_TS(3);
      // pass 2: P = exp2((s - max) log2e) as bf16, in place over S (the bf16 row is half as wide)
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        tmem_ld32(tm + (c + 1) * 32, r[1]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = (dbg & 4) ? __uint_as_float(r[0][2 * i]) : ex2a(fmaf(__uint_as_float(r[0][2 * i]), LOG2E, nm));
          const float p1 = (dbg & 4) ? __uint_as_float(r[0][2 * i + 1]) : ex2a(fmaf(__uint_as_float(r[0][2 * i + 1]), LOG2E, nm));
          sum += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_wait_ld();
        tmem_st16(tm + c * 16, pk);                       // columns [16c,16c+16) were consumed at chunk <= c
        if (c + 2 < 8) tmem_ld32(tm + (c + 2) * 32, r[0]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = (dbg & 4) ? __uint_as_float(r[1][2 * i]) : ex2a(fmaf(__uint_as_float(r[1][2 * i]), LOG2E, nm));
          const float p1 = (dbg & 4) ? __uint_as_float(r[1][2 * i + 1]) : ex2a(fmaf(__uint_as_float(r[1][2 * i + 1]), LOG2E, nm));
          sum += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_wait_ld();
        tmem_st16(tm + (c + 1) * 16, pk);
      }
      const float px = ex2a(fmaf(sx, LOG2E, nm));
      sum += px;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * g);
      HVLA_TS(4);
      // ---- query row 256 (one row per item), cooperatively: this warp takes keys [32*warp, 32*warp+32) ----
      {
        const uint8_t* sk = st + OFF_K;
        const uint8_t* sv = st + OFF_V;
        const uint8_t* q256 = st + OFF_QT;                 // row 0 of the tail tile: no swizzle offset
        const int key = 32 * warp + lane;
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          a = dot8(*reinterpret_cast<const uint4*>(q256 + (u << 4)), *reinterpret_cast<const uint4*>(sk + key * 128 + ((u ^ (key & 7)) << 4)), a);
        float mw = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
        const float pw = ex2a((a - mw) * LOG2E);
        float lw = pw;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lw += __shfl_xor_sync(0xffffffffu, lw, o);
        float o0 = 0.f, o1 = 0.f;                           // lane owns d = 2*lane, 2*lane+1
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
          const int k2 = 32 * warp + l;
          const float p = __shfl_sync(0xffffffffu, pw, l);
          const float2 v2 = __bfloat1622float2(
              *reinterpret_cast<const __nv_bfloat162*>(sv + k2 * 128 + (((lane >> 2) ^ (k2 & 7)) << 4) + (lane & 3) * 4));
          o0 = fmaf(p, v2.x, o0);
          o1 = fmaf(p, v2.y, o1);
        }
        float* pp = part + ((it & 1) * 8 + warp) * 66;
        if (lane == 0) { pp[0] = mw; pp[1] = lw; }
        pp[2 + 2 * lane] = o0;
        pp[3 + 2 * lane] = o1;
      }
      // epilogue: (O + p_256 v_256) / sum -> bf16 -> global
      const float inv = 1.0f / sum;
      const float pxi = px * inv;
      bf16* orow = out + ((int64_t)b * S_ + g * 128 + rl) * DD + h * DHD;
      const uint8_t* vr = st + OFF_V + 256 * 128;
      mbar_wait(o_full + 8 * g, par);
      tc_fence_after();
      HVLA_TS(5);
      tmem_ld32(tm + TM_OREL, r[0]);
      tmem_ld32(tm + TM_OREL + 32, r[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free + 8 * g);          // O is in registers now
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 vq = *reinterpret_cast<const uint4*>(vr + ((c * 4 + i) << 4));
          const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&vq);
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 v2 = __bfloat1622float2(pv[e]);
            w[e] = pack_bf16(fmaf(pxi, v2.x, __uint_as_float(r[c][i * 8 + 2 * e]) * inv),
                             fmaf(pxi, v2.y, __uint_as_float(r[c][i * 8 + 2 * e + 1]) * inv));
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      HVLA_TS(6);
      {
        const uint8_t* sk = st + OFF_K;
        const uint8_t* sv = st + OFF_V;
        const uint8_t* q256 = st + OFF_QT;
        // the last of the 8 warps to publish its share merges them (no CTA-wide barrier: the groups stay decoupled)
        __threadfence_block();
        __syncwarp();
        int prev = 0;
        if (lane == 0) prev = atomicAdd(cnt + (it & 1), 1);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == 7) {
          __threadfence_block();
          if (lane == 0) cnt[it & 1] = 0;
          float sx2 = 0.f;                                   // key 256 itself
#pragma unroll
          for (int u = 0; u < 8; ++u)
            sx2 = dot8(*reinterpret_cast<const uint4*>(q256 + (u << 4)), *reinterpret_cast<const uint4*>(sk + 256 * 128 + (u << 4)), sx2);
          const volatile float* p0 = part + (it & 1) * 8 * 66;
          float M = sx2;
#pragma unroll
          for (int w = 0; w < 8; ++w) M = fmaxf(M, p0[w * 66]);
          const float ex = ex2a((sx2 - M) * LOG2E);
          float L = ex;
          const float2 vx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + 256 * 128 + 4 * lane));
          float r0 = ex * vx.x, r1 = ex * vx.y;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float f = ex2a((p0[w * 66] - M) * LOG2E);
            L = fmaf(f, p0[w * 66 + 1], L);
            r0 = fmaf(f, p0[w * 66 + 2 + 2 * lane], r0);
            r1 = fmaf(f, p0[w * 66 + 3 + 2 * lane], r1);
          }
          const float il = 1.0f / L;
          *reinterpret_cast<uint32_t*>(out + ((int64_t)b * S_ + 256) * DD + h * DHD + 2 * lane) = pack_bf16(r0 * il, r1 * il);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(in_empty + 8 * s);        // done reading this stage's shared memory
      HVLA_TS(7);
    }
  }
#undef HVLA_TS
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline int dino_attention_tc(cudaStream_t st, const bf16* qkv, bf16* out, int B) {
  static bool attr = false;
  if (!attr) {
    HVLA_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  CUtensorMap map, tail;
  HVLA_TRY(make_map_bf16(&map, qkv, (int64_t)B * S_, 3 * DD, 128));
  HVLA_TRY(make_map_bf16(&tail, qkv, (int64_t)B * S_, 3 * DD, 16));
  const int n_items = B * DH;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  ProfScope ps(st, "dino_attention");
  int dbg = 0;
  if (const char* e = getenv("HVLA_ATTN_DEBUG")) dbg = atoi(e);     // timing experiments only (results are wrong when set)
  static long long* d_ts = nullptr;
  if ((dbg & 32) && !d_ts) { cudaMalloc(&d_ts, 2 * 6 * 8 * sizeof(long long)); cudaMemset(d_ts, 0, 2 * 6 * 8 * sizeof(long long)); }
  launch_k(attn_tc_kernel, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, map, tail, qkv, out, n_items, dbg, (dbg & 32) ? d_ts : (long long*)nullptr);
  if (dbg & 32) {
    long long h[96];
    cudaMemcpy(h, d_ts, sizeof h, cudaMemcpyDeviceToHost);
    for (int g = 0; g < 2; ++g)
      for (int it = 0; it < 6; ++it) {
        const long long* t = h + (g * 6 + it) * 8;
        fprintf(stderr, "ts g%d it%d: start %lld | in_full +%lld s_full +%lld pass1 +%lld pass2 +%lld o_full +%lld epi +%lld row256 +%lld\n", g, it,
                t[0] - h[0], t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6]);
      }
  }
  HVLA_LAUNCH_CHECK("attn_tc");
  return HVLA_OK;
}

}  // namespace attn5

// ====================================================================================================================
// v9 (the pipeline's kernel): per CTA ONE softmax group (4 warps) + 4 helper warps + TMA warp + MMA warp, 256 TMEM columns,
// one shared-memory stage, TWO CTAs per SM.
//   * The two query tiles of an item run back to back in a CTA; one tile's MUFU/TMEM work overlaps the other CTA's MMA,
//     loads and stores because the two co-resident CTAs drift freely against each other.
//   * Everything that is not a 128x256 tile -- the scores against key 256 (a rank-1 column for all 256 rows) and the whole
//     of query row 256 -- is CUDA-core work done by the helper warps OFF the softmax warps' critical path; the softmax warps
//     pick the key-256 score up from shared memory at the end of their max pass.
//   * The stage is refilled in three parts with their own full/empty barriers (Q0,K | V | Q1): Q0 and K of the next item
//     stream in as soon as this item's second S = Q K^T has retired, V as soon as its second P V has, so one stage is enough.
//   * Output rows leave through shared memory (the dead Q tile, 128B-swizzled) and one TMA store per warp: a row-per-thread
//     global store costs one LSU cycle per row and was 1-2 k cycles of every tile's critical path.
//   * Registers are moved between roles with setmaxnreg (softmax 160, helpers 40, TMA/MMA 40).
// Small batches (units < CTA slots) split an item into its two tiles (unit = one query tile).
// ====================================================================================================================
namespace attn9 {
using namespace attn5;

constexpr int NT9 = 12 * 32;                        // warps 0-3 softmax, 4-7 helpers, 8 TMA, 9 MMA, 10-11 register donors
constexpr int A_BYTES = 3 * TILE_BYTES + 2 * 16 * 128;   // Q0 | K[0..255] | K[256..271] | Q[256..271]
constexpr int B_BYTES = TILE_BYTES;                      // Q1
constexpr int C_BYTES = 2 * TILE_BYTES + 16 * 128;       // V[0..271]
static_assert(A_BYTES + B_BYTES + C_BYTES == STAGE_BYTES, "stage parts");
constexpr int OFF_BAR = STAGE_BYTES;
constexpr int OFF_SX = OFF_BAR + 256;               // float [2][256]: score of every query row against key 256 (double-buffered by item)
constexpr int OFF_PS = OFF_SX + 2 * 256 * 4;        // float [264]: un-normalised probabilities of query row 256
constexpr int OFF_RED = OFF_PS + 264 * 4;           // float [8]
constexpr int OFF_OP = OFF_RED + 32;                // float [4][64]: per-warp partial outputs of query row 256
constexpr int SMEM9 = 1024 + OFF_OP + 4 * 64 * 4;
static_assert(2 * (SMEM9 + 1024) <= 233472, "two CTAs per SM");

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void helper_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void cvt8(const uint4& a, float (&f)[8]) {
  f[0] = bf_lo(a.x); f[1] = bf_hi(a.x); f[2] = bf_lo(a.y); f[3] = bf_hi(a.y);
  f[4] = bf_lo(a.z); f[5] = bf_hi(a.z); f[6] = bf_lo(a.w); f[7] = bf_hi(a.w);
}
__device__ __forceinline__ float dot8f(const uint4& a, const float (&k)[8], float acc) {
  acc = fmaf(bf_lo(a.x), k[0], acc); acc = fmaf(bf_hi(a.x), k[1], acc);
  acc = fmaf(bf_lo(a.y), k[2], acc); acc = fmaf(bf_hi(a.y), k[3], acc);
  acc = fmaf(bf_lo(a.z), k[4], acc); acc = fmaf(bf_hi(a.z), k[5], acc);
  acc = fmaf(bf_lo(a.w), k[6], acc); acc = fmaf(bf_hi(a.w), k[7], acc);
  return acc;
}
// MMA issue with the 64-bit shared-memory descriptors kept as (lo, hi) words: stepping K only touches the 14-bit address field in
// lo, so the single issuing thread spends one 32-bit add per MMA instead of 64-bit arithmetic on a serial dependency chain.
// The K-step offsets are added INSIDE the asm block so that the compiler cannot hoist 24 pre-stepped descriptors out of the item loop
// (they would not fit the issuing warp's 40 registers).
template <bool ACC, int KOFF>
__device__ __forceinline__ void umma_ss2(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .b64 da, db;\n\t"
      "add.u32 ta, %1, %5;\n\t"
      "add.u32 tb, %2, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, %6;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "n"(KOFF), "n"(ACC ? 1 : 0)
      : "memory");
}
template <bool ACC, int AOFF, int BOFF>
__device__ __forceinline__ void umma_ts2(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .b64 db;\n\t"
      "add.u32 ta, %1, %5;\n\t"
      "add.u32 tb, %2, %6;\n\t"
      "mov.b64 db, {tb, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %4, %7;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(bhi), "r"(idesc), "n"(AOFF), "n"(BOFF), "n"(ACC ? 1 : 0)
      : "memory");
}
template <int K>
__device__ __forceinline__ void pv_chain(uint32_t tmem_d, uint32_t tmem_a, uint32_t vlo, uint32_t vhi, uint32_t idesc) {
  if constexpr (K < 16) {
    umma_ts2<K != 0, 8 * K, K * (2048 >> 4)>(tmem_d, tmem_a, vlo, vhi, idesc);
    pv_chain<K + 1>(tmem_d, tmem_a, vlo, vhi, idesc);
  }
}

__global__ void __launch_bounds__(NT9, 2)
attn_tc9_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTail, const __grid_constant__ CUtensorMap tmOut,
                int n_units, int split_dbg, long long* __restrict__ tstamp, bf16* __restrict__ out) {
  pdl_trigger();
  const int split = split_dbg & 1, dbg = split_dbg >> 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + OFF_BAR;
  const uint32_t a_full = bars, b_full = bars + 8, c_full = bars + 16, a_empty = bars + 24, b_empty = bars + 32, c_empty = bars + 40;
  const uint32_t s_full = bars + 48, p_full = bars + 56, o_full = bars + 64, o_free = bars + 72, sx_full = bars + 80 /* [2]: one per tile index, completes once per item */, tmem_slot = bars + 96;
  uint8_t* st = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(st + OFF_BAR + 96);
  float* sxs = reinterpret_cast<float*>(st + OFF_SX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTail);
    tma_prefetch_desc(&tmOut);
    mbar_init(a_full, 1);
    mbar_init(b_full, 1);
    mbar_init(c_full, 1);
    mbar_init(a_empty, 9);        // MMA commit (last S of the unit retired) + 4 helper warps + 4 softmax warps (tile-0 output staged in Q0 has left)
    mbar_init(b_empty, 9);        // same for Q1
    mbar_init(c_empty, 9);        // MMA commit (last PV retired) + 4 helper warps + 4 softmax warps (key-256 value row)
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    mbar_init(sx_full, 4);
    mbar_init(sx_full + 8, 4);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp >= 8) {
    reg_dec<40>();
    if (warp == 8 && lane == 0) {
      // ============================ TMA producer ============================
      int it = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int item = split ? (unit >> 1) : unit;
        const int b = item / DH, h = item % DH;
        const uint32_t ph = it & 1;
        mbar_wait(a_empty, ph ^ 1);
        mbar_expect_tx(a_full, A_BYTES);
        tma_load_2d(smem_base, &tmQKV, a_full, h * DHD, b * S_);
        tma_load_2d(smem_base + OFF_K, &tmQKV, a_full, DD + h * DHD, b * S_);
        tma_load_2d(smem_base + OFF_K + TILE_BYTES, &tmQKV, a_full, DD + h * DHD, b * S_ + 128);
        tma_load_2d(smem_base + OFF_K + 2 * TILE_BYTES, &tmTail, a_full, DD + h * DHD, b * S_ + 256);
        tma_load_2d(smem_base + OFF_QT, &tmTail, a_full, h * DHD, b * S_ + 256);
        mbar_wait(c_empty, ph ^ 1);
        mbar_expect_tx(c_full, C_BYTES);
        tma_load_2d(smem_base + OFF_V, &tmQKV, c_full, 2 * DD + h * DHD, b * S_);
        tma_load_2d(smem_base + OFF_V + TILE_BYTES, &tmQKV, c_full, 2 * DD + h * DHD, b * S_ + 128);
        tma_load_2d(smem_base + OFF_V + 2 * TILE_BYTES, &tmTail, c_full, 2 * DD + h * DHD, b * S_ + 256);
        mbar_wait(b_empty, ph ^ 1);
        mbar_expect_tx(b_full, B_BYTES);
        tma_load_2d(smem_base + TILE_BYTES, &tmQKV, b_full, h * DHD, b * S_ + 128);
      }
    } else if (warp == 9 && lane == 0) {
      // ============================ MMA issuer ============================
      constexpr uint32_t idesc_s = make_idesc(128, 256);
      constexpr uint32_t idesc_o = make_idesc_bmn(128, 64);
      const uint64_t dk = make_smem_desc(smem_base + OFF_K);
      const uint64_t dv = make_smem_desc_mn(smem_base + OFF_V);
      const uint32_t klo = (uint32_t)dk, khi = (uint32_t)(dk >> 32), vlo = (uint32_t)dv, vhi = (uint32_t)(dv >> 32);
      int it = 0;
      uint32_t n = 0;                                   // tiles so far: every per-tile barrier completes once per tile
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
        const int t0 = split ? (unit & 1) : 0, t1 = split ? t0 + 1 : 2;
        mbar_wait(a_full, it & 1);
        for (int t = t0; t < t1; ++t, ++n) {
          if (t == 1) mbar_wait(b_full, it & 1);
          mbar_wait(o_free, (n & 1) ^ 1);               // the previous tile's O has been read out of TMEM
          tc_fence_after();
          const uint32_t qlo = klo - ((OFF_K - t * TILE_BYTES) >> 4);     // same descriptor fields, Q tile t's address
          umma_ss2<false, 0>(tmem_base, qlo, klo, khi, idesc_s);
          umma_ss2<true, 2>(tmem_base, qlo, klo, khi, idesc_s);
          umma_ss2<true, 4>(tmem_base, qlo, klo, khi, idesc_s);
          umma_ss2<true, 6>(tmem_base, qlo, klo, khi, idesc_s);
          umma_commit(s_full);
          if (t == t1 - 1) {                            // Q and K may be overwritten once these MMAs retire
            umma_commit(a_empty);
            umma_commit(b_empty);
          }
          mbar_wait(p_full, n & 1);
          if (t == t0) mbar_wait(c_full, it & 1);
          tc_fence_after();
          if (!(dbg & 1)) pv_chain<0>(tmem_base + TM_OREL, tmem_base, vlo, vhi, idesc_o);
          umma_commit(o_full);
        }
        umma_commit(c_empty);
      }
    }
  } else if (warp >= 4) {
    // ============================ helpers: key-256 column for all rows, and query row 256 ============================
    reg_dec<40>();
    const int hw = warp - 4, tid = hw * 32 + lane, sw = tid & 7;
    const uint8_t* sk = st + OFF_K;
    const uint8_t* sv = st + OFF_V;
    float* ps = reinterpret_cast<float*>(st + OFF_PS);
    float* red = reinterpret_cast<float*>(st + OFF_RED);
    float* op = reinterpret_cast<float*>(st + OFF_OP);
    int it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int item = split ? (unit >> 1) : unit;
      const int t0 = split ? (unit & 1) : 0, t1 = split ? t0 + 1 : 2;
      const bool own256 = t1 == 2;                         // the unit holding tile 1 also produces query row 256
      const int b = item / DH, h = item % DH;
      const uint32_t par = it & 1;
      float* sxw = sxs + par * 256;
      mbar_wait(a_full, par);
      // thread tid: row tid against key 256; query 256 against keys tid, 128 + tid and 256 (128B swizzle: chunk ^= row & 7)
      float a0 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 2
      for (int u = 0; u < 8; ++u) {
        const int cs = (u ^ sw) << 4;
        float kc[8], qc[8];
        cvt8(*reinterpret_cast<const uint4*>(sk + 256 * 128 + (u << 4)), kc);
        a0 = dot8f(*reinterpret_cast<const uint4*>(st + tid * 128 + cs), kc, a0);
        cvt8(*reinterpret_cast<const uint4*>(st + OFF_QT + (u << 4)), qc);
#pragma unroll
        for (int e = 0; e < 8; ++e) a2 = fmaf(qc[e], kc[e], a2);
        b0 = dot8f(*reinterpret_cast<const uint4*>(sk + tid * 128 + cs), qc, b0);
        b1 = dot8f(*reinterpret_cast<const uint4*>(sk + (128 + tid) * 128 + cs), qc, b1);
      }
      if (t0 == 0) {
        sxw[tid] = a0;
        __syncwarp();
        if (lane == 0) mbar_arrive(sx_full);
      }
      if (t1 == 2) {
        mbar_wait(b_full, par);
        float a1 = 0.f;                                    // row 128 + tid against key 256
#pragma unroll 2
        for (int u = 0; u < 8; ++u) {
          float kc[8];
          cvt8(*reinterpret_cast<const uint4*>(sk + 256 * 128 + (u << 4)), kc);
          a1 = dot8f(*reinterpret_cast<const uint4*>(st + TILE_BYTES + tid * 128 + ((u ^ sw) << 4)), kc, a1);
        }
        sxw[128 + tid] = a1;
        __syncwarp();
        if (lane == 0) mbar_arrive(sx_full + 8);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_empty);
        mbar_arrive(b_empty);
      }
      if (own256) {
        float m = fmaxf(fmaxf(b0, b1), a2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[hw] = m;
        helper_sync();
        m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        const float p0 = ex2a((b0 - m) * LOG2E), p1 = ex2a((b1 - m) * LOG2E), p2 = ex2a((a2 - m) * LOG2E);
        float l = p0 + p1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        ps[tid] = p0;
        ps[128 + tid] = p1;
        if (lane == 0) red[4 + hw] = l;
        helper_sync();
        l = red[4] + red[5] + red[6] + red[7] + p2;
        mbar_wait(c_full, par);
        float o0 = 0.f, o1 = 0.f;                           // warp hw: keys [64 hw, 64 hw + 64); lane owns d = 2 lane, 2 lane + 1
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
          const int k2 = 64 * hw + j;
          const float p = ps[k2];
          const uint32_t v2 = *reinterpret_cast<const uint32_t*>(sv + k2 * 128 + (((lane >> 2) ^ (k2 & 7)) << 4) + (lane & 3) * 4);
          o0 = fmaf(p, bf_lo(v2), o0);
          o1 = fmaf(p, bf_hi(v2), o1);
        }
        if (hw == 0) {
          const uint32_t v2 = *reinterpret_cast<const uint32_t*>(sv + 256 * 128 + lane * 4);
          o0 = fmaf(p2, bf_lo(v2), o0);
          o1 = fmaf(p2, bf_hi(v2), o1);
        }
        op[hw * 64 + 2 * lane] = o0;
        op[hw * 64 + 2 * lane + 1] = o1;
        __syncwarp();
        if (lane == 0) mbar_arrive(c_empty);
        helper_sync();
        if (hw == 0) {
          const float il = 1.0f / l;
          const float r0 = (op[2 * lane] + op[64 + 2 * lane]) + (op[128 + 2 * lane] + op[192 + 2 * lane]);
          const float r1 = (op[2 * lane + 1] + op[64 + 2 * lane + 1]) + (op[128 + 2 * lane + 1] + op[192 + 2 * lane + 1]);
          *reinterpret_cast<uint32_t*>(out + ((int64_t)b * S_ + 256) * DD + h * DHD + 2 * lane) = pack_bf16(r0 * il, r1 * il);
        }
      } else if (lane == 0) {
        mbar_arrive(c_empty);
      }
    }
  } else {
    // ============================ softmax + epilogue (TMEM lane quarter = warp) ============================
    reg_inc<160>();
#define TS9(k) do { if (tstamp && blockIdx.x == 0 && threadIdx.x == 0 && n < 6) tstamp[n * 8 + (k)] = clock64(); } while (0)
    const uint32_t tm = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int rl = warp * 32 + lane;
    const uint8_t* vr = st + OFF_V + 256 * 128;
    int it = 0;
    uint32_t n = 0;
    int pending = -1;                                      // tile whose staged output (in its Q tile) is still being read by TMA
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int item = split ? (unit >> 1) : unit;
      const int t0 = split ? (unit & 1) : 0, t1 = split ? t0 + 1 : 2;
      const int b = item / DH, h = item % DH;
      const uint32_t par = it & 1;
      for (int t = t0; t < t1; ++t, ++n) {
        TS9(0);
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        TS9(1);
        if (pending >= 0) {                                // long done by now: release the previous tile's Q buffer to the producer
          if (lane == 0) {
            bulk_wait_read0();
            mbar_arrive(pending == 0 ? a_empty : b_empty);
          }
          pending = -1;
        }
        uint32_t r[2][32];
        // pass 1: row max (TMEM loads double-buffered against the max reduction)
        float mx = -3.0e38f;
        tmem_ld32(tm, r[0]);
        tmem_wait_ld();
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          tmem_ld32(tm + (c + 1) * 32, r[1]);
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[0][i]));
          tmem_wait_ld();
          tmem_ld32(tm + ((c + 2) & 7) * 32, r[0]);        // wraps to chunk 0 = first chunk of pass 2
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[1][i]));
          tmem_wait_ld();
        }
        mbar_wait(sx_full + 8 * t, par);                   // the helpers' key-256 column of this tile
        const float sx = sxs[par * 256 + t * 128 + rl];
        mx = fmaxf(mx, sx);
        const float nm = -mx * LOG2E;
        float sum = 0.f;
        TS9(2);
        // pass 2: P = exp2((s - max) log2e) as bf16, in place over S (the bf16 row is half as wide)
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          tmem_ld32(tm + (c + 1) * 32, r[1]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2a(fmaf(__uint_as_float(r[0][2 * i]), LOG2E, nm));
            const float p1 = ex2a(fmaf(__uint_as_float(r[0][2 * i + 1]), LOG2E, nm));
            sum += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
          }
          tmem_wait_ld();
          tmem_st16(tm + c * 16, pk);                       // columns [16c,16c+16) were consumed at chunk <= c
          if (c + 2 < 8) tmem_ld32(tm + (c + 2) * 32, r[0]);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2a(fmaf(__uint_as_float(r[1][2 * i]), LOG2E, nm));
            const float p1 = ex2a(fmaf(__uint_as_float(r[1][2 * i + 1]), LOG2E, nm));
            sum += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
          }
          tmem_wait_ld();
          tmem_st16(tm + (c + 1) * 16, pk);
        }
        const float px = ex2a(fmaf(sx, LOG2E, nm));
        sum += px;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        TS9(3);
        // epilogue: (O + p_256 v_256) / sum -> bf16 -> this warp's 32 x 128 B slice of the dead Q tile -> one TMA store
        const float inv = 1.0f / sum;
        const float pxi = px * inv;
        uint8_t* stg = st + t * TILE_BYTES + warp * 4096 + lane * 128;
        mbar_wait(o_full, n & 1);
        if (t == t0) mbar_wait(c_full, par);               // makes the TMA-written V row 256 visible to this thread
        tc_fence_after();
        TS9(4);
        tmem_ld32(tm + TM_OREL, r[0]);
        tmem_ld32(tm + TM_OREL + 32, r[1]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);                  // O is in registers now: the next tile's S may overwrite it
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 vq = *reinterpret_cast<const uint4*>(vr + ((c * 4 + i) << 4));
            const uint32_t vw[4] = {vq.x, vq.y, vq.z, vq.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w[e] = pack_bf16(fmaf(pxi, bf_lo(vw[e]), __uint_as_float(r[c][i * 8 + 2 * e]) * inv),
                               fmaf(pxi, bf_hi(vw[e]), __uint_as_float(r[c][i * 8 + 2 * e + 1]) * inv));
            *reinterpret_cast<uint4*>(stg + (((c * 4 + i) ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(dbg & 2)) {
          tma_store_2d(&tmOut, smem_base + t * TILE_BYTES + warp * 4096, h * DHD, b * S_ + t * 128 + warp * 32);
          bulk_commit();
        }
        pending = t;
        if (split) {                                       // one tile per unit: nothing later in this unit could release the buffers
          if (lane == 0) {
            bulk_wait_read0();
            mbar_arrive(a_empty);
            mbar_arrive(b_empty);
          }
          pending = -1;
        }
        TS9(5);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(c_empty);                 // done reading the value tile's row 256
    }
    if (lane == 0) bulk_wait0();                           // the last staged tile must have left before the CTA's shared memory goes away
#undef TS9
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

inline int dino_attention_tc9(cudaStream_t st, const bf16* qkv, bf16* out, int B) {
  static bool attr = false;
  if (!attr) {
    HVLA_CUDA(cudaFuncSetAttribute(attn_tc9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM9));
    attr = true;
  }
  CUtensorMap map, tail, omap;
  HVLA_TRY(make_map_bf16(&map, qkv, (int64_t)B * S_, 3 * DD, 128));
  HVLA_TRY(make_map_bf16(&tail, qkv, (int64_t)B * S_, 3 * DD, 16));
  HVLA_TRY(make_map_bf16(&omap, out, (int64_t)B * S_, DD, 32));
  const int n_items = B * DH;
  const int slots = 2 * num_sms();
  const int split = (2 * n_items <= slots) ? 1 : 0;          // small batches: one query tile per CTA
  const int n_units = split ? 2 * n_items : n_items;
  const int grid = n_units < slots ? n_units : slots;
  ProfScope ps(st, "dino_attention");
  static long long* d_ts = nullptr;
  const bool ts = getenv("HVLA_ATTN_TS") != nullptr;            // phase timestamps of CTA 0 (experiments)
  if (ts && !d_ts) { cudaMalloc(&d_ts, 48 * sizeof(long long)); cudaMemset(d_ts, 0, 48 * sizeof(long long)); }
  int dbg = 0;
  if (const char* e = getenv("HVLA_ATTN_DEBUG")) dbg = atoi(e);      // timing experiments only (results are wrong when set)
  launch_k(attn_tc9_kernel, dim3(grid), dim3(NT9), (size_t)SMEM9, st, map, tail, omap, n_units, split | (dbg << 4),
           ts ? d_ts : (long long*)nullptr, out);
  if (ts) {
    long long h[48];
    cudaMemcpy(h, d_ts, sizeof h, cudaMemcpyDeviceToHost);
    for (int n = 0; n < 6; ++n) {
      const long long* t = h + n * 8;
      fprintf(stderr, "ts9 tile %d: start %lld | s_full +%lld pass1 +%lld pass2 +%lld o_full +%lld epi +%lld\n", n, t[0] - h[0], t[1] - t[0],
              t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
    }
  }
  HVLA_LAUNCH_CHECK("attn_tc9");
  return HVLA_OK;
}

}  // namespace attn9

}  // namespace hvla
