// DINOv2 self-attention (257 tokens, 12 heads x 64) on bf16 tensor cores, flash style.
// One CTA per (head, image, query half); K and V of that head live in shared memory, each warp
// owns one 16-query tile and streams the 257 keys in chunks of 64 with an online softmax.
// Warp-level mma.sync m16n8k16 is used on purpose: per (image, head) the problem is 257x257x64 --
// far below a tcgen05 tile -- and attention is 5% of the step's FLOPs.
// q arrives pre-divided by sqrt(64) (folded into the QKV GEMM epilogue).
#pragma once
#include "common.cuh"

namespace hvla {
inline int tc_num_sms() { return num_sms(); }
namespace attn {

constexpr int S = DTOK;          // 257
constexpr int SP = 272;          // keys padded to 17 tiles of 16
constexpr int ROW = DHD + 8;     // padded smem row (bf16 elements) -> 144 B, conflict-free ldmatrix
constexpr int WARPS = 9;
constexpr int QSPLIT = 2;
constexpr int SMEM = 2 * SP * ROW * 2;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one chunk of NT*8 keys starting at key0.  kaddr / vaddr are this lane's ldmatrix base addresses
// (row = lane-dependent, key 0); TAIL masks keys >= S (only the last chunk has any).
template <int NT, bool TAIL>
__device__ __forceinline__ void chunk(const uint32_t (&qf)[4][4], uint32_t kaddr, uint32_t vaddr, int key0, int lane,
                                      float (&o)[8][4], float& m0, float& m1, float& l0, float& l1) {
  constexpr float LOG2E = 1.4426950408889634f;
  float s[NT][4];
  const uint32_t ka = kaddr + (uint32_t)(key0 * ROW * 2), va = vaddr + (uint32_t)(key0 * ROW * 2);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {   // pairs of 16-wide d steps
      uint32_t b0, b1, b2, b3;
      ldsm_x4(ka + (uint32_t)((nt * 8 * ROW + kp * 32) * 2), b0, b1, b2, b3);
      mma_bf16(s[nt], qf[2 * kp], b0, b1);
      mma_bf16(s[nt], qf[2 * kp + 1], b2, b3);
    }
  }
  float mx0 = m0, mx1 = m1;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (TAIL) {
      const int kbase = key0 + nt * 8 + (lane & 3) * 2;
      if (kbase >= S) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (kbase + 1 >= S) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float n0 = -mx0 * LOG2E, n1 = -mx1 * LOG2E;
  const float c0 = ex2_approx(fmaf(m0, LOG2E, n0)), c1 = ex2_approx(fmaf(m1, LOG2E, n1));   // m = -inf on the first chunk -> 0
  m0 = mx0; m1 = mx1;
  l0 *= c0; l1 *= c1;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) { o[dn][0] *= c0; o[dn][1] *= c0; o[dn][2] *= c1; o[dn][3] *= c1; }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = ex2_approx(fmaf(s[nt][0], LOG2E, n0));
    s[nt][1] = ex2_approx(fmaf(s[nt][1], LOG2E, n0));
    s[nt][2] = ex2_approx(fmaf(s[nt][2], LOG2E, n1));
    s[nt][3] = ex2_approx(fmaf(s[nt][3], LOG2E, n1));
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  // O += P * V
#pragma unroll
  for (int t = 0; t < NT / 2; ++t) {
    uint32_t pa[4];
    pa[0] = pack2(s[2 * t][0], s[2 * t][1]);
    pa[1] = pack2(s[2 * t][2], s[2 * t][3]);
    pa[2] = pack2(s[2 * t + 1][0], s[2 * t + 1][1]);
    pa[3] = pack2(s[2 * t + 1][2], s[2 * t + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide d tiles
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(va + (uint32_t)((16 * t * ROW + dp * 16) * 2), b0, b1, b2, b3);
      mma_bf16(o[2 * dp], pa, b0, b1);
      mma_bf16(o[2 * dp + 1], pa, b2, b3);
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent version: one CTA per SM loops over (image, head) items; 17 warps = the 17 query tiles of
// 16 rows; K/V of the NEXT item are prefetched with cp.async into the second smem buffer while the
// current item is computed, so the staging latency is hidden and every K/V row is read once.
constexpr int PWARPS = 17;
constexpr int PBUF = 2 * SP * ROW;                  // bf16 elements per buffer (K then V)
constexpr int PSMEM = 2 * PBUF * 2;

__device__ __forceinline__ void stage_item(const bf16* __restrict__ qkv, int item, uint32_t sK, uint32_t sV) {
  const int b = item / DH, h = item % DH;
  const bf16* base = qkv + (int64_t)b * S * (3 * DD) + h * DHD;
  for (int i = threadIdx.x; i < S * 16; i += PWARPS * 32) {
    const int r = i >> 4, part = i & 15;            // 16 x 16-byte pieces per row: 8 of K, 8 of V
    const int cc = (part & 7) * 8;
    const bf16* src = base + (int64_t)r * (3 * DD) + (part < 8 ? DD : 2 * DD) + cc;
    cp_async16((part < 8 ? sK : sV) + (uint32_t)((r * ROW + cc) * 2), src);
  }
}

__global__ void __launch_bounds__(PWARPS * 32, 1)
dino_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int n_items) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* buf = reinterpret_cast<bf16*>(smem);
  const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows 257..271 of K and V in both buffers stay zero (scores masked to -inf, P*V contribution zero)
  for (int i = threadIdx.x; i < 2 * 2 * (SP - S) * 8; i += PWARPS * 32) {
    const int which = i / ((SP - S) * 8), j = i % ((SP - S) * 8);
    const int r = S + (j >> 3), c = (j & 7) * 8;
    *reinterpret_cast<uint4*>(buf + which * (SP * ROW) + r * ROW + c) = make_uint4(0, 0, 0, 0);
  }
  int item = blockIdx.x;
  if (item < n_items) stage_item(qkv, item, sbuf, sbuf + SP * ROW * 2);
  cp_async_commit();
  const int row0 = warp * 16 + (lane >> 2), row1 = row0 + 8;
  for (int it = 0; item < n_items; item += gridDim.x, ++it) {
    const int cur = it & 1;
    const uint32_t sK = sbuf + (uint32_t)(cur * PBUF * 2), sV = sK + SP * ROW * 2;
    const int nxt = item + gridDim.x;
    if (nxt < n_items) stage_item(qkv, nxt, sbuf + (uint32_t)((cur ^ 1) * PBUF * 2), sbuf + (uint32_t)((cur ^ 1) * PBUF * 2) + SP * ROW * 2);
    cp_async_commit();
    const int b = item / DH, h = item % DH;
    const bf16* base = qkv + (int64_t)b * S * (3 * DD) + h * DHD;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + (lane & 3) * 2;
      qf[ks][0] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)row0 * (3 * DD) + c)) : 0u;
      qf[ks][1] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)row1 * (3 * DD) + c)) : 0u;
      qf[ks][2] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)row0 * (3 * DD) + c + 8)) : 0u;
      qf[ks][3] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)row1 * (3 * DD) + c + 8)) : 0u;
    }
    cp_async_wait<1>();          // everything but the prefetch just issued has landed
    __syncthreads();
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    // this lane's ldmatrix row addresses at key 0: K (non-transposed): row lane&7, 16-byte column (lane>>3);
    // V (transposed): row (lane>>3 & 1)*8 + (lane&7), column ((lane>>4) * 8)
    const uint32_t kaddr = sK + (uint32_t)(((lane & 7) * ROW + (lane >> 3) * 8) * 2);
    const uint32_t vaddr = sV + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * ROW + (lane >> 4) * 8) * 2);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) chunk<8, false>(qf, kaddr, vaddr, c * 64, lane, o, m0, m1, l0, l1);
    chunk<2, true>(qf, kaddr, vaddr, 256, lane, o, m0, m1, l0, l1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    bf16* ob = out + (int64_t)b * S * DD + h * DHD;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int c = dn * 8 + (lane & 3) * 2;
      if (row0 < S) *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * DD + c) = pack2(o[dn][0] * i0, o[dn][1] * i0);
      if (row1 < S) *reinterpret_cast<uint32_t*>(ob + (int64_t)row1 * DD + c) = pack2(o[dn][2] * i1, o[dn][3] * i1);
    }
    __syncthreads();             // buffer `cur` may be overwritten by the next iteration's prefetch
  }
  cp_async_wait<0>();
}

inline int dino_attention(cudaStream_t st, const bf16* qkv, bf16* out, int B) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(dino_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM));
  }
  const int n_items = B * DH;
  const int grid = n_items < tc_num_sms() ? n_items : tc_num_sms();
  ProfScope ps(st, "dino_attention");
  dino_attention_kernel<<<grid, PWARPS * 32, PSMEM, st>>>(qkv, out, n_items);
  HVLA_LAUNCH_CHECK("dino_attention");
  return HVLA_OK;
}

}  // namespace attn
}  // namespace hvla
