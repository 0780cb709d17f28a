// DINOv2 with fp32-class accuracy ON THE TENSOR CORES ("bf16x3", dtype HVLA_BF16X3): the measured answer to "fp32 parity vs
// tensor cores" (SURVEY.md section 7).  The fp32 path (HVLA_F32) runs every GEMM and the attention on CUDA cores; here every
// matrix product is a tcgen05 / mma.sync product of SPLIT operands: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa
// bits together) and  A W^T ~= Ahi Whi^T + Alo Whi^T + Ahi Wlo^T  (the dropped lo.lo term is 2^-18 relative), accumulated in fp32.
//   * linear layers: ONE launch of the 2-CTA tcgen05 GEMM per matrix, the three products concatenated along K: A' = [Ahi | Alo | Ahi]
//     (M x 3K, written by the producing kernel / epilogue), W' = [Whi | Whi | Wlo] (N x 3K, packed once by hvla/params.py);
//   * q|k|v and fc1 leave their GEMM already split (EPI_SPLIT_BF16 / EPI_SPLIT_GELU_BF16, gemm_tc.cuh), proj and fc2 reduce-add
//     into the fp32 residual stream (EPI_RESIDUAL_F32 with bias and LayerScale), LayerNorm writes the split operand directly;
//   * attention: warp-level mma.sync flash attention with Q, K, V and P split the same way (three products for S = Q K^T, three
//     for O = P V), fp32 softmax statistics.
// Reference semantics: FlaxDinov2Module (transformers 4.50.0) as called at hypervla/components/base_vit.py:111-122, in fp32.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
#include "gemm_tc2.cuh"
#include "attn_mma.cuh"
#include "t5_embed.cuh"

namespace hvla {
namespace x3 {

using t5::store_split4;

// LayerNorm (eps 1e-6, fast variance, scale / bias) of the fp32 stream straight into the split operand [rows, 3 * 768]
__global__ void __launch_bounds__(256) ln768_split_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                                                          bf16* __restrict__ a, int rows) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * DD);
  float4 v[6];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float mean = s / 768.f;
  const float rstd = 1.0f / sqrtf(fmaxf(0.f, s2 / 768.f - mean * mean) + 1e-6f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(scale + c)), b = __ldg(reinterpret_cast<const float4*>(bias + c));
    const float y[4] = {(v[i].x - mean) * (rstd * g.x) + b.x, (v[i].y - mean) * (rstd * g.y) + b.y, (v[i].z - mean) * (rstd * g.z) + b.z,
                        (v[i].w - mean) * (rstd * g.w) + b.w};
    store_split4(a, row, c, DD, y);
  }
}

// ---- split-operand attention --------------------------------------------------------------------------------------------------
// One CTA per (image, head, query half): 9 warps x 16 query rows; K and V of the head (hi and lo planes) staged in shared memory.
// Input  qkv2 [B*257, 2 * 2304] bf16 = [hi plane | lo plane] of q | k | v (q NOT pre-scaled: 1/8 is applied to the fp32 scores, exactly);
// output a3   [B*257, 3 * 768]  bf16 = [hi | lo | hi] of the attention output, the A operand of the out-projection.
namespace at3 {
using attn::cp_async16;
using attn::cp_async_commit;
using attn::cp_async_wait;
using attn::ex2_approx;
using attn::ldsm_x4;
using attn::ldsm_x4_t;
using attn::mma_bf16;

constexpr int S = DTOK, SP = 272, ROW = DHD + 8, WARPS = 9, HALF_ROWS = WARPS * 16;     // 144 query rows per CTA
constexpr int PLANE = SP * ROW * 2;                                                       // bytes of one staged plane (K or V, hi or lo)
constexpr int SMEM = 4 * PLANE;                                                           // Kh | Kl | Vh | Vl
constexpr int LDQ = 2 * 3 * DD;                                                           // row length of qkv2

__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int NT, bool TAIL>
__device__ __forceinline__ void chunk3(const uint32_t (&qh)[4][4], const uint32_t (&ql)[4][4], uint32_t kaddr, uint32_t vaddr, int key0, int lane,
                                       float (&o)[8][4], float& m0, float& m1, float& l0, float& l1) {
  constexpr float SC = 0.125f * 1.4426950408889634f;      // 1/sqrt(64) (exact) and log2(e) in one factor
  float s[NT][4];
  const uint32_t ka = kaddr + (uint32_t)(key0 * ROW * 2), va = vaddr + (uint32_t)(key0 * ROW * 2);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(ka + (uint32_t)((nt * 8 * ROW + kp * 32) * 2), b0, b1, b2, b3);                 // K hi
      mma_bf16(s[nt], qh[2 * kp], b0, b1);
      mma_bf16(s[nt], qh[2 * kp + 1], b2, b3);
      mma_bf16(s[nt], ql[2 * kp], b0, b1);
      mma_bf16(s[nt], ql[2 * kp + 1], b2, b3);
      ldsm_x4(ka + (uint32_t)(PLANE + (nt * 8 * ROW + kp * 32) * 2), b0, b1, b2, b3);         // K lo
      mma_bf16(s[nt], qh[2 * kp], b0, b1);
      mma_bf16(s[nt], qh[2 * kp + 1], b2, b3);
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
  const float n0 = -mx0 * SC, n1 = -mx1 * SC;
  const float c0 = ex2_approx(fmaf(m0, SC, n0)), c1 = ex2_approx(fmaf(m1, SC, n1));        // m = -inf on the first chunk -> 0
  m0 = mx0; m1 = mx1;
  l0 *= c0; l1 *= c1;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) { o[dn][0] *= c0; o[dn][1] *= c0; o[dn][2] *= c1; o[dn][3] *= c1; }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = ex2_approx(fmaf(s[nt][0], SC, n0));
    s[nt][1] = ex2_approx(fmaf(s[nt][1], SC, n0));
    s[nt][2] = ex2_approx(fmaf(s[nt][2], SC, n1));
    s[nt][3] = ex2_approx(fmaf(s[nt][3], SC, n1));
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
#pragma unroll
  for (int t = 0; t < NT / 2; ++t) {
    uint32_t ph[4], pl[4];
    ph[0] = pack_hi_lo(s[2 * t][0], s[2 * t][1], pl[0]);
    ph[1] = pack_hi_lo(s[2 * t][2], s[2 * t][3], pl[1]);
    ph[2] = pack_hi_lo(s[2 * t + 1][0], s[2 * t + 1][1], pl[2]);
    ph[3] = pack_hi_lo(s[2 * t + 1][2], s[2 * t + 1][3], pl[3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(va + (uint32_t)((16 * t * ROW + dp * 16) * 2), b0, b1, b2, b3);                // V hi
      mma_bf16(o[2 * dp], ph, b0, b1);
      mma_bf16(o[2 * dp + 1], ph, b2, b3);
      mma_bf16(o[2 * dp], pl, b0, b1);
      mma_bf16(o[2 * dp + 1], pl, b2, b3);
      ldsm_x4_t(va + (uint32_t)(PLANE + (16 * t * ROW + dp * 16) * 2), b0, b1, b2, b3);        // V lo
      mma_bf16(o[2 * dp], ph, b0, b1);
      mma_bf16(o[2 * dp + 1], ph, b2, b3);
    }
  }
}

__global__ void __launch_bounds__(WARPS * 32, 1) attention_x3_kernel(const bf16* __restrict__ qkv2, bf16* __restrict__ a3) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sK = sb, sV = sb + 2 * PLANE;                   // [hi, lo] planes each
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int b = item / DH, h = item % DH;
  const bf16* base = qkv2 + (int64_t)b * S * LDQ + h * DHD;
  // rows 257..271 of every plane: zero (scores masked to -inf, P V contribution zero)
  for (int i = threadIdx.x; i < 4 * (SP - S) * 8; i += WARPS * 32) {
    const int pl = i / ((SP - S) * 8), j = i % ((SP - S) * 8);
    *reinterpret_cast<uint4*>(smem + pl * PLANE + ((S + (j >> 3)) * ROW + (j & 7) * 8) * 2) = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < S * 32; i += WARPS * 32) {        // per row: 8 pieces of K hi, K lo, V hi, V lo
    const int r = i >> 5, part = i & 31, pl = part >> 3, cc = (part & 7) * 8;
    const bf16* src = base + (int64_t)r * LDQ + ((pl >> 1) ? 2 * DD : DD) + (pl & 1) * 3 * DD + cc;
    cp_async16(sb + (uint32_t)(pl * PLANE + (r * ROW + cc) * 2), src);
  }
  cp_async_commit();
  const int row0 = half * HALF_ROWS + warp * 16 + (lane >> 2), row1 = row0 + 8;
  uint32_t qh[4][4], ql[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = ks * 16 + (lane & 3) * 2;
    const bf16* q0 = base + (int64_t)row0 * LDQ + c;
    const bf16* q1 = base + (int64_t)row1 * LDQ + c;
    qh[ks][0] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0)) : 0u;
    qh[ks][1] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1)) : 0u;
    qh[ks][2] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 8)) : 0u;
    qh[ks][3] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 8)) : 0u;
    ql[ks][0] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 3 * DD)) : 0u;
    ql[ks][1] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 3 * DD)) : 0u;
    ql[ks][2] = row0 < S ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 3 * DD + 8)) : 0u;
    ql[ks][3] = row1 < S ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 3 * DD + 8)) : 0u;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (half * HALF_ROWS + warp * 16 >= S) return;                  // the second half has 113 real rows: its last warps have none
  float o[8][4];
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t kaddr = sK + (uint32_t)(((lane & 7) * ROW + (lane >> 3) * 8) * 2);
  const uint32_t vaddr = sV + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * ROW + (lane >> 4) * 8) * 2);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) chunk3<8, false>(qh, ql, kaddr, vaddr, c * 64, lane, o, m0, m1, l0, l1);
  chunk3<2, true>(qh, ql, kaddr, vaddr, 256, lane, o, m0, m1, l0, l1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  bf16* ob = a3 + (int64_t)b * S * (3 * DD) + h * DHD;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    const int c = dn * 8 + (lane & 3) * 2;
    uint32_t lo;
    if (row0 < S) {
      const uint32_t hi = pack_hi_lo(o[dn][0] * i0, o[dn][1] * i0, lo);
      bf16* r = ob + (int64_t)row0 * (3 * DD) + c;
      *reinterpret_cast<uint32_t*>(r) = hi; *reinterpret_cast<uint32_t*>(r + DD) = lo; *reinterpret_cast<uint32_t*>(r + 2 * DD) = hi;
    }
    if (row1 < S) {
      const uint32_t hi = pack_hi_lo(o[dn][2] * i1, o[dn][3] * i1, lo);
      bf16* r = ob + (int64_t)row1 * (3 * DD) + c;
      *reinterpret_cast<uint32_t*>(r) = hi; *reinterpret_cast<uint32_t*>(r + DD) = lo; *reinterpret_cast<uint32_t*>(r + 2 * DD) = hi;
    }
  }
}

inline int attention_x3(cudaStream_t st, const bf16* qkv2, bf16* a3, int B) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) HVLA_CUDA(cudaFuncSetAttribute(attention_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  ProfScope ps(st, "dino_attention");
  launch_k(attention_x3_kernel, dim3(2 * B * DH), dim3(WARPS * 32), (size_t)SMEM, st, qkv2, a3);
  HVLA_LAUNCH_CHECK("attention_x3");
  return HVLA_OK;
}
}  // namespace at3

// whole residual stream before the patch-embedding GEMM accumulates onto it: X[b,r,:] = pos[r,:] (+ cls on row 0)
__global__ void __launch_bounds__(256) init_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ X, int64_t total4) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = (int)(i % (DD / 4));
  const int r = (int)((i / (DD / 4)) % DTOK);
  float4 v = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)r * DD) + c4);
  if (r == 0) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(cls) + c4);
    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
  }
  reinterpret_cast<float4*>(X)[i] = v;
}

// bytes of scratch of the x3 flow beyond the fp32 stream: A' of the 768-wide inputs, [hi | lo] of q|k|v, A' of the MLP hidden layer
struct Ws {
  float* X; bf16* A3; bf16* QKV2; bf16* A3L; float* A0;
};

// images -> last_hidden_state (fp32) [B*257, 768]
inline int dino_forward(cudaStream_t st, const float* dv, const bf16* dm3, const uint8_t* images, int B, float* out_emb, const Ws& w) {
  typedef DvecLayout V;
  typedef DmatLayout Mx;
  const int M = B * DTOK;
  auto gemm = [&](const bf16* A, int64_t w_off, int m, int n, int k, int epi, tc::EpiP& ep) -> int {
    ep.part = nullptr;                                          // no K split in this flow (one fp32 accumulation order)
    return tc2::gemm_tc2(st, A, dm3 + 3 * w_off, m, n, 3 * k, epi, ep);
  };
  {
    const int64_t total = (int64_t)B * NPATCH * PATCH_KP;
    {
      ProfScope ps(st, "im2col");
      im2col_norm_kernel<float><<<cdiv(total, 256), 256, 0, st>>>(images, w.A0, B);
      HVLA_LAUNCH_CHECK("im2col");
      t5::split_kernel<<<cdiv(total / 4, 256), 256, 0, st>>>(w.A0, w.A3L, total / 4, PATCH_KP, 0);
      HVLA_LAUNCH_CHECK("split");
    }
    ProfScope ps2(st, "cls_rows");
    const int64_t total4 = (int64_t)M * DD / 4;
    launch_k(init_rows_kernel, dim3(cdiv(total4, 256)), dim3(256), 0, st, dv + V::cls, dv + V::pos, w.X, total4);
    HVLA_LAUNCH_CHECK("dino_init_rows");
  }
  {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = dv + V::patch_b; ep.out = w.X; ep.ldo = DD; ep.patch_rows = 1;
    HVLA_TRY(gemm(w.A3L, Mx::patch_w, B * NPATCH, DD, PATCH_KP, tc::EPI_PATCH_F32, ep));
  }
  for (int l = 0; l < DL; ++l) {
    const float* v = dv + V::layers + (int64_t)l * V::layer_size;
    const int64_t m = Mx::layers + (int64_t)l * Mx::layer_size;
    {
      ProfScope ps(st, "layernorm");
      launch_k(ln768_split_kernel, dim3(cdiv(M, 8)), dim3(256), 0, st, w.X, v + V::ln1_s, v + V::ln1_b, w.A3, M);
      HVLA_LAUNCH_CHECK("ln768_split");
    }
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::bqkv; ep.out = w.QKV2; ep.ldo = 2 * 3 * DD; ep.plane_stride = 3 * DD; ep.nplanes = 2;
      HVLA_TRY(gemm(w.A3, m + Mx::wqkv, M, 3 * DD, DD, tc::EPI_SPLIT_BF16, ep));
    }
    HVLA_TRY(at3::attention_x3(st, w.QKV2, w.A3, B));
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::bo; ep.out = w.X; ep.ldo = DD; ep.ls = v + V::ls1;
      HVLA_TRY(gemm(w.A3, m + Mx::wo, M, DD, DD, tc::EPI_RESIDUAL_F32, ep));
    }
    {
      ProfScope ps(st, "layernorm");
      launch_k(ln768_split_kernel, dim3(cdiv(M, 8)), dim3(256), 0, st, w.X, v + V::ln2_s, v + V::ln2_b, w.A3, M);
      HVLA_LAUNCH_CHECK("ln768_split");
    }
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::b1; ep.out = w.A3L; ep.ldo = 3 * DF; ep.plane_stride = DF; ep.nplanes = 3;
      HVLA_TRY(gemm(w.A3, m + Mx::w1, M, DF, DD, tc::EPI_SPLIT_GELU_BF16, ep));
    }
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::b2; ep.out = w.X; ep.ldo = DD; ep.ls = v + V::ls2;
      HVLA_TRY(gemm(w.A3L, m + Mx::w2, M, DD, DF, tc::EPI_RESIDUAL_F32, ep));
    }
  }
  LnP ln; memset(&ln, 0, sizeof ln);
  ln.x = w.X; ln.ldx = DD; ln.y = out_emb; ln.ldy = DD; ln.scale = dv + V::lnf_s; ln.bias = dv + V::lnf_b; ln.rows = M; ln.rows_per_batch = 1;
  return layernorm<float, float>(st, ln, DD);
}

}  // namespace x3
}  // namespace hvla
