// T5-base token embedder on the GPU: the step upstream of generate.  The reference embeds the
// tokenised instruction with FlaxT5EncoderModel('t5-base') and feeds last_hidden_state to the hypernetwork as
// `token_embedding` (octo/model/components/tokenizers.py:186-211, data/utils/language_tokenizer.py:9-28,
// data/simpler/evaluate.py:240-262).  T5 v1.0 encoder: RMS LayerNorm (no mean, no bias), bias-free linears, un-scaled
// QK^T + bucketed relative-position bias (block 0's table for all blocks) + additive padding mask, ReLU MLP.
// fp32 path on the generic CUDA-core kernels (exact to 1e-5 against the oracle); it runs once per task switch on
// 32 tokens per task.  Weights in the HF torch layout ([out, in]), packed by hvla/t5.py.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
#include "gemm_tc2.cuh"
#include <cfloat>

namespace hvla {
namespace t5 {

constexpr int TD = 768, TH = 12, THD = 64, TFF = 3072, TL = 12, TV = 32128, SMAX = 32;
struct Layout {
  static constexpr int64_t embed = 0;
  static constexpr int64_t layers = embed + (int64_t)TV * TD;
  static constexpr int64_t ln0 = 0, wqkv = ln0 + TD, wo = wqkv + (int64_t)3 * TD * TD, ln1 = wo + (int64_t)TD * TD, wi = ln1 + TD,
                           wo2 = wi + (int64_t)TFF * TD, layer_size = wo2 + (int64_t)TD * TFF;
  static constexpr int64_t lnf = layers + TL * layer_size;
  static constexpr int64_t total = lnf + TD;
};

__global__ void __launch_bounds__(192) gather_kernel(const int32_t* __restrict__ ids, const float* __restrict__ embed, float* __restrict__ X,
                                                     int M) {
  const int m = blockIdx.x;
  if (m >= M) return;
  int id = ids[m];
  id = id < 0 ? 0 : (id >= TV ? TV - 1 : id);
  reinterpret_cast<float4*>(X + (int64_t)m * TD)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(embed + (int64_t)id * TD) + threadIdx.x);
}

// T5LayerNorm: y = x * rsqrt(mean(x^2) + 1e-6) * w ; one warp per row
__global__ void __launch_bounds__(256) rmsnorm768_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * TD);
  float4 v[6];
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = xr[lane + 32 * i];
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const float r = 1.0f / sqrtf(s2 / (float)TD + 1e-6f);
  float4* yr = reinterpret_cast<float4*>(y + (int64_t)row * TD);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    yr[lane + 32 * i] = make_float4(v[i].x * r * g.x, v[i].y * r * g.y, v[i].z * r * g.z, v[i].w * r * g.w);
  }
}

// one CTA (4 warps) per (head h, task t): q, k, v of the head staged in shared memory (k rows padded to 65 floats: lane = key
// reads a column conflict-free), a warp owns every 4th query.  s = q.k (no 1/sqrt(d)) + pos_bias[h,i,j] + (1-mask_j)*finfo.min;
// the dot product runs over d ascending and the output over keys ascending (the summation order of the oracle's einsum is not
// defined; this one is fixed).  The first version ran one single-warp CTA per (query, head, task) straight from global memory:
// 211 us per layer for 64 instructions, half of the whole encoder.
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ mask,
                                                        const float* __restrict__ pos_bias, float* __restrict__ out, int S) {
  __shared__ __align__(16) float sq[SMAX][THD];
  __shared__ float sk[SMAX][THD + 1];
  __shared__ __align__(16) float sv[SMAX][THD];
  const int h = blockIdx.x, t = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < S * (THD / 4); e += blockDim.x) {
    const int r = e / (THD / 4), c4 = e % (THD / 4);
    const float4* src = reinterpret_cast<const float4*>(qkv + ((int64_t)t * S + r) * (3 * TD) + h * THD) + c4;
    const float4 q4 = __ldg(src), k4 = __ldg(src + TD / 4), v4 = __ldg(src + 2 * TD / 4);
    *reinterpret_cast<float4*>(&sq[r][c4 * 4]) = q4;
    *reinterpret_cast<float4*>(&sv[r][c4 * 4]) = v4;
    sk[r][c4 * 4] = k4.x; sk[r][c4 * 4 + 1] = k4.y; sk[r][c4 * 4 + 2] = k4.z; sk[r][c4 * 4 + 3] = k4.w;
  }
  __syncthreads();
  const float neg = (lane < S && mask[t * S + lane] != 0) ? 0.f : -FLT_MAX;
  for (int i = warp; i < S; i += 4) {
    float s = -INFINITY;
    if (lane < S) {
      float a = 0.f;
#pragma unroll 16
      for (int d = 0; d < THD; ++d) a = fmaf(sq[i][d], sk[lane][d], a);
      s = a + __ldg(pos_bias + ((int64_t)h * S + i) * S + lane) + neg;
    }
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = lane < S ? expf(s - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float p = e / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      o0 = fmaf(pj, sv[j][lane], o0);
      o1 = fmaf(pj, sv[j][lane + 32], o1);
    }
    float* o = out + ((int64_t)t * S + i) * TD + h * THD;
    o[lane] = o0;
    o[lane + 32] = o1;
  }
}

inline size_t workspace_bytes(int T, int S) {
  const size_t M = (size_t)T * S;
  return ((M * TD * 4 + 255) & ~(size_t)255) * 3 + ((M * 3 * TD * 4 + 255) & ~(size_t)255) + ((M * TFF * 4 + 255) & ~(size_t)255);
}

inline GemmP gp(const float* A, int lda, const float* Wt, int ldw, float* C, int ldc, int M, int N, int K) {
  GemmP g;
  memset(&g, 0, sizeof g);
  g.A = A; g.lda = lda; g.W = Wt; g.ldw = ldw; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.wt = 1; g.out_scale = 1.f;
  return g;
}

inline int encode(cudaStream_t st, const float* blob, const float* pos_bias, const int32_t* ids, const int32_t* mask, int T, int S, float* out,
                  uint8_t* ws) {
  typedef Layout L;
  const int M = T * S;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  float* X = reinterpret_cast<float*>(ws);
  float* Y = reinterpret_cast<float*>(ws + al((size_t)M * TD * 4));
  float* ATT = reinterpret_cast<float*>(ws + 2 * al((size_t)M * TD * 4));
  float* QKV = reinterpret_cast<float*>(ws + 3 * al((size_t)M * TD * 4));
  float* HID = reinterpret_cast<float*>(ws + 3 * al((size_t)M * TD * 4) + al((size_t)M * 3 * TD * 4));
  ProfScope ps(st, "t5_encode");
  gather_kernel<<<M, 192, 0, st>>>(ids, blob + L::embed, X, M);
  HVLA_LAUNCH_CHECK("t5_gather");
  for (int l = 0; l < TL; ++l) {
    const float* w = blob + L::layers + (int64_t)l * L::layer_size;
    rmsnorm768_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, w + L::ln0, Y, M);
    HVLA_LAUNCH_CHECK("t5_rmsnorm");
    HVLA_TRY((gemm_simt<float, float, float, float>(st, gp(Y, TD, w + L::wqkv, TD, QKV, 3 * TD, M, 3 * TD, TD), 1)));
    attention_kernel<<<dim3(TH, T), 128, 0, st>>>(QKV, mask, pos_bias, ATT, S);
    HVLA_LAUNCH_CHECK("t5_attention");
    {
      GemmP g = gp(ATT, TD, w + L::wo, TD, X, TD, M, TD, TD);
      g.R = X; g.ldr = TD;
      HVLA_TRY((gemm_simt<float, float, float, float>(st, g, 1)));
    }
    rmsnorm768_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, w + L::ln1, Y, M);
    HVLA_LAUNCH_CHECK("t5_rmsnorm");
    {
      GemmP g = gp(Y, TD, w + L::wi, TD, HID, TFF, M, TFF, TD);
      g.act = 3;                                            // ReLU
      HVLA_TRY((gemm_simt<float, float, float, float>(st, g, 1)));
    }
    {
      GemmP g = gp(HID, TFF, w + L::wo2, TFF, X, TD, M, TD, TFF);
      g.R = X; g.ldr = TD;
      HVLA_TRY((gemm_simt<float, float, float, float>(st, g, 1)));
    }
  }
  rmsnorm768_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, blob + L::lnf, out, M);
  HVLA_LAUNCH_CHECK("t5_rmsnorm");
  return HVLA_OK;
}

// ---- tensor-core path ("bf16x3") ---------------------------------------------------------------------------------
// The fp32 path above is exact but runs its GEMMs on CUDA cores (6.7 ms for one instruction, 21 ms for 64: the skinny
// M = 32..2048 GEMMs reach a few % of anything).  Here every GEMM runs on the tcgen05 kernel of the DINOv2 blocks with
// fp32-like accuracy: activations and weights are split into two bf16 numbers (x = hi + lo, hi = bf16(x), lo = bf16(x - hi),
// 16 mantissa bits together) and  A W^T ~= Ahi Whi^T + Alo Whi^T + Ahi Wlo^T  (the dropped Alo Wlo^T term is 2^-18 relative).
// The three products are ONE launch: the operands are concatenated along K,  A' = [Ahi | Alo | Ahi]  (M x 3K, written by the
// producing kernel) and  W' = [Whi | Whi | Wlo]  (N x 3K, packed once by hvla/t5.py), so the large hi.hi part is accumulated
// first and the corrections after it, in the fp32 TMEM accumulator.  Outputs go through the TMA reduce-add epilogue
// (EPI_RESIDUAL_F32, LayerScale 1, zero bias) into the fp32 stream, or into a zeroed fp32 buffer for q|k|v and the MLP hidden
// layer.  Results are run-to-run deterministic (K splits at small row counts are folded in a fixed order).
// Rows are padded to a multiple of 256 (one CTA-pair tile); padding rows are zero / never read back.
struct MatLayout {   // bf16 elements per layer, each matrix [N, 3K] = [hi | hi | lo] per row (HF torch [out,in] orientation)
  static constexpr int64_t wqkv = 0, wo = wqkv + (int64_t)3 * TD * 3 * TD, wi = wo + (int64_t)TD * 3 * TD, wo2 = wi + (int64_t)TFF * 3 * TD,
                           layer_size = wo2 + (int64_t)TD * 3 * TFF;
  static constexpr int64_t total = TL * layer_size;
};

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// four consecutive columns c..c+3 of row `row` of A' [*, 3K]: hi at c and 2K + c, lo at K + c
__device__ __forceinline__ void store_split4(bf16* __restrict__ a, int64_t row, int c, int K, const float y[4]) {
  __align__(8) bf16 h[4];
  __align__(8) bf16 l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16(y[j], h[j], l[j]);
  bf16* r = a + row * (3 * (int64_t)K) + c;
  *reinterpret_cast<uint2*>(r) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(r + K) = *reinterpret_cast<const uint2*>(l);
  *reinterpret_cast<uint2*>(r + 2 * K) = *reinterpret_cast<const uint2*>(h);
}

// T5LayerNorm straight into the split operand: one warp per row
__global__ void __launch_bounds__(256) rmsnorm768_split_kernel(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ a,
                                                               int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * TD);
  float4 v[6];
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = xr[lane + 32 * i];
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const float r = 1.0f / sqrtf(s2 / (float)TD + 1e-6f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    const float y[4] = {v[i].x * r * g.x, v[i].y * r * g.y, v[i].z * r * g.z, v[i].w * r * g.w};
    store_split4(a, row, (lane + 32 * i) * 4, TD, y);
  }
}

// fp32 [rows, K] -> A' [rows, 3K], optionally through ReLU; K % 4 == 0
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ in, bf16* __restrict__ a, int64_t n4, int K, int relu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  float y[4] = {v.x, v.y, v.z, v.w};
  if (relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = fmaxf(y[j], 0.f);
  }
  const int k4 = K / 4;
  store_split4(a, i / k4, (int)(i % k4) * 4, K, y);
}

constexpr size_t PART_BYTES_TC = (size_t)24 << 20;   // split-K scratch: covers every (rows, N, splits) the launcher can choose (<= 16.5 MB)
inline int padded_rows(int M) { return (M + 255) / 256 * 256; }
inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }
inline size_t workspace_bytes_tc(int T, int S) {
  const size_t Mp = (size_t)padded_rows(T * S);
  return 2 * al256(Mp * TD * 4) + al256(Mp * 3 * TD * 4) + al256(Mp * TFF * 4) + al256(Mp * 3 * TFF * 2) + al256((size_t)TFF * 4) + PART_BYTES_TC;
}

// out[i] += part[0][i] + part[1][i] + ... in a fixed order (the K splits of the preceding GEMM); real rows only
__global__ void __launch_bounds__(256) fold_kernel(float* __restrict__ out, const float* __restrict__ part, int64_t stride4, int ns, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(out)[i];
  for (int s = 0; s < ns; ++s) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(part) + s * stride4 + i);
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  }
  reinterpret_cast<float4*>(out)[i] = a;
}


// OUT[Mp,N] (fp32) += A'[Mp,3K] W'[N,3K]^T.  With few instructions the N = 768 GEMMs have 3 output tiles and 3 CTA pairs would
// stream a whole 14 MB matrix (45 us at one instruction): the launcher then splits K over up to 8 CTA pairs (the mechanism of
// the DINOv2 residual GEMMs at batch 1: split 0 reduce-adds into OUT, the others store partial products) and fold_kernel adds
// the partial products in a fixed order -- still deterministic.  HVLA_GEMM_SPLITK=0 turns it off.
inline int gemm_split(cudaStream_t st, const bf16* A3, const bf16* W3, float* OUT, int Mp, int M, int N, int K, const float* zero_bias,
                      float* part) {
  tc::EpiP ep;
  memset(&ep, 0, sizeof ep);
  int splits = 1;
  ep.bias = zero_bias; ep.out = OUT; ep.ldo = N;            // ls == null: LayerScale 1
  ep.part = part; ep.part_bytes = PART_BYTES_TC; ep.splits_used = &splits;
  HVLA_TRY(tc2::gemm_tc2(st, A3, W3, Mp, N, 3 * K, tc::EPI_RESIDUAL_F32, ep));
  if (splits > 1) {
    const int64_t n4 = (int64_t)M * N / 4;
    fold_kernel<<<cdiv(n4, 256), 256, 0, st>>>(OUT, part, (int64_t)Mp * N / 4, splits - 1, n4);
    HVLA_LAUNCH_CHECK("t5_fold");
  }
  return HVLA_OK;
}

inline int encode_tc(cudaStream_t st, const float* blob, const bf16* mat, const float* pos_bias, const int32_t* ids, const int32_t* mask, int T,
                     int S, float* out, uint8_t* ws) {
  typedef Layout L;
  typedef MatLayout W;
  const int M = T * S, Mp = padded_rows(M);
  uint8_t* p = ws;
  float* X = reinterpret_cast<float*>(p);   p += al256((size_t)Mp * TD * 4);
  float* ATT = reinterpret_cast<float*>(p); p += al256((size_t)Mp * TD * 4);
  float* QKV = reinterpret_cast<float*>(p); p += al256((size_t)Mp * 3 * TD * 4);
  float* HID = reinterpret_cast<float*>(p); p += al256((size_t)Mp * TFF * 4);
  bf16* A3 = reinterpret_cast<bf16*>(p);    p += al256((size_t)Mp * 3 * TFF * 2);
  float* ZERO = reinterpret_cast<float*>(p); p += al256((size_t)TFF * 4);
  float* PART = reinterpret_cast<float*>(p);
  ProfScope ps(st, "t5_encode");
  HVLA_CUDA(cudaMemsetAsync(ws, 0, workspace_bytes_tc(T, S) - PART_BYTES_TC, st));     // padding rows, zero bias
  gather_kernel<<<M, 192, 0, st>>>(ids, blob + L::embed, X, M);
  HVLA_LAUNCH_CHECK("t5_gather");
  for (int l = 0; l < TL; ++l) {
    const float* w = blob + L::layers + (int64_t)l * L::layer_size;
    const bf16* m = mat + (int64_t)l * W::layer_size;
    rmsnorm768_split_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, w + L::ln0, A3, M);
    HVLA_LAUNCH_CHECK("t5_rmsnorm_split");
    if (l) HVLA_CUDA(cudaMemsetAsync(QKV, 0, (size_t)Mp * 3 * TD * 4, st));
    HVLA_TRY(gemm_split(st, A3, m + W::wqkv, QKV, Mp, M, 3 * TD, TD, ZERO, PART));
    attention_kernel<<<dim3(TH, T), 128, 0, st>>>(QKV, mask, pos_bias, ATT, S);
    HVLA_LAUNCH_CHECK("t5_attention");
    split_kernel<<<cdiv((int64_t)M * TD / 4, 256), 256, 0, st>>>(ATT, A3, (int64_t)M * TD / 4, TD, 0);
    HVLA_LAUNCH_CHECK("t5_split");
    HVLA_TRY(gemm_split(st, A3, m + W::wo, X, Mp, M, TD, TD, ZERO, PART));                 // x += att Wo^T
    rmsnorm768_split_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, w + L::ln1, A3, M);
    HVLA_LAUNCH_CHECK("t5_rmsnorm_split");
    if (l) HVLA_CUDA(cudaMemsetAsync(HID, 0, (size_t)Mp * TFF * 4, st));
    HVLA_TRY(gemm_split(st, A3, m + W::wi, HID, Mp, M, TFF, TD, ZERO, PART));
    split_kernel<<<cdiv((int64_t)M * TFF / 4, 256), 256, 0, st>>>(HID, A3, (int64_t)M * TFF / 4, TFF, 1);   // ReLU, then split
    HVLA_LAUNCH_CHECK("t5_split");
    HVLA_TRY(gemm_split(st, A3, m + W::wo2, X, Mp, M, TD, TFF, ZERO, PART));               // x += relu(.) Wo2^T
  }
  rmsnorm768_kernel<<<cdiv(M, 8), 256, 0, st>>>(X, blob + L::lnf, out, M);
  HVLA_LAUNCH_CHECK("t5_rmsnorm");
  return HVLA_OK;
}

}  // namespace t5
}  // namespace hvla
