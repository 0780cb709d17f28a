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

// one warp per (query i, head h, task t); lane = key.  s = q.k (no 1/sqrt(d)) + pos_bias[h,i,j] + (1-mask_j)*finfo.min
__global__ void __launch_bounds__(32) attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ mask,
                                                       const float* __restrict__ pos_bias, float* __restrict__ out, int S) {
  const int i = blockIdx.x, h = blockIdx.y, t = blockIdx.z, lane = threadIdx.x;
  const float* q = qkv + ((int64_t)t * S + i) * (3 * TD) + h * THD;
  float s = -INFINITY;
  if (lane < S) {
    const float* k = qkv + ((int64_t)t * S + lane) * (3 * TD) + TD + h * THD;
    float a = 0.f;
#pragma unroll 8
    for (int d = 0; d < THD; ++d) a = fmaf(__ldg(q + d), __ldg(k + d), a);
    s = a + pos_bias[((int64_t)h * S + i) * S + lane] + (mask[t * S + lane] != 0 ? 0.f : -FLT_MAX);
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
    const float* v = qkv + ((int64_t)t * S + j) * (3 * TD) + 2 * TD + h * THD;
    o0 = fmaf(pj, __ldg(v + lane), o0);
    o1 = fmaf(pj, __ldg(v + lane + 32), o1);
  }
  float* o = out + ((int64_t)t * S + i) * TD + h * THD;
  o[lane] = o0;
  o[lane + 32] = o1;
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
    attention_kernel<<<dim3(S, TH, T), 32, 0, st>>>(QKV, mask, pos_bias, ATT, S);
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

}  // namespace t5
}  // namespace hvla
