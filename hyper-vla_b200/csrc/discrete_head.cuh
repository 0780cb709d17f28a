// DiscreteActionHead + BinTokenizer.decode on the GPU (second half):
//   hypervla/components/action_heads.py:252-396 (vocab_proj Dense -> logits (horizon, action_dim, 256) -> argmax) and
//   octo/model/components/tokenizers.py:235-275 (uniform bins over [-1, 1]: decode = bin centre).
// The base ViT then carries A readout tokens instead of one (base_network.py:22-33: discrete_token_type "action_horizon" -> A = 4,
// each token projects to action_dim * 256 logits; "action_dim_and_action_horizon" -> A = 28, each token projects to 256 logits);
// patches do not attend to ANY of them, the readout tokens attend to everything (base_vit.py:209-214).
// Generated row of this variant: proj_w[768,64] proj_b[64] pos[256+A,64] 4 x block (as GenLayout) encn_s encn_b vocab_w[64,V] vocab_b[V].
#pragma once
#include "common.cuh"

namespace hvla {

constexpr int VOCAB = 256;

// Shape of one generated row and of the base-net token sequence (runtime: the mix head is A = 1 with GenLayout's offsets)
struct BaseShape {
  int A;                 // readout (action) tokens
  int S;                 // 256 + A
  int V;                 // discrete head: outputs per token (1792 or 256); 0 = mix head
  int64_t layers, encn_s, encn_b, head_w, head_b, total, ngp;
};
inline BaseShape mix_shape() {
  BaseShape s;
  s.A = 1; s.S = BTOK; s.V = 0;
  s.layers = GenLayout::layers; s.encn_s = GenLayout::encn_s; s.encn_b = GenLayout::encn_b;
  s.head_w = GenLayout::wc; s.head_b = GenLayout::bc; s.total = GenLayout::total; s.ngp = NGP;
  return s;
}
inline bool discrete_shape(int A, BaseShape* out) {
  if (A != AH && A != AH * AD) return false;
  BaseShape s;
  s.A = A; s.S = NPATCH + A; s.V = A == AH ? AD * VOCAB : VOCAB;
  s.layers = GenLayout::pos + (int64_t)s.S * BD;
  s.encn_s = s.layers + BL * GenLayout::layer_size;
  s.encn_b = s.encn_s + BD;
  s.head_w = s.encn_b + BD;
  s.head_b = s.head_w + (int64_t)BD * s.V;
  s.total = s.head_b + s.V;
  s.ngp = (s.total + 31) / 32 * 32;
  *out = s;
  return true;
}

// encoder_norm of the A readout tokens, vocab_proj, argmax over the 256 bins of every (horizon, dim), bin centre.
// One CTA per env, one warp per (horizon, dim) slot in turn; lane l scans bins l, l+32, ...  Ties resolve to the LOWEST bin index
// (jnp.argmax).  tokens [B,4,7] i32, action [B,4,7] f32, top2 [B,4,7,2] f32 (the two largest logits, for margin-aware parity tests) or null.
template <typename TW>
__global__ void __launch_bounds__(256) discrete_head_kernel(const float* __restrict__ X /*[B,S,64]*/, const TW* __restrict__ weights,
                                                            const int* __restrict__ tidx, BaseShape sh, int32_t* __restrict__ tokens,
                                                            float* __restrict__ action, float* __restrict__ top2) {
  __shared__ float hs[AH * AD][BD];                      // normalised readout tokens (A <= 28)
  const int b = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TW* wr = weights + (int64_t)(tidx ? tidx[b] : b) * sh.ngp;
  for (int t = w; t < sh.A; t += 8) {
    const float* x = X + ((int64_t)b * sh.S + NPATCH + t) * BD;
    const float v0 = x[lane], v1 = x[lane + 32];
    float s = v0 + v1, s2 = fmaf(v0, v0, v1 * v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s / 64.f, var = fmaxf(0.f, s2 / 64.f - mean * mean);
    const float rstd = 1.0f / sqrtf(var + 1e-6f);
    hs[t][lane] = (v0 - mean) * (rstd * to_f(wr[sh.encn_s + lane])) + to_f(wr[sh.encn_b + lane]);
    hs[t][lane + 32] = (v1 - mean) * (rstd * to_f(wr[sh.encn_s + lane + 32])) + to_f(wr[sh.encn_b + lane + 32]);
  }
  __syncthreads();
  for (int slot = w; slot < AH * AD; slot += 8) {          // slot = horizon * 7 + dim
    const int tok = sh.A == AH ? slot / AD : slot;         // which readout token holds this slot's logits
    const int col0 = sh.A == AH ? (slot % AD) * VOCAB : 0;
    const TW* W = wr + sh.head_w + col0;
    float best = -FLT_MAX, second = -FLT_MAX;
    int arg = 0;
    for (int j = 0; j < VOCAB / 32; ++j) {
      const int bin = lane + 32 * j;
      float a = 0.f;
      for (int k = 0; k < BD; ++k) a = fmaf(hs[tok][k], to_f(W[(int64_t)k * sh.V + bin]), a);
      a += to_f(wr[sh.head_b + col0 + bin]);
      if (a > best) { second = best; best = a; arg = bin; }
      else if (a > second) second = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { second = fmaxf(best, os); best = ob; arg = oa; }
      else second = fmaxf(second, ob);
    }
    if (lane == 0) {
      tokens[(int64_t)b * (AH * AD) + slot] = arg;
      // BinTokenizer.decode (uniform): thresholds = linspace(-1, 1, 257); centre of bin i
      action[(int64_t)b * (AH * AD) + slot] = -1.0f + (2.0f * (float)arg + 1.0f) / (float)VOCAB;
      if (top2) { top2[((int64_t)b * (AH * AD) + slot) * 2] = best; top2[((int64_t)b * (AH * AD) + slot) * 2 + 1] = second; }
    }
  }
}

}  // namespace hvla
