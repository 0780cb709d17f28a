// libhvla.so -- C ABI + host orchestration of the HyperVLA hot path on sm_100a.
// See include/hvla.h for the contract and the reference code each entry point replaces.
#include "common.cuh"
#include "simt_kernels.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "gemm_chain.cuh"
#include "attn_mma.cuh"
#include "attn_tc.cuh"
#include "base_fused.cuh"
#include "heads_mma.cuh"
#include "ctx_fused.cuh"
#include "postprocess.cuh"
#include "preprocess.cuh"
#include "t5_embed.cuh"
#include "dino_x3.cuh"
#include "discrete_head.cuh"

#include <stdlib.h>
#include <type_traits>

namespace hvla {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
ProfState g_prof;
bool g_pdl = getenv("HVLA_NO_PDL") == nullptr;

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- workspace plan ---------------------------------------------------------------------------------
struct Plan {
  // generate region (fp32)
  size_t tp, ip, xc, yc, qkvc, cc, hc, e;
  // DINO region
  size_t x, y, qkv, att, hid, emb, part, st, chain;
  // base region (fp32)
  size_t pt, xb, yb, qkvb, cb, hb;
  // host staging (hvla_act_host)
  size_t img, act, logit, tidx;
  // split-operand flow (HVLA_BF16X3, dino_x3.cuh): A' of the 768-wide inputs, [hi | lo] of q|k|v, A' of the MLP hidden layer
  size_t a3, qkv2, a3l;
  size_t total;
};

// room for the split-K partial products of one residual GEMM: at most one 256x256 fp32 block per CTA pair
constexpr size_t PART_BYTES = (size_t)74 * 256 * 256 * 4;

static Plan make_plan(int B, int T, int dtype) {
  Plan p;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const size_t es = dtype == HVLA_BF16 ? 2 : 4;
  const size_t Tt = (size_t)(T > 0 ? T : 1), Bb = (size_t)(B > 0 ? B : 1);
  p.tp = take(Tt * LANG * CD * 4);
  p.ip = take(Tt * CD * 4);
  p.xc = take(Tt * CTOK * CD * 4);
  p.yc = take(Tt * CTOK * CD * 4);
  p.qkvc = take(Tt * CTOK * 3 * CD * 4);
  p.cc = take(Tt * CTOK * CD * 4);
  p.hc = take(Tt * CTOK * CF * 4);
  p.e = take(Tt * CD * 4);
  const size_t M = Bb * DTOK;
  p.x = take((M + 31) / 32 * 32 * DD * 4);                 // fp32 residual stream (flow B: blocked layout, whole 32-row blocks)
  p.y = take(M * DD * es);
  p.qkv = take(M * 3 * DD * es);
  p.att = take(M * DD * es);
  p.hid = take(M * DF * es);
  p.emb = take(M * DD * es);
  p.part = take(dtype == HVLA_BF16 ? PART_BYTES : 0);   // split-K partial products (small batches)
  p.st = take(dtype == HVLA_BF16 ? M * 12 * 4 : 0);      // partial row statistics of the stream (LayerNorm-free flow)
  p.chain = take(dtype == HVLA_BF16 ? (size_t)DL * chain::chain_ws_ints((int64_t)M) * 4 : 0);   // scheduler state of the GEMM chains (gemm_chain.cuh)
  p.pt = take(Bb * NPATCH * BD * 4);
  constexpr size_t BTOKMAX = NPATCH + AH * AD;     // discrete head: up to 28 readout tokens (discrete_head.cuh)
  p.xb = take(Bb * BTOKMAX * BD * 4);
  p.yb = take(Bb * BTOKMAX * BD * 4);
  p.qkvb = take(Bb * BTOKMAX * 3 * BD * 4);
  p.cb = take(Bb * BTOKMAX * BD * 4);
  p.hb = take(Bb * BTOKMAX * BF * 4);
  p.img = take(Bb * IMG * IMG * 3);
  p.act = take(Bb * AH * AD * 4);
  p.logit = take(Bb * AH * 4);
  p.tidx = take(Bb * 4);
  const bool x3f = dtype == HVLA_BF16X3;
  p.a3 = take(x3f ? M * 3 * DD * 2 : 0);
  p.qkv2 = take(x3f ? M * 2 * 3 * DD * 2 : 0);
  p.a3l = take(x3f ? M * 3 * DF * 2 : 0);
  p.total = off;
  return p;
}

// debugging aid: the tensor-core GEMM's math on CUDA cores (same bf16 operands, W transposed)
static int gemm_simt_debug(cudaStream_t st, const bf16* A, const bf16* Wt, int M, int N, int K, int epi, const tc::EpiP& ep) {
  if (epi == tc::EPI_BIAS_BF16) {
    bf16* C = reinterpret_cast<bf16*>(ep.out);
    const int nq = ep.qcols;
    if (nq > 0) {
      GemmP g = gemm_params(A, K, Wt, K, ep.bias, C, ep.ldo, M, nq, K);
      g.wt = 1; g.out_scale = ep.qscale;
      HVLA_TRY((gemm_simt<bf16, bf16, float, bf16>(st, g, 1)));
    }
    if (N > nq) {
      GemmP g = gemm_params(A, K, Wt + (int64_t)nq * K, K, ep.bias + nq, C + nq, ep.ldo, M, N - nq, K);
      g.wt = 1;
      HVLA_TRY((gemm_simt<bf16, bf16, float, bf16>(st, g, 1)));
    }
    return HVLA_OK;
  }
  if (epi == tc::EPI_BIAS_GELU_BF16) {
    GemmP g = gemm_params(A, K, Wt, K, ep.bias, ep.out, ep.ldo, M, N, K);
    g.wt = 1; g.act = 2;
    return gemm_simt<bf16, bf16, float, bf16>(st, g, 1);
  }
  if (epi == tc::EPI_RESIDUAL_F32) {
    GemmP g = gemm_params(A, K, Wt, K, ep.bias, ep.out, ep.ldo, M, N, K);
    g.wt = 1; g.R = reinterpret_cast<const float*>(ep.out); g.ldr = ep.ldo; g.ls = ep.ls;
    return gemm_simt<bf16, bf16, float, float>(st, g, 1);
  }
  return fail(HVLA_ERR_ARG, "gemm_simt_debug: unsupported epilogue");
}

static bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

// ---- generate (K1-K3) ---------------------------------------------------------------------------------
// dense [T, 128] context rows -> rows[t] of a persistent [T_max, 128] buffer
__global__ void scatter_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ rows, int T, int T_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * CD) return;
  const int r = rows[i / CD];
  if (r >= 0 && r < T_max) dst[(int64_t)r * CD + i % CD] = src[i];
}

template <typename TW>
static int generate_impl(cudaStream_t st, const float* hn, const void* hn_lp, const void* heads_w, const float* heads_b, const float* tok_emb,
                         const int32_t* tok_mask, const uint8_t* lang_pad, const float* init_cls, int T, void* out_w,
                         float* out_ctx, uint8_t* ws, const Plan& pl, const int32_t* rows = nullptr, int T_max = 0, int64_t ngp = NGP) {
  float* TPj = reinterpret_cast<float*>(ws + pl.tp);
  float* IPj = reinterpret_cast<float*>(ws + pl.ip);
  float* X = reinterpret_cast<float*>(ws + pl.xc);
  float* Y = reinterpret_cast<float*>(ws + pl.yc);
  float* QKV = reinterpret_cast<float*>(ws + pl.qkvc);
  float* CC = reinterpret_cast<float*>(ws + pl.cc);
  float* HC = reinterpret_cast<float*>(ws + pl.hc);
  // rows != null (task-switch scheduler): weights / context rows are scattered into persistent [T_max, .] buffers
  float* E = (out_ctx && !rows) ? out_ctx : reinterpret_cast<float*>(ws + pl.e);
  if (!rows) T_max = T;
  auto scatter_ctx = [&]() -> int {
    if (!rows || !out_ctx) return HVLA_OK;
    scatter_rows_kernel<<<cdiv((int64_t)T * CD, 256), 256, 0, st>>>(E, out_ctx, rows, T, T_max);
    HVLA_LAUNCH_CHECK("scatter_ctx");
    return HVLA_OK;
  };
  const int M = T * CTOK;
  typedef HnLayout L;
  if (std::is_same<TW, bf16>::value && hn_lp && !env_flag("HVLA_DEBUG_GENERIC_CTX")) {
    // bf16 tensor-core path: the whole context encoder is one kernel, the 73 heads another
    HVLA_TRY(ctxf::ctx_encode_lp(st, hn, reinterpret_cast<const ctxf::lp*>(hn_lp), tok_emb, tok_mask, lang_pad, init_cls, T, E));
    HVLA_TRY(scatter_ctx());
    return heads::heads_gemm_bf16(st, E, reinterpret_cast<const bf16*>(heads_w), heads_b, reinterpret_cast<bf16*>(out_w), T, rows, T_max, ngp);
  }
  // K1: projections (hypernetwork.py:112, 126)
  HVLA_TRY((gemm_simt<float, float, float, float>(
      st, gemm_params(tok_emb, LANGD, hn + L::tok_w, CD, hn + L::tok_b, TPj, CD, T * LANG, CD, LANGD), 1)));
  HVLA_TRY((gemm_simt<float, float, float, float>(
      st, gemm_params(init_cls, DD, hn + L::img_w, CD, hn + L::img_b, IPj, CD, T, CD, DD), 1)));
  {
    const int64_t total = (int64_t)M * CD;
    ProfScope ps(st, "ctx_assemble");
    ctx_assemble_kernel<<<cdiv(total, 256), 256, 0, st>>>(TPj, IPj, hn + L::task_pos, hn + L::img_pos, hn + L::layer_pos, X, T);
    HVLA_LAUNCH_CHECK("ctx_assemble");
  }
  // K2: context encoder (transformer.py:247-260)
  for (int l = 0; l < CL; ++l) {
    const float* lw = hn + L::layers + (int64_t)l * L::layer_size;
    LnP ln; memset(&ln, 0, sizeof ln);
    ln.x = X; ln.ldx = CD; ln.y = Y; ln.ldy = CD; ln.scale = lw + L::ln0_s; ln.bias = lw + L::ln0_b; ln.rows = M; ln.rows_per_batch = 1;
    HVLA_TRY((layernorm<float, float>(st, ln, CD)));
    HVLA_TRY((gemm_simt<float, float, float, float>(st, gemm_params(Y, CD, lw + L::wqkv, 3 * CD, lw + L::bqkv, QKV, 3 * CD, M, 3 * CD, CD), 1)));
    AttnP ap; memset(&ap, 0, sizeof ap);
    ap.qkv = QKV; ap.out = CC; ap.S = CTOK; ap.H = CH; ap.nbatch = T; ap.mask = 2; ap.tok_mask = tok_mask; ap.lang_pad = lang_pad;
    HVLA_TRY((attention_simt<float, float>(st, ap, CHD)));
    GemmP g = gemm_params(CC, CD, lw + L::wo, CD, lw + L::bo, X, CD, M, CD, CD);
    g.R = X; g.ldr = CD;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g, 1)));
    ln.scale = lw + L::ln1_s; ln.bias = lw + L::ln1_b;
    HVLA_TRY((layernorm<float, float>(st, ln, CD)));
    GemmP g0 = gemm_params(Y, CD, lw + L::w0, CF, lw + L::b0, HC, CF, M, CF, CD);
    g0.act = 1;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g0, 1)));
    GemmP g1 = gemm_params(HC, CF, lw + L::w1, CD, lw + L::b1, X, CD, M, CD, CF);
    g1.R = X; g1.ldr = CD;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g1, 1)));
  }
  {  // encoder_norm on the layer token only, then / sqrt(128)  (hypernetwork.py:188-192)
    LnP ln; memset(&ln, 0, sizeof ln);
    ln.x = X + (int64_t)(CTOK - 1) * CD; ln.ldx = (int64_t)CTOK * CD; ln.y = E; ln.ldy = CD;
    ln.scale = hn + L::encn_s; ln.bias = hn + L::encn_b; ln.rows = T; ln.rows_per_batch = 1;
    ln.post_div = sqrtf((float)CD);
    HVLA_TRY((layernorm<float, float>(st, ln, CD)));
  }
  // K3: all 73 output heads as one skinny GEMM (hypernetwork.py:205-217, 227)
  HVLA_TRY(scatter_ctx());
  if (std::is_same<TW, bf16>::value && !env_flag("HVLA_DEBUG_SIMT_HEADS"))
    return heads::heads_gemm_bf16(st, E, reinterpret_cast<const bf16*>(heads_w), heads_b, reinterpret_cast<bf16*>(out_w), T, rows, T_max, ngp);
  ProfScope ps(st, "heads_gemm");
  heads_gemm_kernel<TW, TW><<<cdiv(ngp, 1024), 256, 0, st>>>(E, reinterpret_cast<const TW*>(heads_w), heads_b,
                                                             reinterpret_cast<TW*>(out_w), T, rows, T_max, ngp);
  HVLA_LAUNCH_CHECK("heads_gemm");
  return HVLA_OK;
}

// ---- DINOv2 (K4-K7) -----------------------------------------------------------------------------------
static int dino_f32(cudaStream_t st, const float* dv, const float* dm, const uint8_t* images, int B, float* out_emb,
                    uint8_t* ws, const Plan& pl, float* maps = nullptr) {
  typedef DvecLayout V;
  typedef DmatLayout Mx;
  const int M = B * DTOK;
  float* X = reinterpret_cast<float*>(ws + pl.x);
  float* Y = reinterpret_cast<float*>(ws + pl.y);
  float* QKV = reinterpret_cast<float*>(ws + pl.qkv);
  float* ATT = reinterpret_cast<float*>(ws + pl.att);
  float* HID = reinterpret_cast<float*>(ws + pl.hid);
  float* A0 = HID;   // im2col matrix aliases the MLP hidden buffer
  float* P = QKV;    // patch projections alias the qkv buffer
  {
    const int64_t total = (int64_t)B * NPATCH * PATCH_KP;
    ProfScope ps(st, "im2col");
    im2col_norm_kernel<float><<<cdiv(total, 256), 256, 0, st>>>(images, A0, B);
    HVLA_LAUNCH_CHECK("im2col");
  }
  HVLA_TRY((gemm_simt<float, float, float, float>(
      st, gemm_params(A0, PATCH_KP, dm + Mx::patch_w, DD, dv + V::patch_b, P, DD, B * NPATCH, DD, PATCH_KP), 1)));
  {
    const int64_t total = (int64_t)M * DD;
    dino_assemble_kernel<float><<<cdiv(total, 256), 256, 0, st>>>(P, dv + V::cls, dv + V::pos, X, B);
    HVLA_LAUNCH_CHECK("dino_assemble");
  }
  for (int l = 0; l < DL; ++l) {
    const float* v = dv + V::layers + (int64_t)l * V::layer_size;
    const float* m = dm + Mx::layers + (int64_t)l * Mx::layer_size;
    LnP ln; memset(&ln, 0, sizeof ln);
    ln.x = X; ln.ldx = DD; ln.y = Y; ln.ldy = DD; ln.scale = v + V::ln1_s; ln.bias = v + V::ln1_b; ln.rows = M; ln.rows_per_batch = 1;
    HVLA_TRY((layernorm<float, float>(st, ln, DD)));
    HVLA_TRY((gemm_simt<float, float, float, float>(st, gemm_params(Y, DD, m + Mx::wqkv, 3 * DD, v + V::bqkv, QKV, 3 * DD, M, 3 * DD, DD), 1)));
    if (maps) HVLA_TRY((attn_probs<float, DHD>(st, QKV, maps + (int64_t)l * B * DH * DTOK * DTOK, DTOK, DH, B, 0, 0.125f)));
    AttnP ap; memset(&ap, 0, sizeof ap);
    ap.qkv = QKV; ap.out = ATT; ap.S = DTOK; ap.H = DH; ap.nbatch = B; ap.mask = 0;
    HVLA_TRY((attention_simt<float, float>(st, ap, DHD)));
    GemmP g = gemm_params(ATT, DD, m + Mx::wo, DD, v + V::bo, X, DD, M, DD, DD);
    g.R = X; g.ldr = DD; g.ls = v + V::ls1;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g, 1)));
    ln.scale = v + V::ln2_s; ln.bias = v + V::ln2_b;
    HVLA_TRY((layernorm<float, float>(st, ln, DD)));
    GemmP g1 = gemm_params(Y, DD, m + Mx::w1, DF, v + V::b1, HID, DF, M, DF, DD);
    g1.act = 2;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g1, 1)));
    GemmP g2 = gemm_params(HID, DF, m + Mx::w2, DD, v + V::b2, X, DD, M, DD, DF);
    g2.R = X; g2.ldr = DD; g2.ls = v + V::ls2;
    HVLA_TRY((gemm_simt<float, float, float, float>(st, g2, 1)));
  }
  LnP ln; memset(&ln, 0, sizeof ln);
  ln.x = X; ln.ldx = DD; ln.y = out_emb; ln.ldy = DD; ln.scale = dv + V::lnf_s; ln.bias = dv + V::lnf_b; ln.rows = M; ln.rows_per_batch = 1;
  HVLA_TRY((layernorm<float, float>(st, ln, DD)));
  return HVLA_OK;
}

// cls rows of the residual stream: X[b,0,:] = cls + pos[0]
__global__ void dino_cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ X, int B) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * DD) return;
  const int c = idx % DD, b = idx / DD;
  X[(int64_t)b * DTOK * DD + c] = cls[c] + pos[c];
}

// whole residual stream before the patch-embedding GEMM accumulates onto it: X[b,r,:] = pos[r,:] (+ cls on row 0)
__global__ void __launch_bounds__(256) dino_init_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ X,
                                                             int64_t total4) {
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

// cls rows of the BLOCKED stream (gemm_tc.cuh: xblk_f4): X[b,0,:] = cls + pos[0]; the patch rows are written by the patch-embedding GEMM
__global__ void __launch_bounds__(192) dino_cls_rows_blk_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ X,
                                                                bf16* __restrict__ Y, float* __restrict__ ST) {
  pdl_trigger();
  pdl_wait();
  const int c4 = threadIdx.x;
  const float4 c = __ldg(reinterpret_cast<const float4*>(cls) + c4), p = __ldg(reinterpret_cast<const float4*>(pos) + c4);
  const float4 v = make_float4(c.x + p.x, c.y + p.y, c.z + p.z, c.w + p.w);
  const int64_t row = (int64_t)blockIdx.x * DTOK;
  reinterpret_cast<float4*>(X)[tc::xblk_f4((int)row, c4)] = v;
  // bf16 shadow + row statistics of the CLS row for the first q|k|v (the patch rows get theirs from the patch-embedding epilogue)
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(Y + row * DD + 4 * c4) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  float s = (v.x + v.y) + (v.z + v.w), q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ float red[6][2];
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s; red[threadIdx.x >> 5][1] = q; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
      o.x = ((red[0][0] + red[1][0]) + (red[2][0] + red[3][0])) + (red[4][0] + red[5][0]);
      o.y = ((red[0][1] + red[1][1]) + (red[2][1] + red[3][1])) + (red[4][1] + red[5][1]);
    }
    reinterpret_cast<float4*>(ST + row * 12)[threadIdx.x] = o;
  }
}

// ---- large-batch flow ("flow B"): no LayerNorm kernels between the GEMMs (gemm_tc.cuh, "stream LayerNorm without a LayerNorm kernel")
// The fp32 residual stream is kept in the BLOCKED layout; the residual GEMMs (proj, fc2) update it in place in their epilogue and
// emit the un-normalised bf16 shadow (Y) + per-row statistics (ST) that q|k|v / fc1 fold into THEIR epilogue.  Per layer: 4 GEMMs
// + attention (or one GEMM chain + attention), nothing else; the shadow of the embedded tokens comes out of the patch-embedding epilogue
// and the CLS-row kernel, one stream_blk_rows launch per forward is the final LayerNorm.
static int dino_bf16_blk(cudaStream_t st, const float* dv, const bf16* dm, const uint8_t* images, int B, bf16* out_emb,
                         uint8_t* ws, const Plan& pl) {
  typedef DvecLayout V;
  typedef DmatLayout Mx;
  const int M = B * DTOK;
  float* X = reinterpret_cast<float*>(ws + pl.x);          // blocked, (M + 31) / 32 row blocks
  bf16* Y = reinterpret_cast<bf16*>(ws + pl.y);
  bf16* QKV = reinterpret_cast<bf16*>(ws + pl.qkv);
  bf16* ATT = reinterpret_cast<bf16*>(ws + pl.att);
  bf16* HID = reinterpret_cast<bf16*>(ws + pl.hid);
  bf16* A0 = HID;
  float* ST = reinterpret_cast<float*>(ws + pl.st);
  // L2 locality along the launch chain (gemm_tc.cuh, "serpentine tile order"): kernel k walks the row blocks forwards for even k and
  // backwards for odd k, so every consumer starts with the rows its producer wrote last; the residual stream can be accessed with an
  // evict_last policy (HVLA_XHINT=1).  Both are A/B switches whose measured effect is inside the noise (profiles/README.md).
  static const bool serp = !(getenv("HVLA_SERPENTINE") && getenv("HVLA_SERPENTINE")[0] == '0');
  static const int xh_env = getenv("HVLA_XHINT") ? atoi(getenv("HVLA_XHINT")) : -1;
  const int xhint = xh_env >= 0 ? xh_env : 0;      // measured with the GEMM chain: 2.87 ms per forward with the policy, 2.86 without -- off by default
  int chain = 0;
  auto next_rev = [&]() { return serp ? (chain++ & 1) : 0; };
  {
    ProfScope ps(st, "im2col");
    launch_k(im2col_norm_bf16_kernel, dim3(B * GRID), dim3(256), 0, st, images, A0, B);
    HVLA_LAUNCH_CHECK("im2col");
  }
  {
    ProfScope ps(st, "cls_rows");
    launch_k(dino_cls_rows_blk_kernel, dim3(B), dim3(DD / 4), 0, st, dv + V::cls, dv + V::pos, X, Y, ST);
    HVLA_LAUNCH_CHECK("dino_cls_rows_blk");
  }
  {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = dv + V::patch_b; ep.out = X; ep.ldo = DD; ep.rows = B * NPATCH; ep.pos = dv + V::pos_blk; ep.xhint = xhint;
    ep.shadow = Y; ep.stats_out = ST;        // shadow + statistics of the embedded tokens for the first q|k|v: no separate stream pass
    HVLA_TRY(tc2::gemm_tc2(st, A0, dm + Mx::patch_w, B * NPATCH, DD, PATCH_KP, tc::EPI_PATCH_BLK, ep));
  }
  auto ep_qkv = [&](const float* v, int rev) {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = v + V::bqkv_f; ep.out = QKV; ep.ldo = 3 * DD; ep.stats = ST; ep.cs = v + V::cs_qkv; ep.rev = rev;
    return ep;
  };
  auto ep_res = [&](const float* bias, const float* ls, int rev) {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = bias; ep.out = X; ep.ldo = DD; ep.ls = ls; ep.shadow = Y; ep.stats_out = ST; ep.rev = rev; ep.xhint = xhint;
    return ep;
  };
  auto ep_fc1 = [&](const float* v, int rev) {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = v + V::b1_f; ep.out = HID; ep.ldo = DF; ep.stats = ST; ep.cs = v + V::cs_1; ep.rev = rev;
    return ep;
  };
  // One launch per layer for proj -> fc1 -> fc2 -> q|k|v of the next layer (gemm_chain.cuh).
  // Default: while the N = 768 GEMMs have at most 12 waves of tiles (<= ~290 images) -- the chain buys back their tails and the launch
  // boundaries (64 / 128 images: +1.5 % sustained); from 20 waves on (512 images) the tails are noise and one launch per GEMM, with its
  // static tile order and cross-tile stream prefetch, is 2-5 % faster.  HVLA_CHAIN=0 / 1 force either.
  static const int chain_env = getenv("HVLA_CHAIN") ? atoi(getenv("HVLA_CHAIN")) : -1;
  const bool use_chain = chain_env >= 0 ? chain_env != 0 : (int64_t)((M + 255) / 256) * 3 <= (int64_t)12 * (num_sms() / 2);
  // wavefront lags in row blocks (proj -> fc1, fc1 -> fc2, fc2 -> q|k|v): HVLA_CHAIN_LAG="a,b,c" or one number for all three
  static int chain_lags[3] = {1 << 20, 1 << 20, 1 << 20};     // >= row blocks: each GEMM's tiles follow the previous GEMM's (measured: interleaving them is slower)
  static const bool lags_parsed = [&]() {
    if (const char* e = getenv("HVLA_CHAIN_LAG")) {
      int a = 0, b = 0, c = 0;
      const int n = sscanf(e, "%d,%d,%d", &a, &b, &c);
      if (n == 1) chain_lags[0] = chain_lags[1] = chain_lags[2] = a;
      else if (n == 3) { chain_lags[0] = a; chain_lags[1] = b; chain_lags[2] = c; }
    }
    return true;
  }();
  (void)lags_parsed;
  if (use_chain) {
    int* cws = reinterpret_cast<int*>(ws + pl.chain);
    const size_t cints = chain::chain_ws_ints(M);
    HVLA_CUDA(cudaMemsetAsync(cws, 0, (size_t)DL * cints * 4, st));
    const float* v0 = dv + V::layers;
    HVLA_TRY(tc2::gemm_tc2(st, Y, dm + Mx::layers + Mx::wqkv, M, 3 * DD, DD, tc::EPI_BIAS_BF16_FOLD, ep_qkv(v0, 0)));
    for (int l = 0; l < DL; ++l) {
      const float* v = dv + V::layers + (int64_t)l * V::layer_size;
      const bf16* m = dm + Mx::layers + (int64_t)l * Mx::layer_size;
      // the chain walks the row blocks upwards: q|k|v rows of the last images are the freshest in L2, so attention starts there,
      // and the attention rows of the first images are the freshest when the next chain starts
      HVLA_TRY(attn_tc::dino_attention_tc(st, QKV, ATT, B, serp ? 1 : 0));
      chain::ChainDesc d[4];
      d[0] = {ATT, m + Mx::wo, DD, DD, tc::EPI_RESIDUAL_BLK, ep_res(v + V::bo, v + V::ls1, 0)};
      d[1] = {Y, m + Mx::w1, DF, DD, tc::EPI_BIAS_GELU_BF16_FOLD, ep_fc1(v, 0)};
      d[2] = {HID, m + Mx::w2, DD, DF, tc::EPI_RESIDUAL_BLK, ep_res(v + V::b2, v + V::ls2, 0)};
      int n = 3;
      if (l + 1 < DL) {
        d[3] = {Y, m + Mx::layer_size + Mx::wqkv, 3 * DD, DD, tc::EPI_BIAS_BF16_FOLD, ep_qkv(v + V::layer_size, 0)};
        n = 4;
      }
      HVLA_TRY(chain::gemm_chain(st, d, n, M, cws + (size_t)l * cints, chain_lags));
    }
    return stream_blk_rows(st, X, out_emb, nullptr, dv + V::lnf_s, dv + V::lnf_b, M);
  }
  for (int l = 0; l < DL; ++l) {
    const float* v = dv + V::layers + (int64_t)l * V::layer_size;
    const bf16* m = dm + Mx::layers + (int64_t)l * Mx::layer_size;
    HVLA_TRY(tc2::gemm_tc2(st, Y, m + Mx::wqkv, M, 3 * DD, DD, tc::EPI_BIAS_BF16_FOLD, ep_qkv(v, next_rev())));
    HVLA_TRY(attn_tc::dino_attention_tc(st, QKV, ATT, B, next_rev()));
    HVLA_TRY(tc2::gemm_tc2(st, ATT, m + Mx::wo, M, DD, DD, tc::EPI_RESIDUAL_BLK, ep_res(v + V::bo, v + V::ls1, next_rev())));
    HVLA_TRY(tc2::gemm_tc2(st, Y, m + Mx::w1, M, DF, DD, tc::EPI_BIAS_GELU_BF16_FOLD, ep_fc1(v, next_rev())));
    HVLA_TRY(tc2::gemm_tc2(st, HID, m + Mx::w2, M, DD, DF, tc::EPI_RESIDUAL_BLK, ep_res(v + V::b2, v + V::ls2, next_rev())));
  }
  return stream_blk_rows(st, X, out_emb, nullptr, dv + V::lnf_s, dv + V::lnf_b, M);
}

// Flow B is used from this many images on (HVLA_FUSED_LN=0 / =1 force either flow): below it the N = 768 GEMMs have fewer tiles
// than CTA pairs, and the classic flow's split-K with partial products folded by the LayerNorm kernel is the faster one.
static bool use_flow_blk(int B) {
  const char* e = getenv("HVLA_FUSED_LN");
  if (e && e[0]) return e[0] != '0';
  return (int64_t)((B * DTOK + 255) / 256) * 3 >= num_sms() / 2;
}

static int dino_bf16(cudaStream_t st, const float* dv, const bf16* dm, const uint8_t* images, int B, bf16* out_emb,
                     uint8_t* ws, const Plan& pl, float* maps = nullptr) {
  if (!maps && use_flow_blk(B) && !env_flag("HVLA_DEBUG_SIMT_GEMM") && !env_flag("HVLA_DEBUG_SIMT_ATTN") && !env_flag("HVLA_ATTN_MMA") &&
      !env_flag("HVLA_GEMM_1CTA"))
    return dino_bf16_blk(st, dv, dm, images, B, out_emb, ws, pl);
  typedef DvecLayout V;
  typedef DmatLayout Mx;
  const int M = B * DTOK;
  float* X = reinterpret_cast<float*>(ws + pl.x);
  bf16* Y = reinterpret_cast<bf16*>(ws + pl.y);
  bf16* QKV = reinterpret_cast<bf16*>(ws + pl.qkv);
  bf16* ATT = reinterpret_cast<bf16*>(ws + pl.att);
  bf16* HID = reinterpret_cast<bf16*>(ws + pl.hid);
  bf16* A0 = HID;
  float* PART = reinterpret_cast<float*>(ws + pl.part);
  const int64_t part_stride = (int64_t)((M + 255) / 256) * 256 * DD;   // one [M_pad, 768] block per extra K split
  int splits = 1;                                                       // K splits of the most recent residual GEMM
  const bool simt_gemm = env_flag("HVLA_DEBUG_SIMT_GEMM");   // debugging aid: CUDA-core GEMMs on the bf16 data
  const bool simt_attn = env_flag("HVLA_DEBUG_SIMT_ATTN");
  const bool mma_attn = env_flag("HVLA_ATTN_MMA");           // A/B switch: warp-level mma.sync attention instead of tcgen05
  const bool one_cta = env_flag("HVLA_GEMM_1CTA");           // A/B switch: single-CTA 128x256 tiles instead of CTA pairs
  {
    {
      ProfScope ps(st, "im2col");
      launch_k(im2col_norm_bf16_kernel, dim3(B * GRID), dim3(256), 0, st, images, A0, B);   // one CTA per row of 16 patches
      HVLA_LAUNCH_CHECK("im2col");
    }
    ProfScope ps2(st, "cls_rows");
    if (one_cta || simt_gemm) {
      launch_k(dino_cls_rows_kernel, dim3(cdiv(B * DD, 256)), dim3(256), 0, st, dv + V::cls, dv + V::pos, X, B);
    } else {
      const int64_t total4 = (int64_t)M * DD / 4;
      launch_k(dino_init_rows_kernel, dim3(cdiv(total4, 256)), dim3(256), 0, st, dv + V::cls, dv + V::pos, X, total4);
    }
    HVLA_LAUNCH_CHECK("dino_cls_rows");
  }
  auto gemm = [&](const bf16* A, const bf16* Wt, int m, int n, int k, int epi, const tc::EpiP& ep) -> int {
    if (ep.splits_used) *ep.splits_used = 1;      // only the 2-CTA tensor-core path ever splits K
    if (!simt_gemm) return one_cta ? tc::gemm_tc(st, A, Wt, m, n, k, epi, ep) : tc2::gemm_tc2(st, A, Wt, m, n, k, epi, ep);
    // debug path: same math on CUDA cores (W given transposed)
    return gemm_simt_debug(st, A, Wt, m, n, k, epi, ep);
  };
  if (simt_gemm) {   // debug: patch projections to a bf16 temp (aliases QKV), then assemble tokens
    GemmP g = gemm_params(A0, PATCH_KP, dm + Mx::patch_w, PATCH_KP, dv + V::patch_b, QKV, DD, B * NPATCH, DD, PATCH_KP);
    g.wt = 1;
    HVLA_TRY((gemm_simt<bf16, bf16, float, bf16>(st, g, 1)));
    const int64_t total = (int64_t)M * DD;
    dino_assemble_kernel<bf16><<<cdiv(total, 256), 256, 0, st>>>(QKV, dv + V::cls, dv + V::pos, X, B);
    HVLA_LAUNCH_CHECK("dino_assemble");
  } else {
    tc::EpiP ep; memset(&ep, 0, sizeof ep);
    ep.bias = dv + V::patch_b; ep.out = X; ep.ldo = DD; ep.pos = dv + V::pos;
    ep.patch_rows = one_cta ? 0 : 1;     // 2-CTA path: the stream already holds cls/pos, the GEMM reduce-adds onto it
    HVLA_TRY(gemm(A0, dm + Mx::patch_w, B * NPATCH, DD, PATCH_KP, tc::EPI_PATCH_F32, ep));
  }
  for (int l = 0; l < DL; ++l) {
    const float* v = dv + V::layers + (int64_t)l * V::layer_size;
    const bf16* m = dm + Mx::layers + (int64_t)l * Mx::layer_size;
    LnP ln; memset(&ln, 0, sizeof ln);
    // gamma / beta of both LayerNorms are folded into wqkv / w1 and their biases (params.py: pack_dino_tree): plain (x-mean)*rstd here
    ln.x = X; ln.ldx = DD; ln.y = Y; ln.ldy = DD; ln.scale = nullptr; ln.bias = nullptr; ln.rows = M; ln.rows_per_batch = 1;
    ln.part = PART; ln.part_stride = part_stride; ln.nsplit = splits - 1;   // partial products of the previous layer's fc2
    HVLA_TRY((layernorm<float, bf16>(st, ln, DD)));
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::bqkv_f; ep.out = QKV; ep.ldo = 3 * DD;   // q / sqrt(64) is folded into the packed weights and bias
      HVLA_TRY(gemm(Y, m + Mx::wqkv, M, 3 * DD, DD, tc::EPI_BIAS_BF16, ep));
    }
    if (maps) HVLA_TRY((attn_probs<bf16, DHD>(st, QKV, maps + (int64_t)l * B * DH * DTOK * DTOK, DTOK, DH, B, 0, 1.0f)));   // q is pre-scaled
    if (simt_attn) {
      AttnP ap; memset(&ap, 0, sizeof ap);
      ap.qkv = QKV; ap.out = ATT; ap.S = DTOK; ap.H = DH; ap.nbatch = B; ap.mask = 0; ap.prescaled = 1;
      HVLA_TRY((attention_simt<bf16, bf16>(st, ap, DHD)));
    } else if (mma_attn) {
      HVLA_TRY(attn::dino_attention(st, QKV, ATT, B));
    } else {
      HVLA_TRY(attn_tc::dino_attention_tc(st, QKV, ATT, B));
    }
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::bo; ep.out = X; ep.ldo = DD; ep.ls = v + V::ls1;
      ep.part = PART; ep.part_bytes = PART_BYTES; ep.splits_used = &splits;
      HVLA_TRY(gemm(ATT, m + Mx::wo, M, DD, DD, tc::EPI_RESIDUAL_F32, ep));
    }
    ln.nsplit = splits - 1;      // this LayerNorm first folds the split-K partial products into the stream
    HVLA_TRY((layernorm<float, bf16>(st, ln, DD)));
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::b1_f; ep.out = HID; ep.ldo = DF;
      HVLA_TRY(gemm(Y, m + Mx::w1, M, DF, DD, tc::EPI_BIAS_GELU_BF16, ep));
    }
    {
      tc::EpiP ep; memset(&ep, 0, sizeof ep);
      ep.bias = v + V::b2; ep.out = X; ep.ldo = DD; ep.ls = v + V::ls2;
      ep.part = PART; ep.part_bytes = PART_BYTES; ep.splits_used = &splits;
      HVLA_TRY(gemm(HID, m + Mx::w2, M, DD, DF, tc::EPI_RESIDUAL_F32, ep));
    }
  }
  LnP ln; memset(&ln, 0, sizeof ln);
  ln.x = X; ln.ldx = DD; ln.y = out_emb; ln.ldy = DD; ln.scale = dv + V::lnf_s; ln.bias = dv + V::lnf_b; ln.rows = M; ln.rows_per_batch = 1;
  ln.part = PART; ln.part_stride = part_stride; ln.nsplit = splits - 1;
  HVLA_TRY((layernorm<float, bf16>(st, ln, DD)));
  return HVLA_OK;
}

// ---- base ViT + head, generic path (K8, K9) --------------------------------------------------------------
// TE: embedding storage type; TW: generated-weight storage type.  All math fp32.
template <typename TE, typename TW>
static int base_generic(cudaStream_t st, const TE* emb, const TW* weights, const int32_t* tidx, int B, int T,
                        float* out_action, float* out_logit, uint8_t* ws, const Plan& pl, float* maps = nullptr,
                        const BaseShape* shape = nullptr, int32_t* out_tokens = nullptr, float* out_top2 = nullptr) {
  typedef GenLayout G;
  const BaseShape sh = shape ? *shape : mix_shape();     // mix head: one readout token; discrete head: 4 or 28 (discrete_head.cuh)
  const int S = sh.S;
  float* PT = reinterpret_cast<float*>(ws + pl.pt);
  float* X = reinterpret_cast<float*>(ws + pl.xb);
  float* Y = reinterpret_cast<float*>(ws + pl.yb);
  float* QKV = reinterpret_cast<float*>(ws + pl.qkvb);
  float* CB = reinterpret_cast<float*>(ws + pl.cb);
  float* HB = reinterpret_cast<float*>(ws + pl.hb);
  const int64_t sW = sh.ngp;               // weight-batch stride; T==1 is handled by an all-zero index below
  // image_embedding_projection on patch tokens (skip CLS row): base_vit.py:122, 130-133
  {
    GemmP g = gemm_params(emb + DD, DD, weights + G::proj_w, BD, weights + G::proj_b, PT, BD, NPATCH, BD, DD);
    g.sA = (int64_t)DTOK * DD; g.sW = sW; g.sBias = sW; g.sC = (int64_t)NPATCH * BD; g.widx = tidx;
    HVLA_TRY((gemm_simt<TE, TW, TW, float>(st, g, B)));
  }
  {
    const int64_t total = (int64_t)B * S * BD;
    base_assemble_kernel<TW><<<cdiv(total, 256), 256, 0, st>>>(PT, weights, tidx, X, B, S, sh.ngp);
    HVLA_LAUNCH_CHECK("base_assemble");
  }
  for (int l = 0; l < BL; ++l) {
    const TW* lw = weights + sh.layers + (int64_t)l * G::layer_size;
    LnP ln; memset(&ln, 0, sizeof ln);
    ln.x = X; ln.ldx = BD; ln.y = Y; ln.ldy = BD; ln.scale = lw + G::ln0_s; ln.bias = lw + G::ln0_b; ln.sS = sW; ln.widx = tidx;
    ln.rows = B * S; ln.rows_per_batch = S;
    HVLA_TRY((layernorm<TW, float>(st, ln, BD)));
    const int64_t wofs[3] = {G::wq, G::wk, G::wv}, bofs[3] = {G::bq, G::bk, G::bv};
    for (int j = 0; j < 3; ++j) {
      GemmP g = gemm_params(Y, BD, lw + wofs[j], BD, lw + bofs[j], QKV + j * BD, 3 * BD, S, BD, BD);
      g.sA = (int64_t)S * BD; g.sW = sW; g.sBias = sW; g.sC = (int64_t)S * 3 * BD; g.widx = tidx;
      HVLA_TRY((gemm_simt<float, TW, TW, float>(st, g, B)));
    }
    if (maps) HVLA_TRY((attn_probs<float, BHD>(st, QKV, maps + (int64_t)l * B * BH * BTOK * BTOK, BTOK, BH, B, 1, 0.25f)));
    AttnP ap; memset(&ap, 0, sizeof ap);
    ap.qkv = QKV; ap.out = CB; ap.S = S; ap.H = BH; ap.nbatch = B; ap.mask = 1; ap.n_act = sh.A;
    HVLA_TRY((attention_simt<float, float>(st, ap, BHD)));
    {
      GemmP g = gemm_params(CB, BD, lw + G::wo, BD, lw + G::bo, X, BD, S, BD, BD);
      g.sA = (int64_t)S * BD; g.sW = sW; g.sBias = sW; g.sC = (int64_t)S * BD; g.widx = tidx;
      g.R = X; g.ldr = BD; g.sR = (int64_t)S * BD;
      HVLA_TRY((gemm_simt<float, TW, TW, float>(st, g, B)));
    }
    ln.scale = lw + G::ln1_s; ln.bias = lw + G::ln1_b;
    HVLA_TRY((layernorm<TW, float>(st, ln, BD)));
    {
      GemmP g = gemm_params(Y, BD, lw + G::w0, BF, lw + G::b0, HB, BF, S, BF, BD);
      g.sA = (int64_t)S * BD; g.sW = sW; g.sBias = sW; g.sC = (int64_t)S * BF; g.widx = tidx; g.act = 1;
      HVLA_TRY((gemm_simt<float, TW, TW, float>(st, g, B)));
    }
    {
      GemmP g = gemm_params(HB, BF, lw + G::w1, BD, lw + G::b1, X, BD, S, BD, BF);
      g.sA = (int64_t)S * BF; g.sW = sW; g.sBias = sW; g.sC = (int64_t)S * BD; g.widx = tidx;
      g.R = X; g.ldr = BD; g.sR = (int64_t)S * BD;
      HVLA_TRY((gemm_simt<float, TW, TW, float>(st, g, B)));
    }
  }
  if (sh.V > 0) {                          // DiscreteActionHead + BinTokenizer.decode
    ProfScope ps(st, "discrete_head");
    discrete_head_kernel<TW><<<B, 256, 0, st>>>(X, weights, tidx, sh, out_tokens, out_action, out_top2);
    HVLA_LAUNCH_CHECK("discrete_head");
    return HVLA_OK;
  }
  ProfScope ps(st, "mix_head");
  mix_head_kernel<TW><<<cdiv(B, 4), 128, 0, st>>>(X, weights, tidx, out_action, out_logit, B);
  HVLA_LAUNCH_CHECK("mix_head");
  return HVLA_OK;
}

// task index handling: NULL means identity (T==B) or all-zero (T==1)
__global__ void fill_index_kernel(int* idx, int B, int identity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) idx[i] = identity ? i : 0;
}

static int check_common(int B, int T, int dtype, const void* ws, size_t ws_bytes) {
  if (dtype != HVLA_F32 && dtype != HVLA_BF16 && dtype != HVLA_BF16X3) return fail(HVLA_ERR_ARG, "dtype must be HVLA_F32, HVLA_BF16 or HVLA_BF16X3");
  if (B < 0 || T < 0) return fail(HVLA_ERR_ARG, "negative batch");
  if (!ws) return fail(HVLA_ERR_WORKSPACE, "workspace is NULL");
  if (ws_bytes < make_plan(B, T, dtype).total) return fail(HVLA_ERR_WORKSPACE, "workspace too small (see hvla_workspace_bytes)");
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(HVLA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  return HVLA_OK;
}

static int base_act_impl(cudaStream_t st, const void* emb, const void* weights, const int32_t* task_index, int B, int T,
                         float* out_action, float* out_logit, uint8_t* ws, const Plan& pl, int dtype, float* maps = nullptr,
                         const BaseShape* shape = nullptr, int32_t* out_tokens = nullptr, float* out_top2 = nullptr) {
  if (!task_index && !(T == B || T == 1)) return fail(HVLA_ERR_ARG, "task_index is NULL but T != B and T != 1");
  const int32_t* tidx = task_index;
  if (!tidx && T == 1 && B > 1) {   // shared weights: materialise an all-zero index (B == 1: identity already is)
    int* z = reinterpret_cast<int*>(ws + pl.tidx);
    fill_index_kernel<<<cdiv(B, 256), 256, 0, st>>>(z, B, 0);
    HVLA_LAUNCH_CHECK("fill_index");
    tidx = z;
  }
  if (dtype == HVLA_F32 || dtype == HVLA_BF16X3)        // fp32 embeddings and generated weights: the exact CUDA-core base net
    return base_generic<float, float>(st, reinterpret_cast<const float*>(emb), reinterpret_cast<const float*>(weights), tidx, B,
                                      T, out_action, out_logit, ws, pl, maps, shape, out_tokens, out_top2);
  // attention maps / discrete head (several readout tokens) / debugging aid: the same math through the generic CUDA-core kernels
  if (maps || shape || env_flag("HVLA_DEBUG_GENERIC_BASE"))
    return base_generic<bf16, bf16>(st, reinterpret_cast<const bf16*>(emb), reinterpret_cast<const bf16*>(weights), tidx, B, T,
                                    out_action, out_logit, ws, pl, maps, shape, out_tokens, out_top2);
  return basefused::base_act_bf16(st, reinterpret_cast<const bf16*>(emb), reinterpret_cast<const bf16*>(weights), tidx, B,
                                  out_action, out_logit);
}

}  // namespace hvla

// =====================================================================================================
// extern "C"
// =====================================================================================================
using namespace hvla;

extern "C" {
#ifdef HVLA_CHAIN_STATS
int hvla_debug_chain_stats(unsigned long long* out) { return cudaMemcpyFromSymbol(out, chain::g_chain_stats, 64) == cudaSuccess ? 0 : 1; }
#endif

int hvla_version(void) { return 1; }
const char* hvla_last_error(void) { return g_last_error.c_str(); }
int64_t hvla_hn_blob_elems(void) { return HnLayout::total; }
int64_t hvla_generated_elems(void) { return NG; }
int64_t hvla_generated_row_stride(void) { return NGP; }
int64_t hvla_dino_vec_elems(void) { return DvecLayout::total; }
int64_t hvla_dino_mat_elems(void) { return DmatLayout::total; }
int64_t hvla_launch_count(void) { return g_launches.load(); }

int64_t hvla_layout_offset(const char* name) {
  if (!name) return -1;
  std::string s(name);
  int layer = 0;
  auto split_layer = [&](std::string& rest) -> bool {   // "l<k>.<field>" -> layer, field
    if (rest.size() > 2 && rest[0] == 'l' && isdigit((unsigned char)rest[1])) {
      size_t dot = rest.find('.');
      if (dot == std::string::npos) return false;
      layer = atoi(rest.substr(1, dot - 1).c_str());
      rest = rest.substr(dot + 1);
      return true;
    }
    return false;
  };
#define F(tbl, nm) if (f == #nm) return tbl::nm;
#define FL(tbl, nm) if (f == #nm) return tbl::layers + (int64_t)layer * tbl::layer_size + tbl::nm;
  if (s.rfind("hn.", 0) == 0) {
    std::string f = s.substr(3);
    if (split_layer(f)) {
      if (layer < 0 || layer >= CL) return -1;
      FL(HnLayout, ln0_s) FL(HnLayout, ln0_b) FL(HnLayout, wqkv) FL(HnLayout, bqkv) FL(HnLayout, wo) FL(HnLayout, bo)
      FL(HnLayout, ln1_s) FL(HnLayout, ln1_b) FL(HnLayout, w0) FL(HnLayout, b0) FL(HnLayout, w1) FL(HnLayout, b1)
      return -1;
    }
    F(HnLayout, tok_w) F(HnLayout, tok_b) F(HnLayout, img_w) F(HnLayout, img_b) F(HnLayout, task_pos) F(HnLayout, img_pos)
    F(HnLayout, layer_pos) F(HnLayout, encn_s) F(HnLayout, encn_b) F(HnLayout, total)
    return -1;
  }
  if (s.rfind("gen.", 0) == 0) {
    std::string f = s.substr(4);
    if (split_layer(f)) {
      if (layer < 0 || layer >= BL) return -1;
      FL(GenLayout, ln0_s) FL(GenLayout, ln0_b) FL(GenLayout, wq) FL(GenLayout, bq) FL(GenLayout, wk) FL(GenLayout, bk)
      FL(GenLayout, wv) FL(GenLayout, bv) FL(GenLayout, wo) FL(GenLayout, bo) FL(GenLayout, ln1_s) FL(GenLayout, ln1_b)
      FL(GenLayout, w0) FL(GenLayout, b0) FL(GenLayout, w1) FL(GenLayout, b1)
      return -1;
    }
    F(GenLayout, proj_w) F(GenLayout, proj_b) F(GenLayout, pos) F(GenLayout, encn_s) F(GenLayout, encn_b) F(GenLayout, wc)
    F(GenLayout, bc) F(GenLayout, wd) F(GenLayout, bd) F(GenLayout, total)
    return -1;
  }
  if (s.rfind("dvec.", 0) == 0) {
    std::string f = s.substr(5);
    if (split_layer(f)) {
      if (layer < 0 || layer >= DL) return -1;
      FL(DvecLayout, ln1_s) FL(DvecLayout, ln1_b) FL(DvecLayout, bqkv) FL(DvecLayout, bo) FL(DvecLayout, ls1)
      FL(DvecLayout, ln2_s) FL(DvecLayout, ln2_b) FL(DvecLayout, b1) FL(DvecLayout, b2) FL(DvecLayout, ls2)
      FL(DvecLayout, bqkv_f) FL(DvecLayout, cs_qkv) FL(DvecLayout, b1_f) FL(DvecLayout, cs_1)
      return -1;
    }
    F(DvecLayout, patch_b) F(DvecLayout, cls) F(DvecLayout, pos) F(DvecLayout, lnf_s) F(DvecLayout, lnf_b) F(DvecLayout, pos_blk) F(DvecLayout, total)
    return -1;
  }
  if (s.rfind("dmat.", 0) == 0) {
    std::string f = s.substr(5);
    if (split_layer(f)) {
      if (layer < 0 || layer >= DL) return -1;
      FL(DmatLayout, wqkv) FL(DmatLayout, wo) FL(DmatLayout, w1) FL(DmatLayout, w2)
      return -1;
    }
    F(DmatLayout, patch_w) F(DmatLayout, total)
    return -1;
  }
#undef F
#undef FL
  return -1;
}

size_t hvla_workspace_bytes(int B, int T, int dtype) { return make_plan(B, T, dtype).total; }

int hvla_generate(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w, const float* heads_b,
                  const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad, const float* init_cls, int T,
                  void* out_weights, float* out_ctx, void* workspace, size_t workspace_bytes, int dtype) {
  if (!hn_blob || !heads_w || !heads_b || !tok_emb || !tok_mask || !init_cls || !out_weights)
    return fail(HVLA_ERR_ARG, "hvla_generate: NULL argument");
  HVLA_TRY(check_common(0, T, dtype, workspace, workspace_bytes));
  if (T == 0) return HVLA_OK;
  const Plan pl = make_plan(0, T, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (dtype != HVLA_BF16)      // HVLA_F32 and HVLA_BF16X3: fp32 generated weights from the fp32 path
    return generate_impl<float>(st, hn_blob, nullptr, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, out_weights, out_ctx, ws, pl);
  return generate_impl<bf16>(st, hn_blob, hn_blob_f16, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, out_weights, out_ctx, ws, pl);
}

int hvla_generate_rows(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w, const float* heads_b,
                       const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad, const float* init_cls, int T,
                       const int32_t* row_index, int T_max, void* weights, float* ctx, void* workspace, size_t workspace_bytes,
                       int dtype) {
  if (!hn_blob || !heads_w || !heads_b || !tok_emb || !tok_mask || !init_cls || !weights || !row_index)
    return fail(HVLA_ERR_ARG, "hvla_generate_rows: NULL argument");
  if (T_max <= 0 || T > T_max) return fail(HVLA_ERR_ARG, "hvla_generate_rows: need 0 <= T <= T_max");
  HVLA_TRY(check_common(0, T, dtype, workspace, workspace_bytes));
  if (T == 0) return HVLA_OK;
  const Plan pl = make_plan(0, T, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (dtype != HVLA_BF16)
    return generate_impl<float>(st, hn_blob, nullptr, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, weights, ctx, ws, pl,
                                row_index, T_max);
  return generate_impl<bf16>(st, hn_blob, hn_blob_f16, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, weights, ctx, ws, pl,
                             row_index, T_max);
}

int64_t hvla_discrete_generated_elems(int n_action_tokens) {
  BaseShape sh;
  return discrete_shape(n_action_tokens, &sh) ? sh.total : -1;
}
int64_t hvla_discrete_row_stride(int n_action_tokens) {
  BaseShape sh;
  return discrete_shape(n_action_tokens, &sh) ? sh.ngp : -1;
}

int hvla_generate_n(hvla_stream_t stream, const float* hn_blob, const void* hn_blob_f16, const void* heads_w, const float* heads_b,
                    const float* tok_emb, const int32_t* tok_mask, const uint8_t* lang_pad, const float* init_cls, int T,
                    int64_t row_stride, void* out_weights, float* out_ctx, void* workspace, size_t workspace_bytes, int dtype) {
  if (!hn_blob || !heads_w || !heads_b || !tok_emb || !tok_mask || !init_cls || !out_weights)
    return fail(HVLA_ERR_ARG, "hvla_generate_n: NULL argument");
  if (row_stride <= 0 || row_stride % 32 != 0) return fail(HVLA_ERR_ARG, "hvla_generate_n: row_stride must be a positive multiple of 32");
  HVLA_TRY(check_common(0, T, dtype, workspace, workspace_bytes));
  if (T == 0) return HVLA_OK;
  const Plan pl = make_plan(0, T, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (dtype != HVLA_BF16)
    return generate_impl<float>(st, hn_blob, nullptr, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, out_weights, out_ctx, ws, pl,
                                nullptr, 0, row_stride);
  return generate_impl<bf16>(st, hn_blob, hn_blob_f16, heads_w, heads_b, tok_emb, tok_mask, lang_pad, init_cls, T, out_weights, out_ctx, ws, pl,
                             nullptr, 0, row_stride);
}

int hvla_act_discrete(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images, const void* weights,
                      const int32_t* task_index, int B, int T, int n_action_tokens, float* out_action, int32_t* out_tokens, float* out_top2,
                      void* workspace, size_t workspace_bytes, int dtype) {
  if (!dino_vec || !dino_mat || !images || !weights || !out_action || !out_tokens) return fail(HVLA_ERR_ARG, "hvla_act_discrete: NULL argument");
  BaseShape sh;
  if (!discrete_shape(n_action_tokens, &sh))
    return fail(HVLA_ERR_UNSUPPORTED, "hvla_act_discrete: n_action_tokens must be 4 (action_horizon) or 28 (action_dim_and_action_horizon)");
  if (T <= 0 && B > 0) return fail(HVLA_ERR_ARG, "hvla_act_discrete: T must be >= 1");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* emb = ws + pl.emb;
  HVLA_TRY(hvla_dino_forward(stream, dino_vec, dino_mat, images, B, emb, workspace, workspace_bytes, dtype));
  return base_act_impl(reinterpret_cast<cudaStream_t>(stream), emb, weights, task_index, B, T, out_action, nullptr, ws, pl, dtype, nullptr, &sh,
                       out_tokens, out_top2);
}

int hvla_dino_forward(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images, int B,
                      void* out_emb, void* workspace, size_t workspace_bytes, int dtype) {
  if (!dino_vec || !dino_mat || !images || !out_emb) return fail(HVLA_ERR_ARG, "hvla_dino_forward: NULL argument");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (dtype == HVLA_F32)
    return dino_f32(st, dino_vec, reinterpret_cast<const float*>(dino_mat), images, B, reinterpret_cast<float*>(out_emb), ws, pl);
  if (dtype == HVLA_BF16X3) {
    x3::Ws w;
    w.X = reinterpret_cast<float*>(ws + pl.x); w.A3 = reinterpret_cast<bf16*>(ws + pl.a3); w.QKV2 = reinterpret_cast<bf16*>(ws + pl.qkv2);
    w.A3L = reinterpret_cast<bf16*>(ws + pl.a3l); w.A0 = reinterpret_cast<float*>(ws + pl.qkv2);      // im2col matrix aliases the q|k|v planes
    return x3::dino_forward(st, dino_vec, reinterpret_cast<const bf16*>(dino_mat), images, B, reinterpret_cast<float*>(out_emb), w);
  }
  return dino_bf16(st, dino_vec, reinterpret_cast<const bf16*>(dino_mat), images, B, reinterpret_cast<bf16*>(out_emb), ws, pl);
}

int hvla_base_act(hvla_stream_t stream, const void* emb, const void* weights, const int32_t* task_index, int B, int T,
                  float* out_action, float* out_logit, void* workspace, size_t workspace_bytes, int dtype) {
  if (!emb || !weights || !out_action) return fail(HVLA_ERR_ARG, "hvla_base_act: NULL argument");
  if (T <= 0 && B > 0) return fail(HVLA_ERR_ARG, "hvla_base_act: T must be >= 1");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  return base_act_impl(reinterpret_cast<cudaStream_t>(stream), emb, weights, task_index, B, T, out_action, out_logit,
                       reinterpret_cast<uint8_t*>(workspace), pl, dtype);
}

int hvla_act(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images, const void* weights,
             const int32_t* task_index, int B, int T, float* out_action, float* out_logit, void* workspace,
             size_t workspace_bytes, int dtype) {
  if (!dino_vec || !dino_mat || !images || !weights || !out_action) return fail(HVLA_ERR_ARG, "hvla_act: NULL argument");
  if (T <= 0 && B > 0) return fail(HVLA_ERR_ARG, "hvla_act: T must be >= 1");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* emb = ws + pl.emb;
  HVLA_TRY(hvla_dino_forward(stream, dino_vec, dino_mat, images, B, emb, workspace, workspace_bytes, dtype));
  return base_act_impl(reinterpret_cast<cudaStream_t>(stream), emb, weights, task_index, B, T, out_action, out_logit, ws, pl, dtype);
}

int hvla_act_debug(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images, const void* weights,
                   const int32_t* task_index, int B, int T, float* out_action, float* out_logit, float* dino_maps, float* base_maps,
                   void* workspace, size_t workspace_bytes, int dtype) {
  if (!dino_vec || !dino_mat || !images || !weights || !out_action) return fail(HVLA_ERR_ARG, "hvla_act_debug: NULL argument");
  if (T <= 0 && B > 0) return fail(HVLA_ERR_ARG, "hvla_act_debug: T must be >= 1");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* emb = ws + pl.emb;
  if (dtype == HVLA_BF16X3) return fail(HVLA_ERR_UNSUPPORTED, "hvla_act_debug: attention maps come from the HVLA_F32 or HVLA_BF16 paths");
  if (dtype == HVLA_F32)
    HVLA_TRY(dino_f32(st, dino_vec, reinterpret_cast<const float*>(dino_mat), images, B, reinterpret_cast<float*>(emb), ws, pl, dino_maps));
  else
    HVLA_TRY(dino_bf16(st, dino_vec, reinterpret_cast<const bf16*>(dino_mat), images, B, reinterpret_cast<bf16*>(emb), ws, pl, dino_maps));
  return base_act_impl(st, emb, weights, task_index, B, T, out_action, out_logit, ws, pl, dtype, base_maps);
}

int hvla_act_host(hvla_stream_t stream, const float* dino_vec, const void* dino_mat, const uint8_t* images_host,
                  const void* weights, const int32_t* task_index, int B, int T, float* out_action_host, float* out_logit_host,
                  void* workspace, size_t workspace_bytes, int dtype) {
  if (!images_host || !out_action_host) return fail(HVLA_ERR_ARG, "hvla_act_host: NULL host buffer");
  HVLA_TRY(check_common(B, 0, dtype, workspace, workspace_bytes));
  if (B == 0) return HVLA_OK;
  const Plan pl = make_plan(B, 0, dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* d_img = ws + pl.img;
  float* d_act = reinterpret_cast<float*>(ws + pl.act);
  float* d_logit = reinterpret_cast<float*>(ws + pl.logit);
  HVLA_CUDA(cudaMemcpyAsync(d_img, images_host, (size_t)B * IMG * IMG * 3, cudaMemcpyHostToDevice, st));
  HVLA_TRY(hvla_act(stream, dino_vec, dino_mat, d_img, weights, task_index, B, T, d_act, d_logit, workspace, workspace_bytes, dtype));
  HVLA_CUDA(cudaMemcpyAsync(out_action_host, d_act, (size_t)B * AH * AD * 4, cudaMemcpyDeviceToHost, st));
  if (out_logit_host)
    HVLA_CUDA(cudaMemcpyAsync(out_logit_host, d_logit, (size_t)B * AH * 4, cudaMemcpyDeviceToHost, st));
  HVLA_CUDA(cudaStreamSynchronize(st));
  return HVLA_OK;
}

int hvla_gemm_bf16(hvla_stream_t stream, const void* A, const void* Wt, const float* bias, void* C, int M, int N, int K, int act) {
  if (!A || !Wt || !bias || !C) return fail(HVLA_ERR_ARG, "hvla_gemm_bf16: NULL argument");
  tc::EpiP ep; memset(&ep, 0, sizeof ep);
  ep.bias = bias; ep.out = C; ep.ldo = N;
  const int epi = act == 2 ? tc::EPI_BIAS_GELU_BF16 : tc::EPI_BIAS_BF16;
  if (const char* e = getenv("HVLA_GEMM_DEBUG")) ep.debug = atoi(e);
  if (env_flag("HVLA_GEMM_1CTA")) return tc::gemm_tc(reinterpret_cast<cudaStream_t>(stream), A, Wt, M, N, K, epi, ep);
  return tc2::gemm_tc2(reinterpret_cast<cudaStream_t>(stream), A, Wt, M, N, K, epi, ep);
}

int hvla_profile_enable(int on) {
  for (auto& r : g_prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.recs.clear();
  g_prof.on = on != 0;
  return HVLA_OK;
}

int hvla_profile_report(char* buf, size_t cap) {
  if (!buf || cap == 0) return fail(HVLA_ERR_ARG, "hvla_profile_report: NULL buffer");
  HVLA_CUDA(cudaDeviceSynchronize());
  struct Agg { const char* name; int n; double ms; };
  std::vector<Agg> agg;
  for (auto& r : g_prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    bool found = false;
    for (auto& a : agg) if (!strcmp(a.name, r.name)) { a.n++; a.ms += ms; found = true; break; }
    if (!found) agg.push_back({r.name, 1, (double)ms});
  }
  size_t off = 0;
  buf[0] = 0;
  for (auto& a : agg) {
    int w = snprintf(buf + off, cap - off, "%s %d %.6f\n", a.name, a.n, a.ms);
    if (w < 0 || (size_t)w >= cap - off) break;
    off += (size_t)w;
  }
  return HVLA_OK;
}

int hvla_dino_attention(hvla_stream_t stream, const void* qkv, void* out, int B, int impl) {
  if (!qkv || !out || B <= 0) return fail(HVLA_ERR_ARG, "hvla_dino_attention: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == 0) return attn::dino_attention(st, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B);
  return attn_tc::dino_attention_tc(st, reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), B);
}

int64_t hvla_postprocess_state_floats(void) { return post::STATE_FLOATS; }

int hvla_postprocess(hvla_stream_t stream, const float* raw_action, float* state, const uint8_t* reset, int B, int norm_type,
                     const float* stat_a, const float* stat_b, const uint8_t* mask, int ensemble, float temp, int policy_setup,
                     int sticky_repeat, float* out_raw, float* out_action) {
  if (!raw_action || !state || !stat_a || !stat_b || !out_raw || !out_action || B < 0) return fail(HVLA_ERR_ARG, "hvla_postprocess: bad argument");
  if (norm_type < 0 || norm_type > 1 || policy_setup < 0 || policy_setup > 2) return fail(HVLA_ERR_UNSUPPORTED, "hvla_postprocess: unknown mode");
  if (B == 0) return HVLA_OK;
  post::PostP p;
  memset(&p, 0, sizeof p);
  p.raw = raw_action; p.state = state; p.reset = reset; p.out_raw = out_raw; p.out_action = out_action;
  for (int d = 0; d < AD; ++d) { p.a[d] = stat_a[d]; p.b[d] = stat_b[d]; p.mask[d] = mask ? mask[d] : 1; }
  p.B = B; p.norm_type = norm_type; p.ensemble = ensemble; p.temp = temp; p.policy = policy_setup; p.sticky_repeat = sticky_repeat;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope ps(st, "postprocess");
  post::postprocess_kernel<<<cdiv(B, 128), 128, 0, st>>>(p);
  HVLA_LAUNCH_CHECK("postprocess");
  return HVLA_OK;
}

size_t hvla_resize_workspace_bytes(int B, int H, int W, int S, int crop) {
  if (B <= 0 || H <= 0 || W <= 0 || S <= 0) return 0;
  return align_up((size_t)B * S * W * 3 * 4) + (crop ? align_up((size_t)B * S * S * 3 * 4) : 0);
}

int hvla_resize_lanczos3(hvla_stream_t stream, const uint8_t* images, int B, int H, int W, int S, const int32_t* starts_y,
                         const float* weights_y, int span_y, const int32_t* starts_x, const float* weights_x, int span_x, int crop,
                         const float* crop_params, uint8_t* out, void* workspace, size_t workspace_bytes) {
  if (B < 0 || H <= 0 || W <= 0 || S <= 1 || span_y <= 0 || span_x <= 0) return fail(HVLA_ERR_ARG, "hvla_resize_lanczos3: bad sizes");
  if (B == 0) return HVLA_OK;
  if (!images || !starts_y || !weights_y || !starts_x || !weights_x || !out || !workspace || (crop && !crop_params))
    return fail(HVLA_ERR_ARG, "hvla_resize_lanczos3: null argument");
  if (workspace_bytes < hvla_resize_workspace_bytes(B, H, W, S, crop)) return fail(HVLA_ERR_WORKSPACE, "hvla_resize_lanczos3: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* tmp = reinterpret_cast<float*>(workspace);
  float* full = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + align_up((size_t)B * S * W * 3 * 4));
  ProfScope ps(st, "resize");
  const int64_t n1 = (int64_t)B * S * W * 3, n2 = (int64_t)B * S * S * 3;
  const size_t row_bytes = (size_t)W * 3 * 4;
  if (row_bytes <= 48 * 1024) {
    // fused path: no float32 intermediate in HBM (preprocess.cuh)
    const bool vec = (W * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(images) & 3) == 0;
    const dim3 grid(S, B);
    if (!crop) {
      if (vec) prep::resize_fused_kernel<uint8_t, true><<<grid, 256, row_bytes, st>>>(images, out, starts_y, weights_y, span_y, starts_x, weights_x, span_x, H, W, S);
      else prep::resize_fused_kernel<uint8_t, false><<<grid, 256, row_bytes, st>>>(images, out, starts_y, weights_y, span_y, starts_x, weights_x, span_x, H, W, S);
    } else {
      if (vec) prep::resize_fused_kernel<float, true><<<grid, 256, row_bytes, st>>>(images, full, starts_y, weights_y, span_y, starts_x, weights_x, span_x, H, W, S);
      else prep::resize_fused_kernel<float, false><<<grid, 256, row_bytes, st>>>(images, full, starts_y, weights_y, span_y, starts_x, weights_x, span_x, H, W, S);
    }
    HVLA_LAUNCH_CHECK("resize_fused");
  } else {
    prep::resize_rows_kernel<<<cdiv(n1, 256), 256, 0, st>>>(images, tmp, starts_y, weights_y, span_y, H, W * 3, S, n1);
    HVLA_LAUNCH_CHECK("resize_rows");
    if (!crop) prep::resize_cols_kernel<uint8_t><<<cdiv(n2, 256), 256, 0, st>>>(tmp, out, starts_x, weights_x, span_x, W, S, n2);
    else prep::resize_cols_kernel<float><<<cdiv(n2, 256), 256, 0, st>>>(tmp, full, starts_x, weights_x, span_x, W, S, n2);
    HVLA_LAUNCH_CHECK("resize_cols");
  }
  if (crop) {
    prep::crop_bilinear_kernel<<<cdiv(n2, 256), 256, 0, st>>>(full, out, S, crop_params[0], crop_params[1], crop_params[2], crop_params[3], n2);
    HVLA_LAUNCH_CHECK("crop_bilinear");
  }
  return HVLA_OK;
}

int64_t hvla_t5_blob_elems(void) { return t5::Layout::total; }
size_t hvla_t5_workspace_bytes(int T, int S) { return (T > 0 && S > 0) ? t5::workspace_bytes(T, S) : 0; }

int hvla_t5_encode(hvla_stream_t stream, const float* t5_blob, const float* pos_bias, const int32_t* input_ids, const int32_t* attention_mask,
                   int T, int S, float* out_emb, void* workspace, size_t workspace_bytes) {
  if (T < 0 || T > 65535 || S <= 0 || S > t5::SMAX) return fail(HVLA_ERR_ARG, "hvla_t5_encode: need 0 <= T <= 65535 and 1 <= S <= 32");
  if (T == 0) return HVLA_OK;
  if (!t5_blob || !pos_bias || !input_ids || !attention_mask || !out_emb || !workspace) return fail(HVLA_ERR_ARG, "hvla_t5_encode: null argument");
  if (workspace_bytes < t5::workspace_bytes(T, S)) return fail(HVLA_ERR_WORKSPACE, "hvla_t5_encode: workspace too small");
  return t5::encode(reinterpret_cast<cudaStream_t>(stream), t5_blob, pos_bias, input_ids, attention_mask, T, S, out_emb,
                    reinterpret_cast<uint8_t*>(workspace));
}

int64_t hvla_t5_mat_elems(void) { return t5::MatLayout::total; }
size_t hvla_t5_tc_workspace_bytes(int T, int S) { return (T > 0 && S > 0) ? t5::workspace_bytes_tc(T, S) : 0; }

int hvla_t5_encode_tc(hvla_stream_t stream, const float* t5_blob, const void* t5_mat, const float* pos_bias, const int32_t* input_ids,
                      const int32_t* attention_mask, int T, int S, float* out_emb, void* workspace, size_t workspace_bytes) {
  if (T < 0 || T > 65535 || S <= 0 || S > t5::SMAX) return fail(HVLA_ERR_ARG, "hvla_t5_encode_tc: need 0 <= T <= 65535 and 1 <= S <= 32");
  if (T == 0) return HVLA_OK;
  if (!t5_blob || !t5_mat || !pos_bias || !input_ids || !attention_mask || !out_emb || !workspace)
    return fail(HVLA_ERR_ARG, "hvla_t5_encode_tc: null argument");
  if (workspace_bytes < t5::workspace_bytes_tc(T, S)) return fail(HVLA_ERR_WORKSPACE, "hvla_t5_encode_tc: workspace too small");
  return t5::encode_tc(reinterpret_cast<cudaStream_t>(stream), t5_blob, reinterpret_cast<const bf16*>(t5_mat), pos_bias, input_ids,
                       attention_mask, T, S, out_emb, reinterpret_cast<uint8_t*>(workspace));
}

// ---- legacy XLA custom-call wrappers -----------------------------------------------------------------------
// Status-returning form (API_VERSION_STATUS_RETURNING): void f(stream, buffers, opaque, opaque_len, XlaCustomCallStatus*).
// The failure hook is XLA's own XlaCustomCallStatusSetFailure, resolved as a WEAK symbol (present when the library is
// loaded into a process that links xla / jaxlib) or registered explicitly with hvla_xla_register_status_setter.
// On any error the outputs are zero-filled (when the opaque is long enough to know their sizes) so that a caller that
// ignores the status never reads uninitialised memory.
extern "C" void XlaCustomCallStatusSetFailure(void* status, const char* message, size_t message_len) __attribute__((weak));
static std::atomic<hvla_xla_status_setter> g_status_setter{nullptr};

static void xla_fail(void* status, const char* who) {
  std::string msg = std::string(who) + ": " + g_last_error;
  fprintf(stderr, "%s\n", msg.c_str());
  if (!status) return;
  hvla_xla_status_setter fn = g_status_setter.load();
  if (fn) fn(status, msg.c_str(), msg.size());
  else if (XlaCustomCallStatusSetFailure) XlaCustomCallStatusSetFailure(status, msg.c_str(), msg.size());
}

void hvla_xla_register_status_setter(hvla_xla_status_setter fn) { g_status_setter.store(fn); }

void hvla_xla_generate_status(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < sizeof(hvla_xla_opaque)) {
    fail(HVLA_ERR_ARG, "opaque is shorter than struct hvla_xla_opaque");
    xla_fail(status, "hvla_xla_generate");
    return;
  }
  hvla_xla_opaque o; memcpy(&o, opaque, sizeof o);
  int r = hvla_generate(stream, (const float*)b[0], b[1], b[2], (const float*)b[3], (const float*)b[4], (const int32_t*)b[5],
                        (const uint8_t*)b[6], (const float*)b[7], o.T, b[8], (float*)b[9], b[10], (size_t)o.workspace_bytes, o.dtype);
  if (r != HVLA_OK) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t es = o.dtype == HVLA_BF16 ? 2 : 4;
    if (o.T > 0 && b[8]) cudaMemsetAsync(b[8], 0, (size_t)o.T * NGP * es, st);
    if (o.T > 0 && b[9]) cudaMemsetAsync(b[9], 0, (size_t)o.T * CD * 4, st);
    xla_fail(status, "hvla_xla_generate");
  }
}
void hvla_xla_act_status(void* stream, void** b, const char* opaque, size_t opaque_len, void* status) {
  if (!b || !opaque || opaque_len < sizeof(hvla_xla_opaque)) {
    fail(HVLA_ERR_ARG, "opaque is shorter than struct hvla_xla_opaque");
    xla_fail(status, "hvla_xla_act");
    return;
  }
  hvla_xla_opaque o; memcpy(&o, opaque, sizeof o);
  int r = hvla_act(stream, (const float*)b[0], b[1], (const uint8_t*)b[2], b[3], (const int32_t*)b[4], o.B, o.T, (float*)b[5],
                   (float*)b[6], b[7], (size_t)o.workspace_bytes, o.dtype);
  if (r != HVLA_OK) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (o.B > 0 && b[5]) cudaMemsetAsync(b[5], 0, (size_t)o.B * AH * AD * 4, st);
    if (o.B > 0 && b[6]) cudaMemsetAsync(b[6], 0, (size_t)o.B * AH * 4, st);
    xla_fail(status, "hvla_xla_act");
  }
}
// original (API_VERSION_ORIGINAL) form: no status channel -- same behaviour (message on stderr, outputs zero-filled)
void hvla_xla_generate(void* stream, void** b, const char* opaque, size_t opaque_len) { hvla_xla_generate_status(stream, b, opaque, opaque_len, nullptr); }
void hvla_xla_act(void* stream, void** b, const char* opaque, size_t opaque_len) { hvla_xla_act_status(stream, b, opaque, opaque_len, nullptr); }

}  // extern "C"
