// Shared constants, blob layouts and small helpers for libhvla (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/hvla.h"

namespace hvla {

// ---- model constants (README configuration of the reference) ----------
constexpr int IMG = 224, PATCH = 14, GRID = 16, NPATCH = 256;
constexpr int DTOK = 257, DD = 768, DL = 12, DH = 12, DHD = 64, DF = 3072;
constexpr int PATCH_K = 588, PATCH_KP = 640;
constexpr int CD = 128, CL = 6, CH = 4, CHD = 32, CF = 512, LANG = 32, LANGD = 768, CTOK = 34;
constexpr int BD = 64, BL = 4, BH = 4, BHD = 16, BF = 128, BTOK = 257;
constexpr int AH = 4, AD = 7, NCONT = 24;
constexpr int64_t NG = 201500, NGP = 201504;

// ---- HN blob layout (fp32 elements) ----------------------------------------------------
struct HnLayout {
  static constexpr int64_t tok_w = 0;
  static constexpr int64_t tok_b = tok_w + (int64_t)LANGD * CD;
  static constexpr int64_t img_w = tok_b + CD;
  static constexpr int64_t img_b = img_w + (int64_t)DD * CD;
  static constexpr int64_t task_pos = img_b + CD;
  static constexpr int64_t img_pos = task_pos + LANG * CD;
  static constexpr int64_t layer_pos = img_pos + CD;
  static constexpr int64_t layers = layer_pos + CD;
  // per layer
  static constexpr int64_t ln0_s = 0, ln0_b = ln0_s + CD, wqkv = ln0_b + CD, bqkv = wqkv + CD * 3 * CD,
                           wo = bqkv + 3 * CD, bo = wo + CD * CD, ln1_s = bo + CD, ln1_b = ln1_s + CD,
                           w0 = ln1_b + CD, b0 = w0 + CD * CF, w1 = b0 + CF, b1 = w1 + CF * CD,
                           layer_size = b1 + CD;
  static constexpr int64_t encn_s = layers + CL * layer_size;
  static constexpr int64_t encn_b = encn_s + CD;
  static constexpr int64_t total = encn_b + CD;
};

// ---- generated (per-task) row layout ----------------------------------------------------
struct GenLayout {
  static constexpr int64_t proj_w = 0;
  static constexpr int64_t proj_b = proj_w + (int64_t)DD * BD;
  static constexpr int64_t pos = proj_b + BD;
  static constexpr int64_t layers = pos + BTOK * BD;
  static constexpr int64_t ln0_s = 0, ln0_b = 64, wq = 128, bq = wq + 4096, wk = bq + 64, bk = wk + 4096,
                           wv = bk + 64, bv = wv + 4096, wo = bv + 64, bo = wo + 4096, ln1_s = bo + 64,
                           ln1_b = ln1_s + 64, w0 = ln1_b + 64, b0 = w0 + 8192, w1 = b0 + 128, b1 = w1 + 8192,
                           layer_size = b1 + 64;
  static constexpr int64_t encn_s = layers + BL * layer_size;
  static constexpr int64_t encn_b = encn_s + 64;
  static constexpr int64_t wc = encn_b + 64;
  static constexpr int64_t bc = wc + 64 * NCONT;
  static constexpr int64_t wd = bc + NCONT;
  static constexpr int64_t bd = wd + 64 * AH;
  static constexpr int64_t total = bd + AH;
};
static_assert(GenLayout::layer_size == 33472, "base block size");
static_assert(GenLayout::total == NG, "generated row size");

// ---- DINO vector blob (fp32) -----------------------------------------------------------------
struct DvecLayout {
  static constexpr int64_t patch_b = 0, cls = DD, pos = 2 * DD, layers = pos + (int64_t)DTOK * DD;
  static constexpr int64_t ln1_s = 0, ln1_b = DD, bqkv = 2 * DD, bo = bqkv + 3 * DD, ls1 = bo + DD, ln2_s = ls1 + DD,
                           ln2_b = ln2_s + DD, b1 = ln2_b + DD, b2 = b1 + DF, ls2 = b2 + DD,
                           // LayerNorm folded into the next linear (tensor-core path): folded biases and column sums of gamma*W
                           bqkv_f = ls2 + DD, cs_qkv = bqkv_f + 3 * DD, b1_f = cs_qkv + 3 * DD, cs_1 = b1_f + DF,
                           layer_size = cs_1 + DF;
  static constexpr int64_t lnf_s = layers + DL * layer_size, lnf_b = lnf_s + DD;
  // position rows of the 256 patch tokens in the blocked stream layout, indexed by patch p: [p >> 5][col >> 2][p & 31][4]
  static constexpr int64_t pos_blk = lnf_b + DD, total = pos_blk + (int64_t)NPATCH * DD;
};

// ---- DINO matrix blob (fp32 [K,N] or bf16 [N,K]; same element offsets) ------------------------
struct DmatLayout {
  static constexpr int64_t patch_w = 0, layers = (int64_t)PATCH_KP * DD;
  static constexpr int64_t wqkv = 0, wo = wqkv + (int64_t)DD * 3 * DD, w1 = wo + (int64_t)DD * DD,
                           w2 = w1 + (int64_t)DD * DF, layer_size = w2 + (int64_t)DF * DD;
  static constexpr int64_t total = layers + DL * layer_size;
};

// ---- error plumbing -------------------------------------------------------------------------
extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a, b);
  g_last_error = buf;
  return code;
}

#define HVLA_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) return ::hvla::fail(HVLA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define HVLA_LAUNCH_CHECK(name)                                                             \
  do {                                                                                      \
    ::hvla::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    cudaError_t e__ = cudaPeekAtLastError();                                                \
    if (e__ != cudaSuccess) return ::hvla::fail(HVLA_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// ---- optional per-kernel-class event timing (bench.py's live roofline measurement) --------------
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct ProfState { bool on = false; std::vector<ProfRec> recs; };
extern ProfState g_prof;
struct ProfScope {
  cudaStream_t st; cudaEvent_t b = nullptr;
  ProfScope(cudaStream_t s, const char* name) : st(s) {
    if (!g_prof.on) return;
    ProfRec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    b = r.b;
    g_prof.recs.push_back(r);
  }
  ~ProfScope() { if (b) cudaEventRecord(b, st); }
};

#define HVLA_TRY(expr)        \
  do {                        \
    int r__ = (expr);         \
    if (r__ != HVLA_OK) return r__; \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Every kernel of the act path starts with pdl_trigger() (its successor in the stream may be scheduled as soon as
// all CTAs of this grid have started), does its local prologue (barrier init, TMEM alloc, descriptor prefetch) and
// only then pdl_wait()s for the full completion + memory flush of its predecessor before touching global memory.
// This hides launch latency and prologue behind the predecessor's tail (~90 kernel boundaries per control step).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
extern bool g_pdl;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- dtype helpers -------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(bf16& d, float v) { d = __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // flax.linen.gelu(approximate=True): 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))  (transformer.py:66)
  const float c = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * (x * x * x))));
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  // exact GELU used by DINOv2 (HF ACT2FN["gelu"])
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// erf-GELU for bf16 epilogues: erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7) with MUFU rcp/ex2;
// gelu(x) = x * Phi(x), Phi = 1 - q (x >= 0) or q (x < 0), q = 0.5 * poly(t) * exp(-x^2/2).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  p *= t;
  const float e = exp2f(x * x * -0.72134752044448170f);   // exp(-x^2/2)
  const float h = x * (p * e);
  return x >= 0.f ? x - h : h;
}

// erf-GELU for the fc1 epilogue at 1 MUFU + 7 FMA-pipe ops: gelu(x) = 0.5 x (1 + tanh(x (a + b x^2 + c x^4))) with
// (a, b, c) fitted so that the result deviates from the EXACT erf-GELU by < 2.6e-5 absolute over all x
// (the textbook tanh-GELU constants deviate by 4.7e-4); tanh.approx.f32 adds < 2^-11 relative.  Both are far
// below the bf16 rounding of the stored activation (2^-9 relative).  x^2 is clamped where tanh is saturated.
__device__ __forceinline__ float gelu_erf_tanhfit(float x) {
  const float t = fminf(x * x, 80.0f);
  float w = fmaf(-0.00035151678755022096f, t, 0.0370056460170224f);
  w = fmaf(w, t, 0.7975078842849359f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x * w));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// two elements at a time on the packed fp32 pipe (FMUL2 / FFMA2): 6 packed + 2 FMNMX + 2 MUFU per pair instead of 14 + 2
__device__ __forceinline__ float2 gelu_erf_tanhfit2(float2 x) {
  float2 t = __fmul2_rn(x, x);
  t.x = fminf(t.x, 80.0f);
  t.y = fminf(t.y, 80.0f);
  float2 w = __ffma2_rn(make_float2(-0.00035151678755022096f, -0.00035151678755022096f), t,
                        make_float2(0.0370056460170224f, 0.0370056460170224f));
  w = __ffma2_rn(w, t, make_float2(0.7975078842849359f, 0.7975078842849359f));
  const float2 a = __fmul2_rn(x, w);
  float2 th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(a.y));
  const float2 hx = __fmul2_rn(make_float2(0.5f, 0.5f), x);
  return __ffma2_rn(hx, th, hx);
}

// One-time setup per DEVICE (cudaFuncSetAttribute applies to the current device only): bit d of `mask` = done on device d.
inline bool device_once(std::atomic<uint64_t>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  return !(mask.fetch_or(bit) & bit);
}
// SM count of the current device (cached per device ordinal)
inline int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace hvla
