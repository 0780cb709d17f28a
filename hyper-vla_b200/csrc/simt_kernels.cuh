// CUDA-core (fp32 math) kernels: the exact HVLA_F32 path and the glue of the bf16 path.
// Storage types are templated (float or bf16); all arithmetic is fp32 in the reference's
// op order, so this path matches the fp32 oracle to ~1e-6.
#pragma once
#include "common.cuh"
#include <float.h>
#include <type_traits>

namespace hvla {

// =============================================================================================
// Batched GEMM  C[b] = epi( A[b][M,K] * W[w(b)][K,N] + bias[w(b)][N] )
//   epi: out_scale * (.) -> activation -> optional  R + ls * (.)
// 64x64x16 tiles, 256 threads, 4x4 micro-tile.  K % 16 == 0.
// =============================================================================================
struct GemmP {
  const void* A; int lda; int64_t sA;
  const void* W; int ldw; int64_t sW;
  const void* bias; int64_t sBias;
  void* C; int ldc; int64_t sC;
  const float* R; int ldr; int64_t sR;
  const float* ls;
  const int* widx;
  int M, N, K;
  float out_scale;
  int act;  // 0 none, 1 tanh-GELU, 2 erf-GELU, 3 ReLU
  int wt;   // W is stored transposed: element (k,n) at W[n*ldw + k]
};

template <typename TA, typename TW, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmP p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int b = blockIdx.z;
  const int wb = p.sW == 0 ? 0 : (p.widx ? p.widx[b] : b);
  const TA* A = reinterpret_cast<const TA*>(p.A) + (int64_t)b * p.sA;
  const TW* W = reinterpret_cast<const TW*>(p.W) + (int64_t)wb * p.sW;
  const TB* bias = p.bias ? reinterpret_cast<const TB*>(p.bias) + (int64_t)wb * p.sBias : nullptr;
  TC* C = reinterpret_cast<TC*>(p.C) + (int64_t)b * p.sC;
  const float* R = p.R ? p.R + (int64_t)b * p.sR : nullptr;

  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ar = tid >> 2, ak = (tid & 3) * 4;   // A tile: 64 rows x 16 k
  const int wk = tid >> 4, wn = (tid & 15) * 4;  // W tile: 16 k x 64 n
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    {
      const int m = m0 + ar;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[ak + j][ar] = (m < p.M) ? to_f(A[(int64_t)m * p.lda + k0 + ak + j]) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn + j;
        Ws[wk][wn + j] = (n < p.N) ? to_f(p.wt ? W[(int64_t)n * p.ldw + k0 + wk] : W[(int64_t)(k0 + wk) * p.ldw + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (bias) v += to_f(bias[n]);
      if (p.out_scale != 1.0f) v *= p.out_scale;
      if (p.act == 1) v = gelu_tanh_f(v);
      else if (p.act == 2) v = gelu_erf_f(v);
      else if (p.act == 3) v = fmaxf(v, 0.f);
      if (p.ls) v *= p.ls[n];
      if (R) v = R[(int64_t)m * p.ldr + n] + v;
      from_f(C[(int64_t)m * p.ldc + n], v);
    }
  }
}

template <typename TA, typename TW, typename TB, typename TC>
inline int gemm_simt(cudaStream_t st, const GemmP& p, int batch) {
  if (p.K % 16 != 0 || p.M <= 0 || p.N <= 0 || batch <= 0) return fail(HVLA_ERR_ARG, "gemm_simt: bad shape");
  dim3 grid(cdiv(p.N, 64), cdiv(p.M, 64), batch);
  ProfScope ps(st, "gemm_simt");
  gemm_simt_kernel<TA, TW, TB, TC><<<grid, 256, 0, st>>>(p);
  HVLA_LAUNCH_CHECK("gemm_simt");
  return HVLA_OK;
}

inline GemmP gemm_params(const void* A, int lda, const void* W, int ldw, const void* bias, void* C, int ldc,
                         int M, int N, int K) {
  GemmP p;
  memset(&p, 0, sizeof p);
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.bias = bias; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.out_scale = 1.0f;
  return p;
}

template <typename TW> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};


// =============================================================================================
// LayerNorm (flax: eps 1e-6, fast variance E[x^2]-E[x]^2 clipped at 0), one warp per row.
// scale/bias may be per-batch (per-sample generated weights).  D in {64,128,768}.
// =============================================================================================
struct LnP {
  const float* x; int64_t ldx;      // row stride (elements)
  void* y; int64_t ldy;
  const void* scale; const void* bias; int64_t sS;   // per weight-batch stride (0 = shared)
  const int* widx; int rows_per_batch;
  int rows; float post_div;          // y /= post_div when != 0
  const float* part; int nsplit; int64_t part_stride;   // 768-wide fp32 path: x += part[s*part_stride + ...], s < nsplit, first
};

template <typename TS, typename TO, int D>
__global__ void __launch_bounds__(128) layernorm_kernel(LnP p) {
  constexpr int PER = D / 32;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const float* x = p.x + (int64_t)row * p.ldx;
  float v[PER];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = x[lane + 32 * i];
    s += v[i];
    s2 = fmaf(v[i], v[i], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float mean = s / (float)D;
  const float mean2 = s2 / (float)D;
  const float var = fmaxf(0.f, mean2 - mean * mean);
  const float rstd = 1.0f / sqrtf(var + 1e-6f);
  int wb = 0;
  if (p.sS != 0) {
    const int b = row / p.rows_per_batch;
    wb = p.widx ? p.widx[b] : b;
  }
  const TS* sc = reinterpret_cast<const TS*>(p.scale) + (int64_t)wb * p.sS;
  const TS* bi = reinterpret_cast<const TS*>(p.bias) + (int64_t)wb * p.sS;
  TO* y = reinterpret_cast<TO*>(p.y) + (int64_t)row * p.ldy;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    float o = (v[i] - mean) * (rstd * to_f(sc[c])) + to_f(bi[c]);
    if (p.post_div != 0.f) o = o / p.post_div;
    from_f(y[c], o);
  }
}

// DINOv2 LayerNorm: 768-wide fp32 rows (contiguous, shared scale/bias), vectorised; one warp per row.
template <typename TO, int NS>
__global__ void __launch_bounds__(256) layernorm768_kernel(float* __restrict__ x, TO* __restrict__ y,
                                                           const float* __restrict__ scale, const float* __restrict__ bias, int rows,
                                                           const float* __restrict__ part, int64_t part_stride) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + (int64_t)row * 768);
  float4 v[6];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = xr[lane + 32 * i];
  if (NS > 0) {          // split-K partial products of the preceding residual GEMM, added in a fixed order (gemm_tc.cuh);
    float4 q[NS > 0 ? NS : 1][6];     // every load is issued before the first add: one memory round trip (small batches are latency-bound)
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
      const float4* pr = reinterpret_cast<const float4*>(part + sp * part_stride + (int64_t)row * 768);
#pragma unroll
      for (int i = 0; i < 6; ++i) q[sp][i] = pr[lane + 32 * i];
    }
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { v[i].x += q[sp][i].x; v[i].y += q[sp][i].y; v[i].z += q[sp][i].z; v[i].w += q[sp][i].w; }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xr[lane + 32 * i] = v[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float mean = s / 768.f;
  const float var = fmaxf(0.f, s2 / 768.f - mean * mean);
  const float rstd = 1.0f / sqrtf(var + 1e-6f);
  TO* yr = y + (int64_t)row * 768;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = scale ? __ldg(reinterpret_cast<const float4*>(scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);   // null: unit scale,
    const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);     // zero bias
    float o[4];
    o[0] = (v[i].x - mean) * (rstd * g.x) + b.x;
    o[1] = (v[i].y - mean) * (rstd * g.y) + b.y;
    o[2] = (v[i].z - mean) * (rstd * g.z) + b.z;
    o[3] = (v[i].w - mean) * (rstd * g.w) + b.w;
    Vec4<TO>::store(yr + c, o);
  }
}

// Shadow + row statistics of the fp32 stream for the LayerNorm-free flow (gemm_tc.cuh): xb = bf16(x) (UN-normalised),
// stats[row] = {(sum, sumsq), 0 x 5}.  Used where no residual-GEMM epilogue produced them: in front of the first layer and
// after a split-K GEMM, whose NS partial-product blocks it folds into the stream first (like layernorm768_kernel).
template <int NS>
__global__ void __launch_bounds__(256) stream_shadow_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float* __restrict__ stats,
                                                            int rows, const float* __restrict__ part, int64_t part_stride) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + (int64_t)row * 768);
  float4 v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = xr[lane + 32 * i];
  if (NS > 0) {
    float4 q[NS > 0 ? NS : 1][6];
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
      const float4* pr = reinterpret_cast<const float4*>(part + sp * part_stride + (int64_t)row * 768);
#pragma unroll
      for (int i = 0; i < 6; ++i) q[sp][i] = pr[lane + 32 * i];
    }
#pragma unroll
    for (int sp = 0; sp < NS; ++sp) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { v[i].x += q[sp][i].x; v[i].y += q[sp][i].y; v[i].z += q[sp][i].z; v[i].w += q[sp][i].w; }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xr[lane + 32 * i] = v[i];
  }
  float s = 0.f, s2 = 0.f;
  __nv_bfloat16* br = xb + (int64_t)row * 768;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
    float o[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    Vec4<__nv_bfloat16>::store(br + (lane + 32 * i) * 4, o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane < 12) stats[(int64_t)row * 12 + lane] = lane == 0 ? s : (lane == 1 ? s2 : 0.f);
}

inline int stream_shadow(cudaStream_t st, float* x, __nv_bfloat16* xb, float* stats, int rows, const float* part, int nsplit, int64_t part_stride) {
  ProfScope ps(st, "layernorm");
  const dim3 grid(cdiv(rows, 8)), block(256);
  switch (nsplit) {
    case 0: launch_k(stream_shadow_kernel<0>, grid, block, 0, st, x, xb, stats, rows, part, part_stride); break;
    case 1: launch_k(stream_shadow_kernel<1>, grid, block, 0, st, x, xb, stats, rows, part, part_stride); break;
    case 3: launch_k(stream_shadow_kernel<3>, grid, block, 0, st, x, xb, stats, rows, part, part_stride); break;
    case 7: launch_k(stream_shadow_kernel<7>, grid, block, 0, st, x, xb, stats, rows, part, part_stride); break;
    default: return fail(HVLA_ERR_ARG, "stream_shadow: unsupported number of split-K partial blocks");
  }
  HVLA_LAUNCH_CHECK("stream_shadow");
  return HVLA_OK;
}

// ---- blocked fp32 stream (gemm_tc.cuh: xblk_f4) -> row-major bf16, one CTA per 32-row block ------------------------------------
// NORMALIZE = false: the un-normalised bf16 shadow + row statistics the folded-LayerNorm GEMM epilogues consume (in front of the
// first layer, where no residual epilogue has produced them yet); NORMALIZE = true: the final LayerNorm (scale / bias) -> embeddings.
// Loads are the blocked layout's 512-byte runs (lane = row); the 32 x 768 tile is transposed through shared memory so that the
// row-major stores are 16-byte, fully coalesced as well.
constexpr int XB_LD = 776;                                 // bf16 row stride of the staged tile (1552 B: conflict-free 8-byte column writes)
template <bool NORMALIZE>
__global__ void __launch_bounds__(256) stream_blk_rows_kernel(const float* __restrict__ xb, __nv_bfloat16* __restrict__ y, float* __restrict__ stats,
                                                              const float* __restrict__ scale, const float* __restrict__ bias, int rows) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t xb_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(xb_smem);                     // [32][XB_LD]
  float* red = reinterpret_cast<float*>(xb_smem + 32 * XB_LD * 2);                     // [8 warps][32 rows][2]
  float* fin = red + 8 * 32 * 2;                                                       // [32 rows][2]: mean, rstd
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(xb) + (int64_t)rb * 192 * 32 + lane;
  float4 v[24];                                            // this warp's column groups cc = warp + 8 i of row `lane`
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = __ldcg(src + (warp + 8 * i) * 32);
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, s2))));
  }
  red[(warp * 32 + lane) * 2] = s;
  red[(warp * 32 + lane) * 2 + 1] = s2;
  __syncthreads();
  if (warp == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[(w * 32 + lane) * 2]; b += red[(w * 32 + lane) * 2 + 1]; }
    const int row = rb * 32 + lane;
    if (!NORMALIZE) {
      if (row < rows) {
        float4* sp = reinterpret_cast<float4*>(stats + (int64_t)row * 12);
        sp[0] = make_float4(a, b, 0.f, 0.f); sp[1] = make_float4(0.f, 0.f, 0.f, 0.f); sp[2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const float mean = a / 768.f;
      fin[lane * 2] = mean;
      fin[lane * 2 + 1] = 1.0f / sqrtf(fmaxf(0.f, b / 768.f - mean * mean) + 1e-6f);
    }
  }
  if (NORMALIZE) __syncthreads();
  const float mean = NORMALIZE ? fin[lane * 2] : 0.f, rstd = NORMALIZE ? fin[lane * 2 + 1] : 1.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = (warp + 8 * i) * 4;
    float o[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    if (NORMALIZE) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(scale + c)), bb = __ldg(reinterpret_cast<const float4*>(bias + c));
      o[0] = (o[0] - mean) * (rstd * g.x) + bb.x; o[1] = (o[1] - mean) * (rstd * g.y) + bb.y;
      o[2] = (o[2] - mean) * (rstd * g.z) + bb.z; o[3] = (o[3] - mean) * (rstd * g.w) + bb.w;
    }
    Vec4<__nv_bfloat16>::store(tile + lane * XB_LD + c, o);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 96; i += 256) {       // 96 16-byte pieces per row
    const int r = i / 96, c8 = (i % 96) * 8;
    if (rb * 32 + r < rows)
      *reinterpret_cast<uint4*>(y + (int64_t)(rb * 32 + r) * 768 + c8) = *reinterpret_cast<const uint4*>(tile + r * XB_LD + c8);
  }
}
constexpr int XB_SMEM = 32 * XB_LD * 2 + 8 * 32 * 2 * 4 + 32 * 2 * 4;

inline int stream_blk_rows(cudaStream_t st, const float* xb, __nv_bfloat16* y, float* stats, const float* scale, const float* bias, int rows) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(stream_blk_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_SMEM));
    HVLA_CUDA(cudaFuncSetAttribute(stream_blk_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_SMEM));
  }
  ProfScope ps(st, "layernorm");
  const dim3 grid(cdiv(rows, 32)), block(256);
  if (scale) launch_k(stream_blk_rows_kernel<true>, grid, block, (size_t)XB_SMEM, st, xb, y, stats, scale, bias, rows);
  else launch_k(stream_blk_rows_kernel<false>, grid, block, (size_t)XB_SMEM, st, xb, y, stats, scale, bias, rows);
  HVLA_LAUNCH_CHECK("stream_blk_rows");
  return HVLA_OK;
}

template <typename TS, typename TO>
inline int layernorm(cudaStream_t st, const LnP& p, int D) {
  if (p.nsplit > 0 && !(D == 768 && std::is_same<TS, float>::value && p.sS == 0 && p.ldx == 768 && p.ldy == 768 && p.post_div == 0.f))
    return fail(HVLA_ERR_ARG, "layernorm: split-K partials need the 768-wide fp32 path");
  if (D == 768 && std::is_same<TS, float>::value && p.sS == 0 && p.ldx == 768 && p.ldy == 768 && p.post_div == 0.f) {
    ProfScope ps(st, "layernorm");
    float* xp = const_cast<float*>(p.x);
    TO* yp = reinterpret_cast<TO*>(p.y);
    const float* sc = reinterpret_cast<const float*>(p.scale);
    const float* bi = reinterpret_cast<const float*>(p.bias);
    const dim3 grid(cdiv(p.rows, 8)), block(256);
    switch (p.nsplit) {
      case 0: launch_k(layernorm768_kernel<TO, 0>, grid, block, 0, st, xp, yp, sc, bi, p.rows, p.part, p.part_stride); break;
      case 1: launch_k(layernorm768_kernel<TO, 1>, grid, block, 0, st, xp, yp, sc, bi, p.rows, p.part, p.part_stride); break;
      case 3: launch_k(layernorm768_kernel<TO, 3>, grid, block, 0, st, xp, yp, sc, bi, p.rows, p.part, p.part_stride); break;
      case 7: launch_k(layernorm768_kernel<TO, 7>, grid, block, 0, st, xp, yp, sc, bi, p.rows, p.part, p.part_stride); break;
      default: return fail(HVLA_ERR_ARG, "layernorm: unsupported number of split-K partial blocks");
    }
    HVLA_LAUNCH_CHECK("layernorm768");
    return HVLA_OK;
  }
  dim3 grid(cdiv(p.rows, 4));
  ProfScope ps(st, "layernorm");
  if (D == 64) layernorm_kernel<TS, TO, 64><<<grid, 128, 0, st>>>(p);
  else if (D == 128) layernorm_kernel<TS, TO, 128><<<grid, 128, 0, st>>>(p);
  else if (D == 768) layernorm_kernel<TS, TO, 768><<<grid, 128, 0, st>>>(p);
  else return fail(HVLA_ERR_ARG, "layernorm: unsupported width");
  HVLA_LAUNCH_CHECK("layernorm");
  return HVLA_OK;
}

// =============================================================================================
// Attention over a packed qkv buffer [rows, 3*H*DH] (q | k | v), one warp per query row.
// q is divided by sqrt(DH) first (flax dot_product_attention_weights), masked scores are
// finfo(f32).min (-FLT_MAX), softmax fp32.
//   mask 0: none (DINOv2)
//   mask 1: base ViT   -- key S-1 (action token) visible only to query S-1 (base_vit.py:209-214)
//   mask 2: context    -- keys 0..31: tok_mask & lang_pad; key 32: 1; key 33: only query 33
//                         (hypernetwork.py:151-181)
// =============================================================================================
struct AttnP {
  const void* qkv; void* out;      // out [rows, H*DH]
  int S, H, nbatch, mask;
  const int32_t* tok_mask;         // [nbatch, 32] (mask 2)
  const uint8_t* lang_pad;         // [nbatch] or null
  int prescaled;                   // q already divided by sqrt(DH)
  int n_act;                       // mask 1: number of trailing action tokens the other tokens cannot see (0 means 1)
};

template <typename T, typename TO, int DH_>
__global__ void __launch_bounds__(128) attention_simt_kernel(AttnP p) {
  constexpr int MAXK = 9;  // keys per lane: S <= 288
  __shared__ float qs[4][DH_];
  __shared__ float ps[4][288];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + w;
  const int h = blockIdx.y, b = blockIdx.z;
  const int D = p.H * DH_, ld = 3 * D;
  if (q >= p.S) return;  // whole warp exits together (q is warp-uniform); no block-level sync below
  const T* base = reinterpret_cast<const T*>(p.qkv) + (int64_t)b * p.S * ld;
  const T* qp = base + (int64_t)q * ld + h * DH_;
  const float inv = sqrtf((float)DH_);
  for (int d = lane; d < DH_; d += 32) qs[w][d] = p.prescaled ? to_f(qp[d]) : to_f(qp[d]) / inv;
  __syncwarp();
  float sc[MAXK];
  float m = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    const int key = lane + 32 * i;
    float s = -FLT_MAX;
    if (key < p.S) {
      const T* kp = base + (int64_t)key * ld + D + h * DH_;
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < DH_; ++d) a = fmaf(qs[w][d], to_f(kp[d]), a);
      bool ok = true;
      if (p.mask == 1) { const int na = p.n_act > 0 ? p.n_act : 1; ok = (key < p.S - na) || (q >= p.S - na); }
      else if (p.mask == 2) {
        if (key < 32) ok = (p.tok_mask[b * 32 + key] != 0) && (p.lang_pad ? p.lang_pad[b] != 0 : true);
        else if (key == 33) ok = (q == 33);
      }
      s = ok ? a : -FLT_MAX;
      m = fmaxf(m, s);
    }
    sc[i] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    const int key = lane + 32 * i;
    if (key < p.S) {
      const float e = expf(sc[i] - m);
      sc[i] = e;
      sum += e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    const int key = lane + 32 * i;
    if (key < p.S) ps[w][key] = sc[i] / sum;
  }
  __syncwarp();
  TO* op = reinterpret_cast<TO*>(p.out) + ((int64_t)b * p.S + q) * D + h * DH_;
  for (int d = lane; d < DH_; d += 32) {
    float a = 0.f;
    const T* vp = base + 2 * D + h * DH_ + d;
    for (int key = 0; key < p.S; ++key) a = fmaf(ps[w][key], to_f(vp[(int64_t)key * ld]), a);
    from_f(op[d], a);
  }
}

// Base-ViT attention of the fp32 paths (HVLA_F32 and HVLA_BF16X3): 257 tokens, 4 heads x 16, per-env q|k|v in fp32 [B*257, 192].
// One CTA per (head, env): K (rows padded to 17 floats) and V of the head live in shared memory, a warp owns every 8th query; lane = key
// for the scores, lane = (key parity, d) for P V.  Same operation order as attention_simt_kernel (q / sqrt(16) first, masked scores =
// finfo.min, d ascending, keys ascending within a parity class); the generic kernel read every K / V row from global memory per
// query and took 4.4 ms for 64 envs x 4 blocks -- more than the twelve DINOv2 blocks of the split-operand flow together.
__global__ void __launch_bounds__(256) base_attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  constexpr int S = BTOK, HD = BHD, D = BH * BHD, LD = 3 * D;
  __shared__ float sk[S][HD + 1];
  __shared__ __align__(16) float sv[S][HD];
  __shared__ float ps[8][S + 7];
  const int h = blockIdx.x, b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (int64_t)b * S * LD + h * HD;
  for (int e = threadIdx.x; e < S * (HD / 4); e += 256) {
    const int r = e >> 2, c4 = e & 3;
    const float4 k4 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)r * LD + D) + c4);
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)r * LD + 2 * D) + c4);
    sk[r][c4 * 4] = k4.x; sk[r][c4 * 4 + 1] = k4.y; sk[r][c4 * 4 + 2] = k4.z; sk[r][c4 * 4 + 3] = k4.w;
    *reinterpret_cast<float4*>(&sv[r][c4 * 4]) = v4;
  }
  __syncthreads();
  for (int q = w; q < S; q += 8) {
    float qv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) qv[d] = __ldg(base + (int64_t)q * LD + d) / 4.0f;      // q / sqrt(16): flax scales the query first
    float sc[9];
    float m = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int key = lane + 32 * i;
      float s = -FLT_MAX;
      if (key < S) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) a = fmaf(qv[d], sk[key][d], a);
        s = ((key != S - 1) || (q == S - 1)) ? a : -FLT_MAX;                             // patches never see the action token (base_vit.py:209-214)
        m = fmaxf(m, s);
      }
      sc[i] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (lane + 32 * i < S) { sc[i] = expf(sc[i] - m); sum += sc[i]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (lane + 32 * i < S) ps[w][lane + 32 * i] = sc[i] / sum;
    __syncwarp();
    const int d = lane & 15, par = lane >> 4;                       // even keys on lanes 0..15, odd keys on lanes 16..31: conflict-free V reads
    float a = 0.f;
    for (int key = par; key < S; key += 2) a = fmaf(ps[w][key], sv[key][d], a);
    a += __shfl_xor_sync(0xffffffffu, a, 16);
    if (par == 0) out[((int64_t)b * S + q) * D + h * HD + d] = a;
    __syncwarp();
  }
}

template <typename T, typename TO>
inline int attention_simt(cudaStream_t st, const AttnP& p, int dh) {
  if (p.S > 288) return fail(HVLA_ERR_ARG, "attention_simt: S too large");
  ProfScope ps(st, "attention_simt");
  if (std::is_same<T, float>::value && std::is_same<TO, float>::value && dh == BHD && p.S == BTOK && p.H == BH && p.mask == 1 && !p.prescaled && p.n_act <= 1) {
    base_attention_f32_kernel<<<dim3(BH, p.nbatch), 256, 0, st>>>(reinterpret_cast<const float*>(p.qkv), reinterpret_cast<float*>(p.out));
    HVLA_LAUNCH_CHECK("base_attention_f32");
    return HVLA_OK;
  }
  dim3 grid(cdiv(p.S, 4), p.H, p.nbatch);
  if (dh == 16) attention_simt_kernel<T, TO, 16><<<grid, 128, 0, st>>>(p);
  else if (dh == 32) attention_simt_kernel<T, TO, 32><<<grid, 128, 0, st>>>(p);
  else if (dh == 64) attention_simt_kernel<T, TO, 64><<<grid, 128, 0, st>>>(p);
  else return fail(HVLA_ERR_ARG, "attention_simt: unsupported head dim");
  HVLA_LAUNCH_CHECK("attention_simt");
  return HVLA_OK;
}

// =============================================================================================
// Small data-movement kernels
// =============================================================================================
// (u8/255 - mean)/std, im2col of the VALID 14x14/14 conv: A0[b*256+p, k], k = (kh,kw,c), K padded to 640.
template <typename TO>
__global__ void im2col_norm_kernel(const uint8_t* __restrict__ img, TO* __restrict__ out, int B) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * NPATCH * PATCH_KP;
  if (idx >= total) return;
  const int k = (int)(idx % PATCH_KP);
  const int64_t row = idx / PATCH_KP;
  const int pidx = (int)(row % NPATCH), b = (int)(row / NPATCH);
  float v = 0.f;
  if (k < PATCH_K) {
    const int kh = k / 42, rem = k % 42, kw = rem / 3, c = rem % 3;
    const int py = pidx / GRID, px = pidx % GRID;
    const uint8_t pix = img[(((int64_t)b * IMG + py * PATCH + kh) * IMG + px * PATCH + kw) * 3 + c];
    const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    v = ((float)pix / 255.0f - mean) / sd;   // base_vit.py:111-114
  }
  from_f(out[idx], v);
}

// The same for the bf16 tensor-core path, as a gather: a CTA owns one row of 16 patches of one image = 14 image rows =
// one contiguous 9,408-byte block, which it stages in shared memory with 16-byte loads; the 3 x 256 possible outputs
// (channel, byte) are tabulated per CTA with the expression above; a thread then produces 8 consecutive k (one 16-byte
// store into the 20,480 contiguous output bytes of the 16 patches) from 8 shared-memory bytes.
__global__ void __launch_bounds__(256) im2col_norm_bf16_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int B) {
  pdl_trigger();
  constexpr int ROWB = IMG * 3;                              // 672 bytes per image row
  constexpr int BLK = PATCH * ROWB;                          // 9,408 bytes: the 14 image rows of this patch row
  __shared__ bf16 lut[3][256];
  __shared__ __align__(16) uint8_t pix[BLK];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, p = i & 255;
    const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    lut[c][p] = __float2bfloat16_rn(((float)p / 255.0f - mean) / sd);
  }
  pdl_wait();
  const int py = blockIdx.x % GRID, b = blockIdx.x / GRID;
  const uint4* src = reinterpret_cast<const uint4*>(img + ((int64_t)b * IMG + py * PATCH) * ROWB);
  for (int i = threadIdx.x; i < BLK / 16; i += blockDim.x) reinterpret_cast<uint4*>(pix)[i] = __ldg(src + i);
  __syncthreads();
  constexpr int GROUPS = PATCH_KP / 8;                       // 80 groups of 8 k per patch
  bf16* dst = out + ((int64_t)b * NPATCH + py * GRID) * PATCH_KP;
  for (int it = threadIdx.x; it < GRID * GROUPS; it += blockDim.x) {
    const int g = it % GROUPS, px = it / GROUPS;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      uint32_t bits = 0;
      if (k < PATCH_K) {
        const int kh = k / 42, rem = k - kh * 42;              // rem = kw * 3 + c
        bits = __bfloat16_as_ushort(lut[rem % 3][pix[kh * ROWB + px * 42 + rem]]);
      }
      if (j & 1) w[j >> 1] |= bits << 16; else w[j >> 1] = bits;
    }
    *reinterpret_cast<uint4*>(dst + (int64_t)px * PATCH_KP + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// X[b,0,:] = cls + pos[0];  X[b,1+p,:] = P[b*256+p,:] + pos[1+p]   (fp32 residual stream)
template <typename TP>
__global__ void dino_assemble_kernel(const TP* __restrict__ P, const float* __restrict__ cls,
                                     const float* __restrict__ pos, float* __restrict__ X, int B) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * DTOK * DD;
  if (idx >= total) return;
  const int c = (int)(idx % DD);
  const int64_t row = idx / DD;
  const int t = (int)(row % DTOK), b = (int)(row / DTOK);
  const float base = t == 0 ? cls[c] : to_f(P[((int64_t)b * NPATCH + (t - 1)) * DD + c]);
  X[idx] = base + pos[t * DD + c];
}

// context tokens (hypernetwork.py:112-147): [T,34,128]
__global__ void ctx_assemble_kernel(const float* __restrict__ TPj, const float* __restrict__ IPj,
                                    const float* __restrict__ task_pos, const float* __restrict__ img_pos,
                                    const float* __restrict__ layer_pos, float* __restrict__ X, int T) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)T * CTOK * CD;
  if (idx >= total) return;
  const int c = (int)(idx % CD);
  const int64_t row = idx / CD;
  const int j = (int)(row % CTOK), t = (int)(row / CTOK);
  float v;
  if (j < LANG) v = TPj[((int64_t)t * LANG + j) * CD + c] + task_pos[j * CD + c];
  else if (j == LANG) v = IPj[(int64_t)t * CD + c] + img_pos[c];
  else v = 0.f + layer_pos[c];
  X[idx] = v;
}

// base-ViT tokens (base_vit.py:182-204): X[b,p,:] = patches + pos;  X[b,256,:] = 0 + pos[256]
template <typename TW>
__global__ void base_assemble_kernel(const float* __restrict__ Pt, const TW* __restrict__ weights,
                                     const int* __restrict__ tidx, float* __restrict__ X, int B, int S, int64_t ngp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * S * BD;
  if (idx >= total) return;
  const int c = (int)(idx % BD);
  const int64_t row = idx / BD;
  const int t = (int)(row % S), b = (int)(row / S);
  const TW* pos = weights + (int64_t)(tidx ? tidx[b] : b) * ngp + GenLayout::pos;
  const float base = t < NPATCH ? Pt[((int64_t)b * NPATCH + t) * BD + c] : 0.f;
  X[idx] = base + to_f(pos[t * BD + c]);
}

// =============================================================================================
// Final encoder_norm on the action token + mix action head (action_heads.py:455-470, 536-537)
// one warp per environment.
// =============================================================================================
template <typename TW>
__global__ void __launch_bounds__(128) mix_head_kernel(const float* __restrict__ X /*[B,257,64]*/,
                                                       const TW* __restrict__ weights, const int* __restrict__ tidx,
                                                       float* __restrict__ action, float* __restrict__ logit_out, int B) {
  __shared__ float hs[4][BD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + w;
  if (b >= B) return;
  const TW* wr = weights + (int64_t)(tidx ? tidx[b] : b) * NGP;
  const float* x = X + ((int64_t)b * BTOK + (BTOK - 1)) * BD;
  float v0 = x[lane], v1 = x[lane + 32];
  float s = v0 + v1, s2 = fmaf(v0, v0, v1 * v1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float mean = s / 64.f, var = fmaxf(0.f, s2 / 64.f - mean * mean);
  const float rstd = 1.0f / sqrtf(var + 1e-6f);
  hs[w][lane] = (v0 - mean) * (rstd * to_f(wr[GenLayout::encn_s + lane])) + to_f(wr[GenLayout::encn_b + lane]);
  hs[w][lane + 32] = (v1 - mean) * (rstd * to_f(wr[GenLayout::encn_s + lane + 32])) + to_f(wr[GenLayout::encn_b + lane + 32]);
  __syncwarp();
  if (lane < NCONT) {
    float a = 0.f;
    for (int k = 0; k < BD; ++k) a = fmaf(hs[w][k], to_f(wr[GenLayout::wc + k * NCONT + lane]), a);
    a += to_f(wr[GenLayout::bc + lane]);
    a = tanhf(a / 5.0f) * 5.0f;
    action[(int64_t)b * (AH * AD) + (lane / 6) * AD + (lane % 6)] = a;
  } else if (lane < NCONT + AH) {
    const int j = lane - NCONT;
    float a = 0.f;
    for (int k = 0; k < BD; ++k) a = fmaf(hs[w][k], to_f(wr[GenLayout::wd + k * AH + j]), a);
    a += to_f(wr[GenLayout::bd + j]);
    action[(int64_t)b * (AH * AD) + j * AD + 6] = a >= 0.f ? 1.0f : 0.0f;
    if (logit_out) logit_out[(int64_t)b * AH + j] = a;
  }
}

// =============================================================================================
// Generation heads: out[T, NGP] = E[T,128] * W[128, NGP] + b   (the 73 Dense heads of
// hypernetwork.py:205-217 as one skinny GEMM).  HBM-bound on W for small T.
// Each thread owns 4 adjacent columns; tasks are processed 8 at a time from smem.
// =============================================================================================
// Debug aid (hvla_act_debug: the `intermediates` the reference sows, hypervla/model.py:135-137): attention probabilities of one layer,
// softmax(q k^T * qscale) as fp32 [nbatch, H, S, S], from the layer's q|k|v buffer [nbatch*S, 3*H*HD].  One warp per (batch, head, query);
// mask_mode 1 = base ViT (rows < S-1 do not see column S-1, base_vit.py:209-214).  Not on the hot path: no tuning.
template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_probs_kernel(const T* __restrict__ qkv, float* __restrict__ probs, int S, int H, int nbatch,
                                                         int mask_mode, float qscale) {
  __shared__ float qs[4][HD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + w;              // (b * H + h) * S + i
  if (row >= (int64_t)nbatch * H * S) return;
  const int i = (int)(row % S), h = (int)((row / S) % H), b = (int)(row / ((int64_t)S * H));
  const int ld = 3 * H * HD;
  const T* q = qkv + ((int64_t)b * S + i) * ld + h * HD;
  for (int d = lane; d < HD; d += 32) qs[w][d] = to_f(q[d]) * qscale;
  __syncwarp();
  float sc[9];                                                  // S <= 288
  float mx = -3.4028234663852886e38f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int j = lane + 32 * t;
    float a = -3.4028234663852886e38f;
    if (j < S) {
      const T* k = qkv + ((int64_t)b * S + j) * ld + H * HD + h * HD;
      a = 0.f;
      for (int d = 0; d < HD; ++d) a = fmaf(qs[w][d], to_f(k[d]), a);
      if (mask_mode == 1 && i < S - 1 && j == S - 1) a = -3.4028234663852886e38f;      // where(mask, s, finfo.min)
    }
    sc[t] = a;
    mx = fmaxf(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    sc[t] = (lane + 32 * t < S) ? expf(sc[t] - mx) : 0.f;
    sum += sc[t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float* out = probs + row * S;
#pragma unroll
  for (int t = 0; t < 9; ++t)
    if (lane + 32 * t < S) out[lane + 32 * t] = sc[t] * inv;
}
template <typename T, int HD>
inline int attn_probs(cudaStream_t st, const T* qkv, float* probs, int S, int H, int nbatch, int mask_mode, float qscale) {
  if (S > 288) return fail(HVLA_ERR_ARG, "attn_probs: S > 288");
  const int64_t rows = (int64_t)nbatch * H * S;
  attn_probs_kernel<T, HD><<<cdiv(rows, 4), 128, 0, st>>>(qkv, probs, S, H, nbatch, mask_mode, qscale);
  HVLA_LAUNCH_CHECK("attn_probs");
  return HVLA_OK;
}

constexpr int HEADS_TT = 8;
template <typename TW, typename TO>
__global__ void __launch_bounds__(256) heads_gemm_kernel(const float* __restrict__ E, const TW* __restrict__ W,
                                                         const float* __restrict__ bias, TO* __restrict__ out, int T,
                                                         const int32_t* __restrict__ rows, int T_max, int64_t ngp) {
  __shared__ float es[HEADS_TT][CD];
  const int64_t col = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  const bool active = col < ngp;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) Vec4<float>::load(bias + col, bv);
  for (int t0 = 0; t0 < T; t0 += HEADS_TT) {
    __syncthreads();
    for (int i = threadIdx.x; i < HEADS_TT * CD; i += 256) {
      const int tt = i / CD, k = i % CD;
      es[tt][k] = (t0 + tt < T) ? E[(int64_t)(t0 + tt) * CD + k] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    float acc[HEADS_TT][4];
#pragma unroll
    for (int tt = 0; tt < HEADS_TT; ++tt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[tt][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < CD; ++k) {
      float w[4];
      Vec4<TW>::load(W + (int64_t)k * ngp + col, w);
#pragma unroll
      for (int tt = 0; tt < HEADS_TT; ++tt) {
        const float e = es[tt][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[tt][j] = fmaf(e, w[j], acc[tt][j]);
      }
    }
#pragma unroll
    for (int tt = 0; tt < HEADS_TT; ++tt) {
      if (t0 + tt < T) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = acc[tt][j] + bv[j];
        const int orow = rows ? __ldg(rows + t0 + tt) : t0 + tt;      // task-switch scheduler: scattered rows of a persistent buffer
        if (orow >= 0 && orow < T_max) Vec4<TO>::store(out + (int64_t)orow * ngp + col, o);
      }
    }
  }
}

}  // namespace hvla
