// Hypernetwork context encoder in ONE kernel (fp16-operand tensor-core path of HVLA_BF16 mode): token / initial-image projections +
// position embeddings, the 34-token block-masked 6-layer pre-LN transformer (4 heads x 32, mlp 512, tanh-GELU),
// encoder_norm on the layer token and the 1/sqrt(128) scaling.
//   reference: hypervla/components/hypernetwork.py:99-197, transformer.py:127-262.
//
// One CTA per task, 8 warps.  The fp32 residual stream (34 x 128), the fp16 A operands and q|k|v / the MLP
// hidden stay in shared memory for the whole network; the (task-shared) weights are streamed from L2 in K chunks through a
// 4-deep cp.async ring shared by consecutive GEMMs and consumed by warp-level mma.sync m16n8k16 (M = 34 rows is far below a
// tcgen05 tile; the kernel is latency-bound, the 2.8 MB of fp16 weights per task come from L2); the block-masked attention
// runs as 12 (head, m-tile) mma.sync units.
#pragma once
#include "common.cuh"
#include "attn_mma.cuh"
#include <cuda_fp16.h>

namespace hvla {
namespace ctxf {

using attn::cp_async16;
using attn::cp_async_commit;
using attn::cp_async_wait;
using attn::ldsm_x4;
using attn::ldsm_x4_t;

// fp16 operands (fp32 accumulate): LN-normalised activations and O(0.1) weights fit fp16 comfortably and its 11-bit
// mantissa keeps the context embedding ~8x closer to the fp32 reference than bf16 would .
typedef __half lp;
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack2(const void* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }

// tanh-GELU (transformer.py:66) with the MUFU tanh (relative error < 2^-11, below the fp16 rounding of the stored hidden)
__device__ __forceinline__ float gelu_tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.7978845608028654f * (x + 0.044715f * (x * x * x))));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

constexpr int NTH = 256, MT = 3;                         // 8 warps; 3 m-tiles = 48 rows (34 used)
constexpr int XLD = 132, ALD = 136, QLD = 392, HLD = 520, ELD = 776, WCH = 32;
constexpr int OFF_X = 0;                                 // fp32 [48][132]
constexpr int OFF_A = OFF_X + 48 * XLD * 4;              // fp16 [48][136]  LN output / attention output
constexpr int OFF_BIG = OFF_A + 48 * ALD * 2;            // fp16: q|k|v [48][392]  |  hidden [48][520]  |  token emb [32][776]
constexpr int BIG_BYTES = 48 * HLD * 2;
constexpr int NSTG = 4;                                  // weight-chunk ring: 3 chunks in flight while one is consumed
constexpr int OFF_W = OFF_BIG + BIG_BYTES;               // NSTG x weight chunk [32][<=520] fp16
constexpr int WBUF_BYTES = 128 * ALD * 2;               // holds a [32][520] chunk (N = 512) or a [128][136] chunk (N = 128)
static_assert(WCH * HLD * 2 <= WBUF_BYTES, "ring buffer");
constexpr int OFF_MISC = OFF_W + NSTG * WBUF_BYTES;      // key-valid flags [34]
constexpr int SMEM = OFF_MISC + 256;
static_assert(SMEM <= 232448, "shared memory budget");
static_assert(32 * ELD * 2 <= BIG_BYTES && 48 * QLD * 2 <= BIG_BYTES, "BIG region");

// C[48 x N] (+)= A[48 x K] (fp16 smem, row stride lda) * W[K x N] (fp16 global, row-major), warp w owns N/8 columns.
// epi(row, col, v0, v1) is called for column pairs (col, col+1) of rows < 34.
// The weight-chunk ring is shared by consecutive GEMMs: `g` counts chunks since kernel start (buffer = g % NSTG), and the
// last NSTG-1 iterations of a GEMM stage the first chunks of the NEXT one (next_W / next_N / next_K), so that its pipeline is
// already full when it starts instead of paying an L2 round trip per GEMM (26 GEMMs per task).
struct NextW {
  const lp* W;
  int N, K, rows;                                        // rows per chunk of that GEMM (32, or 128 for the 128-wide deep ones)
};
// chunk rows: the deep, narrow GEMMs (K >= 512, N = 128) take 128-row chunks -- a quarter of the barrier/stage rounds and
// four times the bytes in flight (the ring is bound by L2 latency, three chunks in flight)
template <int N, int K>
struct ChunkRows { static constexpr int value = (K >= 512 && N == 128) ? 128 : WCH; };

__device__ __forceinline__ void stage_chunk(uint32_t sW, int g, const lp* Wg, int N, int c, int rows) {
  const uint32_t dst = sW + (uint32_t)((g % NSTG) * WBUF_BYTES);
  const lp* src = Wg + (int64_t)c * rows * N;
  const int per_row = N / 8, wld = N + 8;
  for (int i = threadIdx.x; i < rows * per_row; i += NTH) {
    const int r = i / per_row, cc = (i - r * per_row) * 8;
    cp_async16(dst + (uint32_t)((r * wld + cc) * 2), src + (int64_t)r * N + cc);
  }
}

template <int N, int K, class Epi>
__device__ __forceinline__ void cta_gemm(uint8_t* smem, const lp* As, int lda, const lp* __restrict__ Wg, Epi epi, int& g, bool prefetched,
                                         NextW next) {
  constexpr int NTW = N / 64;                            // n-tiles (8 columns) per warp
  constexpr int WLD = N + 8;
  constexpr int CR = ChunkRows<N, K>::value;
  constexpr int NCH = K / CR;
  static_assert(NCH >= NSTG - 1, "every GEMM has at least NSTG-1 chunks");
  static_assert(CR * WLD * 2 <= WBUF_BYTES, "chunk fits a ring buffer");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sW = (uint32_t)__cvta_generic_to_shared(smem + OFF_W);
  const uint32_t sA = (uint32_t)__cvta_generic_to_shared(As);
  float acc[MT][NTW][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NTW; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  // One commit group per staged slot (empty when there is nothing to stage) keeps the wait_group arithmetic uniform.
  const int g0 = g;
  if (!prefetched) {
#pragma unroll
    for (int c = 0; c < NSTG - 1; ++c) { stage_chunk(sW, g0 + c, Wg, N, c, CR); cp_async_commit(); }
  }
  const int next_nch = next.W ? next.K / next.rows : 0;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    cp_async_wait<NSTG - 2>();                           // chunk c has landed (this thread's copies) ...
    __syncthreads();                                     // ... and everyone's; everyone is also done with the previous chunk
    {                                                    // refill the buffer the previous chunk was read from
      const int cn = c + NSTG - 1;
      if (cn < NCH) stage_chunk(sW, g0 + cn, Wg, N, cn, CR);
      else if (cn - NCH < next_nch) stage_chunk(sW, g0 + cn, next.W, next.N, cn - NCH, next.rows);
      cp_async_commit();
    }
    const uint32_t wb = sW + (uint32_t)(((g0 + c) % NSTG) * WBUF_BYTES);
#pragma unroll
    for (int ks = 0; ks < CR / 16; ++ks) {
      uint32_t a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m)
        ldsm_x4(sA + (uint32_t)(((m * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * lda + c * CR + ks * 16 + (lane >> 4) * 8) * 2), a[m][0], a[m][1],
                a[m][2], a[m][3]);
#pragma unroll
      for (int np = 0; np < NTW / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int i = lane >> 3;
        ldsm_x4_t(wb + (uint32_t)(((ks * 16 + (i & 1) * 8 + (lane & 7)) * WLD + warp * (N / 8) + (np * 2 + (i >> 1)) * 8) * 2), b0, b1, b2, b3);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          mma_f16(acc[m][2 * np], a[m], b0, b1);
          mma_f16(acc[m][2 * np + 1], a[m], b2, b3);
        }
      }
    }
  }
  g = g0 + NCH;
  // no trailing barrier: the next refill of the buffer read last happens after the first barrier of the next GEMM
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NTW; ++n) {
      const int col = warp * (N / 8) + n * 8 + (lane & 3) * 2;
      const int r0 = m * 16 + (lane >> 2);
      if (r0 < CTOK) epi(r0, col, acc[m][n][0], acc[m][n][1]);
      if (r0 + 8 < CTOK) epi(r0 + 8, col, acc[m][n][2], acc[m][n][3]);
    }
}

// flax LayerNorm of the 34 residual rows -> fp16 A operand (one warp per row, 4 values per lane)
__device__ __forceinline__ void ln_rows(const float* X, lp* A, const float* __restrict__ sc, const float* __restrict__ bi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 g = __ldg(reinterpret_cast<const float4*>(sc) + lane), b = __ldg(reinterpret_cast<const float4*>(bi) + lane);
  for (int r = warp; r < CTOK; r += 8) {
    const float4 v = *reinterpret_cast<const float4*>(X + r * XLD + lane * 4);
    float s = (v.x + v.y) + (v.z + v.w);
    float q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float mean = s * (1.f / 128.f);
    const float rstd = 1.0f / sqrtf(fmaxf(0.f, q * (1.f / 128.f) - mean * mean) + 1e-6f);
    const uint32_t p0 = pack2((v.x - mean) * (rstd * g.x) + b.x, (v.y - mean) * (rstd * g.y) + b.y);
    const uint32_t p1 = pack2((v.z - mean) * (rstd * g.z) + b.z, (v.w - mean) * (rstd * g.w) + b.w);
    *reinterpret_cast<uint2*>(A + r * ALD + lane * 4) = make_uint2(p0, p1);
  }
}

#ifdef HVLA_CTX_TS
#define CTX_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 24) ts_[i] = clock64(); } while (0)
#else
#define CTX_TS(i) do { } while (0)
#endif

__global__ void __launch_bounds__(NTH, 1)
ctx_fused_kernel(const float* __restrict__ hn, const lp* __restrict__ hnb, const float* __restrict__ tok_emb,
                 const int32_t* __restrict__ tok_mask, const uint8_t* __restrict__ lang_pad, const float* __restrict__ init_cls,
                 float* __restrict__ out_ctx) {
  extern __shared__ __align__(16) uint8_t smem[];
#ifdef HVLA_CTX_TS
  long long ts_[24];
#endif
  CTX_TS(0);
  typedef HnLayout L;
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* X = reinterpret_cast<float*>(smem + OFF_X);
  lp* A = reinterpret_cast<lp*>(smem + OFF_A);
  lp* BIG = reinterpret_cast<lp*>(smem + OFF_BIG);
  int* kvalid = reinterpret_cast<int*>(smem + OFF_MISC);

  // ---- inputs: token embeddings -> fp16 A operand [32][776]; key-valid flags ----
  for (int i = threadIdx.x; i < LANG * (LANGD / 4); i += NTH) {
    const int r = i / (LANGD / 4), c = (i % (LANGD / 4)) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(tok_emb + ((int64_t)t * LANG + r) * LANGD + c));
    *reinterpret_cast<uint2*>(BIG + r * ELD + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
  }
  // (rows 32..47 of this A operand alias the weight-chunk buffers: their products land in output rows that are never read)
  if (threadIdx.x < CTOK) {
    const int j = threadIdx.x;
    const bool pad = lang_pad ? lang_pad[t] != 0 : true;
    kvalid[j] = j < LANG ? ((tok_mask[t * LANG + j] != 0) && pad) : 1;          // hypernetwork.py:151-163 (key 33: see attention)
  }
  for (int i = threadIdx.x; i < 48 * XLD; i += NTH) X[i] = 0.f;
  for (int i = threadIdx.x; i < 48 * ALD / 2; i += NTH) reinterpret_cast<uint32_t*>(A)[i] = 0u;
  __syncthreads();
  // ---- K1: task tokens = tok_emb * W + b + task_pos (hypernetwork.py:112-115) --------------------------------------
  CTX_TS(1);
  int ring = 0;                                          // chunks staged since kernel start (ring position)
  cta_gemm<CD, LANGD>(smem, BIG, ELD, hnb + L::tok_w, [&](int r, int c, float v0, float v1) {
    if (r < LANG) {
      X[r * XLD + c] = v0 + hn[L::tok_b + c] + hn[L::task_pos + r * CD + c];
      X[r * XLD + c + 1] = v1 + hn[L::tok_b + c + 1] + hn[L::task_pos + r * CD + c + 1];
    }
  }, ring, false, NextW{hnb + L::img_w, CD, DD, ChunkRows<CD, DD>::value});
  // initial-image token (row 32): cls * W + b + pos (hypernetwork.py:126-127) as a second small GEMM whose only
  // live row is row 0 of the A operand; layer token (row 33) = 0 + pos (hypernetwork.py:144-145)
  __syncthreads();
  for (int i = threadIdx.x; i < DD / 4; i += NTH) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(init_cls + (int64_t)t * DD) + i);
    *reinterpret_cast<uint2*>(BIG + i * 4) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
  }
  if (threadIdx.x < CD) X[33 * XLD + threadIdx.x] = hn[L::layer_pos + threadIdx.x];
  __syncthreads();
  cta_gemm<CD, DD>(smem, BIG, ELD, hnb + L::img_w, [&](int r, int c, float v0, float v1) {
    if (r == 0) {
      X[32 * XLD + c] = v0 + hn[L::img_b + c] + hn[L::img_pos + c];
      X[32 * XLD + c + 1] = v1 + hn[L::img_b + c + 1] + hn[L::img_pos + c + 1];
    }
  }, ring, true, NextW{hnb + L::layers + L::wqkv, 3 * CD, CD, WCH});
  __syncthreads();

  CTX_TS(2);
  // ---- K2: 6 encoder blocks ----------------------------------------------------------------------------------------
  for (int l = 0; l < CL; ++l) {
    const float* lw = hn + L::layers + (int64_t)l * L::layer_size;
    const lp* lb = hnb + L::layers + (int64_t)l * L::layer_size;
    ln_rows(X, A, lw + L::ln0_s, lw + L::ln0_b);
    __syncthreads();
    if (l < 2) CTX_TS(3 + 8 * l);
    // q|k|v = LN(x) Wqkv + b; q pre-divided by sqrt(32)
    cta_gemm<3 * CD, CD>(smem, A, ALD, lb + L::wqkv, [&](int r, int c, float v0, float v1) {
      v0 += lw[L::bqkv + c];
      v1 += lw[L::bqkv + c + 1];
      if (c < CD) { v0 *= 0.17677669529663687f; v1 *= 0.17677669529663687f; }
      *reinterpret_cast<uint32_t*>(BIG + r * QLD + c) = pack2(v0, v1);
    }, ring, true, NextW{lb + L::wo, CD, CD, WCH});
    __syncthreads();
    if (l < 2) CTX_TS(4 + 8 * l);
    // attention with warp-level mma: a unit is (head, 16-query m-tile), 12 units over the 8 warps; the 48 key slots
    // (34 live) are one chunk.  Block mask of hypernetwork.py:151-181: keys 0..31 = token mask & language pad, key 32
    // (initial image) always visible, key 33 (layer token) only to query 33.  Rows 34..47 of V are zeroed first: their
    // probabilities are exactly 0, but 0 x (stale fp16 Inf/NaN) would still poison the accumulator.
    for (int i = threadIdx.x; i < (48 - CTOK) * (CD / 8); i += NTH) {
      const int r = CTOK + i / (CD / 8), c = (i % (CD / 8)) * 8;
      *reinterpret_cast<uint4*>(BIG + r * QLD + 2 * CD + c) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    {
      const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(BIG);
      const int i4 = lane >> 3;
      for (int unit = warp; unit < CH * MT; unit += 8) {
        const int h = unit / MT, mt = unit % MT;
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          ldsm_x4(sQ + (uint32_t)(((mt * 16 + (lane & 7) + (i4 & 1) * 8) * QLD + h * CHD + ks * 16 + (i4 >> 1) * 8) * 2), qa[ks][0], qa[ks][1],
                  qa[ks][2], qa[ks][3]);
        float sc[6][4];
#pragma unroll
        for (int np = 0; np < 3; ++np) {
          sc[2 * np][0] = sc[2 * np][1] = sc[2 * np][2] = sc[2 * np][3] = 0.f;
          sc[2 * np + 1][0] = sc[2 * np + 1][1] = sc[2 * np + 1][2] = sc[2 * np + 1][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sQ + (uint32_t)((((np * 2 + (i4 >> 1)) * 8 + (lane & 7)) * QLD + CD + h * CHD + ks * 16 + (i4 & 1) * 8) * 2), b0, b1, b2, b3);
            mma_f16(sc[2 * np], qa[ks], b0, b1);
            mma_f16(sc[2 * np + 1], qa[ks], b2, b3);
          }
        }
        const int r0 = mt * 16 + (lane >> 2);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = nt * 8 + (lane & 3) * 2 + e;
            const bool vis = key < LANG ? kvalid[key] != 0 : key == LANG;
            const bool v0 = vis || (key == CTOK - 1 && r0 == CTOK - 1), v1 = vis || (key == CTOK - 1 && r0 + 8 == CTOK - 1);
            sc[nt][e] = v0 ? sc[nt][e] : -INFINITY;
            sc[nt][2 + e] = v1 ? sc[nt][2 + e] : -INFINITY;
            mx0 = fmaxf(mx0, sc[nt][e]);
            mx1 = fmaxf(mx1, sc[nt][2 + e]);
          }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            sc[nt][e] = sc[nt][e] == -INFINITY ? 0.f : __expf(sc[nt][e] - mx0);
            sc[nt][2 + e] = sc[nt][2 + e] == -INFINITY ? 0.f : __expf(sc[nt][2 + e] - mx1);
            l0 += sc[nt][e];
            l1 += sc[nt][2 + e];
          }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          uint32_t pa[4];
          pa[0] = pack2(sc[2 * kt][0], sc[2 * kt][1]); pa[1] = pack2(sc[2 * kt][2], sc[2 * kt][3]);
          pa[2] = pack2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]); pa[3] = pack2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
          for (int dp = 0; dp < 2; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(sQ + (uint32_t)(((16 * kt + (i4 & 1) * 8 + (lane & 7)) * QLD + 2 * CD + h * CHD + (dp * 2 + (i4 >> 1)) * 8) * 2), b0, b1, b2, b3);
            mma_f16(o[2 * dp], pa, b0, b1);
            mma_f16(o[2 * dp + 1], pa, b2, b3);
          }
        }
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = h * CHD + nt * 8 + (lane & 3) * 2;
          if (r0 < CTOK) *reinterpret_cast<uint32_t*>(A + r0 * ALD + col) = pack2(o[nt][0] * inv0, o[nt][1] * inv0);
          if (r0 + 8 < CTOK) *reinterpret_cast<uint32_t*>(A + (r0 + 8) * ALD + col) = pack2(o[nt][2] * inv1, o[nt][3] * inv1);
        }
      }
    }
    __syncthreads();
    if (l < 2) CTX_TS(5 + 8 * l);
    // x += attn Wo + b
    cta_gemm<CD, CD>(smem, A, ALD, lb + L::wo, [&](int r, int c, float v0, float v1) {
      X[r * XLD + c] += v0 + lw[L::bo + c];
      X[r * XLD + c + 1] += v1 + lw[L::bo + c + 1];
    }, ring, true, NextW{lb + L::w0, CF, CD, WCH});
    __syncthreads();
    if (l < 2) CTX_TS(6 + 8 * l);
    ln_rows(X, A, lw + L::ln1_s, lw + L::ln1_b);
    __syncthreads();
    if (l < 2) CTX_TS(7 + 8 * l);
    // h = gelu_tanh(LN(x) W0 + b0)
    cta_gemm<CF, CD>(smem, A, ALD, lb + L::w0, [&](int r, int c, float v0, float v1) {
      *reinterpret_cast<uint32_t*>(BIG + r * HLD + c) = pack2(gelu_tanh_approx(v0 + lw[L::b0 + c]), gelu_tanh_approx(v1 + lw[L::b0 + c + 1]));
    }, ring, true, NextW{lb + L::w1, CD, CF, ChunkRows<CD, CF>::value});
    __syncthreads();
    if (l < 2) CTX_TS(8 + 8 * l);
    // x += h W1 + b1
    cta_gemm<CD, CF>(smem, BIG, HLD, lb + L::w1, [&](int r, int c, float v0, float v1) {
      X[r * XLD + c] += v0 + lw[L::b1 + c];
      X[r * XLD + c + 1] += v1 + lw[L::b1 + c + 1];
    }, ring, true, l + 1 < CL ? NextW{lb + L::layer_size + L::wqkv, 3 * CD, CD, WCH} : NextW{nullptr, 0, 0, WCH});
    __syncthreads();
    if (l < 2) CTX_TS(9 + 8 * l);
  }
  // ---- encoder_norm on the layer token, / sqrt(128)  (hypernetwork.py:188-192) --------------------------------------
  if (warp == 0) {
    const float4 v = *reinterpret_cast<const float4*>(X + (CTOK - 1) * XLD + lane * 4);
    float s = (v.x + v.y) + (v.z + v.w);
    float q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float mean = s * (1.f / 128.f);
    const float rstd = 1.0f / sqrtf(fmaxf(0.f, q * (1.f / 128.f) - mean * mean) + 1e-6f);
    const float4 g = __ldg(reinterpret_cast<const float4*>(hn + L::encn_s) + lane), b = __ldg(reinterpret_cast<const float4*>(hn + L::encn_b) + lane);
    const float d = sqrtf((float)CD);
    float4 o;
    o.x = ((v.x - mean) * (rstd * g.x) + b.x) / d;
    o.y = ((v.y - mean) * (rstd * g.y) + b.y) / d;
    o.z = ((v.z - mean) * (rstd * g.z) + b.z) / d;
    o.w = ((v.w - mean) * (rstd * g.w) + b.w) / d;
    *reinterpret_cast<float4*>(out_ctx + (int64_t)t * CD + lane * 4) = o;
  }
#ifdef HVLA_CTX_TS
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    printf("ctx_ts (in, tok+img proj, | ln0 qkv attn wo ln1 w0 w1 | ...):");
    for (int i = 1; i < 18; ++i) printf(" %lld", ts_[i] - ts_[i - 1]);
    printf(" total %lld\n", clock64() - ts_[0]);
  }
#endif
}

inline int ctx_encode_lp(cudaStream_t st, const float* hn, const lp* hnb, const float* tok_emb, const int32_t* tok_mask,
                           const uint8_t* lang_pad, const float* init_cls, int T, float* out_ctx) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(ctx_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  ProfScope ps(st, "ctx_fused");
  ctx_fused_kernel<<<T, NTH, SMEM, st>>>(hn, hnb, tok_emb, tok_mask, lang_pad, init_cls, out_ctx);
  HVLA_LAUNCH_CHECK("ctx_fused");
  return HVLA_OK;
}

}  // namespace ctxf
}  // namespace hvla
