// Per-sample base network in ONE kernel: image_embedding_projection (768->64) + pos-emb, 4 pre-LN
// encoder blocks (4 heads x 16, mlp 128, tanh-GELU), encoder_norm and the mix action head, with the
// sample's own GENERATED weights (one packed row of 201,500 bf16 per task).
//   reference: hypervla/components/base_vit.py:130-226, transformer.py:127-262,
//              action_heads.py:455-470, 536-537; batching = jax.vmap of scripts/train.py:559-579.
//
// A thread-block cluster of two CTAs per environment (C = 2; C = 1 kept as an A/B switch), 12 warps each.  Everything stays
// on chip between the embedding read and the 28 output floats: the fp32 residual stream (128 patch tokens per CTA) and the
// layer's K/V (all 256 rows, broadcast between the CTAs through distributed shared memory) live in shared memory, the layer's
// weights are staged there once (cp.async), GEMMs are warp-level mma.sync m16n8k16 (bf16 operands, fp32 accumulation) --
// the problem is 64-wide and per-sample, far below a tcgen05 tile.
//
// Structure exploited (base_vit.py:209-214): patch tokens cannot attend to the action token, so the 256 patch rows form an
// unmasked 256-token transformer (16 m-tiles: warps 0..7 of each CTA own one each); the action token is a single query
// row over 256 patch keys + itself, handled in fp32 on CUDA cores by warps 8..11 of rank 0 (one head each).  Only the
// action token is needed from the last block, so patch rows stop after writing that block's K/V.
#pragma once
#include "common.cuh"
#include "attn_mma.cuh"

namespace hvla {
namespace basefused {

using attn::ldsm_x4;
using attn::ldsm_x4_t;
using attn::mma_bf16;
using attn::pack2;
using attn::cp_async16;
using attn::cp_async_commit;
using attn::cp_async_wait;

constexpr int NW = 12, NT = NW * 32;      // warps 0..7: patch rows; warps 8..11: the action token (rank 0), one head each
constexpr int XLD = 72;                 // fp32 residual row stride (floats): conflict-free float2 C-fragment access
constexpr int KLD = 72;                 // bf16 row stride for 64-wide matrices (144 B)
constexpr int W0LD = 136;               // bf16 row stride for the 64x128 Dense_0 kernel
constexpr int PLD = 72;                 // staged projection kernel [768][72]

// shared memory map (bytes)
constexpr int OFF_X = 0;
constexpr int OFF_K = OFF_X + 256 * XLD * 4;            // 73728
constexpr int OFF_V = OFF_K + 256 * KLD * 2;            // +36864
constexpr int OFF_W = OFF_V + 256 * KLD * 2;            // weights region (also: tail of the staged projection kernel)
constexpr int OFF_WQ = OFF_W, OFF_WK = OFF_WQ + 64 * KLD * 2, OFF_WV = OFF_WK + 64 * KLD * 2, OFF_WO = OFF_WV + 64 * KLD * 2;
constexpr int OFF_W0 = OFF_WO + 64 * KLD * 2;
constexpr int OFF_W1 = OFF_W0 + 64 * W0LD * 2;
constexpr int OFF_VEC = OFF_W1 + 128 * KLD * 2;         // fp32 vectors
constexpr int V_LN0S = 0, V_LN0B = 64, V_BQ = 128, V_BK = 192, V_BV = 256, V_BO = 320, V_LN1S = 384, V_LN1B = 448, V_B0 = 512,
              V_B1 = 640, V_COUNT = 704;
constexpr int OFF_ACT = OFF_VEC + V_COUNT * 4;          // action-token scratch (fp32)
constexpr int A_X = 0, A_XN = 64, A_Q = 128, A_K = 192, A_V = 256, A_O = 320, A_H = 384, A_P = 512, A_COUNT = 512 + 4 * 264;   // A_P: one probability row per head
constexpr int OFF_VECB = OFF_ACT + A_COUNT * 4;       // the layer's vectors as they arrive (bf16), converted into VEC
constexpr int SMEM = OFF_VECB + V_COUNT * 2;
constexpr int ELD = 72;                               // staged embedding chunk [256][72] bf16 (two of them fill the X region)
static_assert(2 * 256 * ELD * 2 <= OFF_K, "embedding chunks are double-buffered in the (not yet written) X region");
static_assert(768 * PLD * 2 <= OFF_VEC - OFF_K, "projection kernel staging must fit in the K/V/W region");
static_assert(SMEM <= 232448, "shared memory budget");

// tanh-GELU of the patch rows (transformer.py:66) with tanh.approx.f32 (one MUFU op, 2^-11 relative: the result is rounded to bf16
// right after): tanhf costs ~25 instructions, and this kernel is bound by instruction issue / dependent latency, not by any pipe.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float c = 0.7978845608028654f;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(c * fmaf(0.044715f * x, x * x, x)));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// exp2 of a non-positive argument: one MUFU op (exp2f adds a range check and two scalings per call; 512 calls per thread and layer)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// C[16 x NTL*8] += A[16 x KS*16] * W[k0.., n0..]   (W row-major bf16 in smem, row stride ld elements)
template <int NTL, int KS>
__device__ __forceinline__ void gemm_tile(float (&c)[NTL][4], const uint32_t (&a)[KS][4], uint32_t sW, int ld, int k0, int n0, int lane) {
  const int i = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int np = 0; np < NTL / 2; ++np) {
      uint32_t b0, b1, b2, b3;
      const uint32_t addr = sW + (uint32_t)(((k0 + ks * 16 + (i & 1) * 8 + (lane & 7)) * ld + n0 + (np * 2 + (i >> 1)) * 8) * 2);
      ldsm_x4_t(addr, b0, b1, b2, b3);
      mma_bf16(c[2 * np], a[ks], b0, b1);
      mma_bf16(c[2 * np + 1], a[ks], b2, b3);
    }
  }
}

// C-layout fp32 [16 x 64] -> bf16 A fragments for 4 k-steps
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

__device__ __forceinline__ void load_x(const float* X, int row0, int lane, float (&c)[8][4]) {
  const float* p0 = X + (row0 + (lane >> 2)) * XLD + (lane & 3) * 2;
  const float* p1 = p0 + 8 * XLD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 u = *reinterpret_cast<const float2*>(p0 + j * 8);
    const float2 w = *reinterpret_cast<const float2*>(p1 + j * 8);
    c[j][0] = u.x; c[j][1] = u.y; c[j][2] = w.x; c[j][3] = w.y;
  }
}
__device__ __forceinline__ void store_x(float* X, int row0, int lane, const float (&c)[8][4]) {
  float* p0 = X + (row0 + (lane >> 2)) * XLD + (lane & 3) * 2;
  float* p1 = p0 + 8 * XLD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float2*>(p0 + j * 8) = make_float2(c[j][0], c[j][1]);
    *reinterpret_cast<float2*>(p1 + j * 8) = make_float2(c[j][2], c[j][3]);
  }
}

// flax LayerNorm (eps 1e-6, fast variance) of a C-layout tile, result as bf16 A fragments
__device__ __forceinline__ void ln_tile(const float (&c)[8][4], const float* sc, const float* bi, int lane, uint32_t (&a)[4][4]) {
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s0 += c[j][0] + c[j][1]; q0 = fmaf(c[j][0], c[j][0], fmaf(c[j][1], c[j][1], q0));
    s1 += c[j][2] + c[j][3]; q1 = fmaf(c[j][2], c[j][2], fmaf(c[j][3], c[j][3], q1));
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  const float m0 = s0 * (1.f / 64.f), m1 = s1 * (1.f / 64.f);
  const float r0 = rsqrtf(fmaxf(0.f, q0 * (1.f / 64.f) - m0 * m0) + 1e-6f);
  const float r1 = rsqrtf(fmaxf(0.f, q1 * (1.f / 64.f) - m1 * m1) + 1e-6f);
  float n[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    const float2 g = *reinterpret_cast<const float2*>(sc + col);
    const float2 b = *reinterpret_cast<const float2*>(bi + col);
    n[j][0] = (c[j][0] - m0) * (r0 * g.x) + b.x;
    n[j][1] = (c[j][1] - m0) * (r0 * g.y) + b.y;
    n[j][2] = (c[j][2] - m1) * (r1 * g.x) + b.x;
    n[j][3] = (c[j][3] - m1) * (r1 * g.y) + b.y;
  }
  c_to_a(n, a);
}

__device__ __forceinline__ void add_bias(float (&c)[8][4], const float* b, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 v = *reinterpret_cast<const float2*>(b + j * 8 + (lane & 3) * 2);
    c[j][0] += v.x; c[j][1] += v.y; c[j][2] += v.x; c[j][3] += v.y;
  }
}
__device__ __forceinline__ void zero8(float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
}

// copy a [rows x cols] bf16 matrix from global (dense) to padded smem (row stride ld); cols % 8 == 0
__device__ __forceinline__ void stage_matrix(uint8_t* smem, int off, const bf16* g, int rows, int cols, int ld) {
  const int per_row = cols / 8;
  for (int i = threadIdx.x; i < rows * per_row; i += NT) {
    const int r = i / per_row, c = (i % per_row) * 8;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (int64_t)r * cols + c));
    *reinterpret_cast<uint4*>(smem + off + (r * ld + c) * 2) = v;
  }
}
// the same through cp.async (16-byte pieces, no register round trip): issue everything, wait once
__device__ __forceinline__ void stage_matrix_async(uint32_t sdst, const bf16* g, int rows, int cols, int ld) {
  const int per_row = cols / 8;
  for (int i = threadIdx.x; i < rows * per_row; i += NT) {
    const int r = i / per_row, c = (i % per_row) * 8;
    cp_async16(sdst + (uint32_t)((r * ld + c) * 2), g + (int64_t)r * cols + c);
  }
}
__device__ __forceinline__ void stage_vec_async(uint32_t sdst, const bf16* g, int n) {   // n % 8 == 0
  if (threadIdx.x < n / 8) cp_async16(sdst + threadIdx.x * 16, g + threadIdx.x * 8);
}
__device__ __forceinline__ void stage_vec(float* dst, const bf16* g, int n) {
  for (int i = threadIdx.x; i < n; i += NT) dst[i] = __bfloat162float(g[i]);
}

// fp32 GEMV helper for the action-token warp: out[n] = sum_k x[k] * W[k][n] (W bf16 smem, row stride ld)
template <int K>
__device__ __forceinline__ float gemv_col(const float* x, const bf16* W, int ld, int n) {
  float a = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) a = fmaf(x[k], __bfloat162float(W[k * ld + n]), a);
  return a;
}

// barrier among the four action-token warps (threads 256..383)
__device__ __forceinline__ void act_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// out[n] for n = n0 + (lane & 15): the two 16-lane halves of the warp each sum half of the K range
template <int K>
__device__ __forceinline__ float gemv_col_split(const float* x, const bf16* W, int ld, int n0, int lane) {
  const int n = n0 + (lane & 15), k0 = (lane >> 4) * (K / 2);
  float a = 0.f;
#pragma unroll 8
  for (int k = k0; k < k0 + K / 2; ++k) a = fmaf(x[k], __bfloat162float(W[k * ld + n]), a);
  return a + __shfl_xor_sync(0xffffffffu, a, 16);
}

__device__ __forceinline__ void warp_ln64(const float* x, const float* sc, const float* bi, float* out, int lane) {
  const float v0 = x[lane], v1 = x[lane + 32];
  float s = v0 + v1, q = fmaf(v0, v0, v1 * v1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float m = s / 64.f, var = fmaxf(0.f, q / 64.f - m * m);
  const float r = 1.0f / sqrtf(var + 1e-6f);
  out[lane] = (v0 - m) * (r * sc[lane]) + bi[lane];
  out[lane + 32] = (v1 - m) * (r * sc[lane + 32]) + bi[lane + 32];
}

// phase timestamps of CTA 0 (experiments only: build with HVLA_NVCC_EXTRA=-DHVLA_BASE_TS)
#ifdef HVLA_BASE_TS
#define BASE_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) ts_[i] = clock64(); } while (0)
#else
#define BASE_TS(i) do { } while (0)
#endif

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// copy 16 bytes of this CTA's shared memory to the same offset of CTA `peer` of the cluster (distributed shared memory)
__device__ __forceinline__ void copy16_to_peer(uint32_t local_addr, uint32_t peer) {
  uint32_t r, v0, v1, v2, v3;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(local_addr) : "memory");
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(peer));
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
}
// all warps of the CTA meet on the (sleeping) CTA barrier first, so nobody spins on the cluster barrier while
// a slower warp of the same SM is still working
__device__ __forceinline__ void cluster_sync_after_cta() {
  __syncthreads();
  cluster_barrier();
}

// C = CTAs per environment (a thread-block cluster): each CTA owns 256/C patch rows (residual stream, Q, MLP) and
// broadcasts the K/V rows it computes into every CTA of the cluster through distributed shared memory; the action
// token lives in rank 0.  C = 2 puts 128 CTAs on the 148 SMs at 64 envs and halves the batch-1 latency.
template <int C>
__global__ void __cluster_dims__(C, 1, 1) __launch_bounds__(NT, 1)
base_fused_kernel(const bf16* __restrict__ emb, const bf16* __restrict__ weights, const int* __restrict__ tidx,
                  float* __restrict__ action, float* __restrict__ logit_out) {
  constexpr int ROWS = 256 / C;          // patch rows of this CTA
  constexpr int MT = 2 / C;              // 16-row m-tiles per patch warp
  static_assert(C == 1 || C == 2, "cluster size");
  pdl_trigger();
  pdl_wait();
#ifdef HVLA_BASE_TS
  long long ts_[16];
#endif
  BASE_TS(0);
  extern __shared__ __align__(16) uint8_t smem[];
  typedef GenLayout G;
  const int b = blockIdx.x / C;
  const uint32_t rank = C > 1 ? cluster_rank() : 0u;
  const int grow0 = (int)rank * ROWS;     // first global patch row of this CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* wrow = weights + (int64_t)(tidx ? tidx[b] : b) * NGP;
  const bf16* erow = emb + ((int64_t)b * DTOK + 1) * DD;     // skip the CLS row (base_vit.py:122)
  float* X = reinterpret_cast<float*>(smem + OFF_X);
  float* VEC = reinterpret_cast<float*>(smem + OFF_VEC);
  float* ACT = reinterpret_cast<float*>(smem + OFF_ACT);
  bf16* Ks = reinterpret_cast<bf16*>(smem + OFF_K);
  bf16* Vs = reinterpret_cast<bf16*>(smem + OFF_V);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sK = sbase + OFF_K, sV = sbase + OFF_V;

  // ------------------------------------------------------------------ P0: projection + pos-emb
  // The 768x64 kernel is staged once; the sample's 256x768 embeddings stream through two 256x64 chunks
  // (cp.async, double-buffered in the X region, which is only written after the last chunk is consumed).
  const uint32_t sE = sbase + OFF_X;
  auto stage_emb = [&](int kc) {
    const uint32_t dst = sE + (uint32_t)((kc & 1) * 256 * ELD * 2);
    for (int i = threadIdx.x; i < ROWS * 8; i += NT) {
      const int r = i >> 3, c = (i & 7) * 8;
      cp_async16(dst + (uint32_t)((r * ELD + c) * 2), erow + (int64_t)(grow0 + r) * DD + kc * 64 + c);
    }
  };
  stage_matrix_async(sK, wrow + G::proj_w, DD, BD, PLD);
  stage_emb(0);
  cp_async_commit();
  stage_vec(VEC, wrow + G::proj_b, BD);
  BASE_TS(1);
  float acc[MT][8][4];
#pragma unroll
  for (int t = 0; t < MT; ++t) zero8(acc[t]);
#pragma unroll 1
  for (int kc = 0; kc < DD / 64; ++kc) {
    cp_async_wait<0>();
    __syncthreads();                                   // chunk kc landed; everyone is done with chunk kc-1
    if (kc + 1 < DD / 64) { stage_emb(kc + 1); cp_async_commit(); }
    if (warp < 8) {
      const uint32_t src = sE + (uint32_t)((kc & 1) * 256 * ELD * 2);
      const int i = lane >> 3;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[MT][4];
#pragma unroll
        for (int t = 0; t < MT; ++t)
          ldsm_x4(src + (uint32_t)((((warp * MT + t) * 16 + (lane & 7) + (i & 1) * 8) * ELD + ks * 16 + (i >> 1) * 8) * 2),
                  a[t][0], a[t][1], a[t][2], a[t][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t addr = sK + (uint32_t)(((kc * 64 + ks * 16 + (i & 1) * 8 + (lane & 7)) * PLD + (np * 2 + (i >> 1)) * 8) * 2);
          ldsm_x4_t(addr, b0, b1, b2, b3);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            mma_bf16(acc[t][2 * np], a[t], b0, b1);
            mma_bf16(acc[t][2 * np + 1], a[t], b2, b3);
          }
        }
      }
    }
  }
  __syncthreads();                                     // the X region is free: every warp has consumed the last chunk
  if (warp < 8) {
    const bf16* pos = wrow + G::pos;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      const int row0 = (warp * MT + t) * 16;             // local row in X
      add_bias(acc[t], VEC, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = j * 8 + (lane & 3) * 2;
        const __nv_bfloat162 pa = *reinterpret_cast<const __nv_bfloat162*>(pos + (grow0 + row0 + (lane >> 2)) * BD + col);
        const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(pos + (grow0 + row0 + 8 + (lane >> 2)) * BD + col);
        acc[t][j][0] += __low2float(pa); acc[t][j][1] += __high2float(pa);
        acc[t][j][2] += __low2float(pb); acc[t][j][3] += __high2float(pb);
      }
      store_x(X, row0, lane, acc[t]);
    }
  } else if (rank == 0 && warp == 8) {
    // action token: zeros + pos_embedding[256]  (base_vit.py:182-204)
    const bf16* pos = wrow + G::pos + 256 * BD;
    ACT[A_X + lane] = __bfloat162float(pos[lane]);
    ACT[A_X + lane + 32] = __bfloat162float(pos[lane + 32]);
  }
  __syncthreads();
  BASE_TS(2);

  // ------------------------------------------------------------------ encoder blocks
  for (int l = 0; l < BL; ++l) {
    const bf16* lw = wrow + G::layers + (int64_t)l * G::layer_size;
    stage_matrix_async(sbase + OFF_WQ, lw + G::wq, 64, 64, KLD);
    stage_matrix_async(sbase + OFF_WK, lw + G::wk, 64, 64, KLD);
    stage_matrix_async(sbase + OFF_WV, lw + G::wv, 64, 64, KLD);
    stage_matrix_async(sbase + OFF_WO, lw + G::wo, 64, 64, KLD);
    stage_matrix_async(sbase + OFF_W0, lw + G::w0, 64, 128, W0LD);
    stage_matrix_async(sbase + OFF_W1, lw + G::w1, 128, 64, KLD);
    {
      const uint32_t sv = sbase + OFF_VECB;
      stage_vec_async(sv + V_LN0S * 2, lw + G::ln0_s, 128);     // ln0 scale | bias are adjacent in the row
      stage_vec_async(sv + V_BQ * 2, lw + G::bq, 64); stage_vec_async(sv + V_BK * 2, lw + G::bk, 64);
      stage_vec_async(sv + V_BV * 2, lw + G::bv, 64); stage_vec_async(sv + V_BO * 2, lw + G::bo, 64);
      stage_vec_async(sv + V_LN1S * 2, lw + G::ln1_s, 128);     // ln1 scale | bias
      stage_vec_async(sv + V_B0 * 2, lw + G::b0, 128); stage_vec_async(sv + V_B1 * 2, lw + G::b1, 64);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int i = threadIdx.x; i < V_COUNT; i += NT) VEC[i] = __bfloat162float(reinterpret_cast<const bf16*>(smem + OFF_VECB)[i]);
    __syncthreads();
    BASE_TS(3 + 3 * l);

    // ---- phase A: K and V of every token -------------------------------------------------------
    if (warp < 8) {
#pragma unroll 1
      for (int t = 0; t < MT; ++t) {
        const int row0 = (warp * MT + t) * 16;
        float c[8][4];
        load_x(X, row0, lane, c);
        uint32_t a[4][4];
        ln_tile(c, VEC + V_LN0S, VEC + V_LN0B, lane, a);
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
          zero8(c);
          gemm_tile<8, 4>(c, a, sbase + (kv ? OFF_WV : OFF_WK), KLD, 0, 0, lane);
          add_bias(c, VEC + (kv ? V_BV : V_BK), lane);
          // K/V rows are indexed globally; every CTA of the cluster gets a copy
          const uint32_t dst = (kv ? sV : sK) + (uint32_t)(((grow0 + row0 + (lane >> 2)) * KLD + (lane & 3) * 2) * 2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t lo = pack2(c[j][0], c[j][1]), hi = pack2(c[j][2], c[j][3]);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + j * 16), "r"(lo) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + 8 * KLD * 2 + j * 16), "r"(hi) : "memory");
          }
          if (C > 1) {                                   // the warp's 16 x 64 tile goes to the peer in 16-byte pieces
            __syncwarp();
            const uint32_t tile = (kv ? sV : sK) + (uint32_t)((grow0 + row0) * KLD * 2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int piece = lane + 32 * i;               // 16 rows x 8 pieces
              copy16_to_peer(tile + (uint32_t)((piece >> 3) * KLD * 2 + (piece & 7) * 16), rank ^ 1u);
            }
          }
        }
      }
    } else if (rank == 0) {
      // action token: LayerNorm by warp 8, then q / k / v on warps 8 / 9 / 10
      if (warp == 8) warp_ln64(ACT + A_X, VEC + V_LN0S, VEC + V_LN0B, ACT + A_XN, lane);
      act_bar();
      if (warp < 11) {
        const int j = warp - 8;
        const bf16* Wj = reinterpret_cast<const bf16*>(smem + (j == 0 ? OFF_WQ : (j == 1 ? OFF_WK : OFF_WV)));
        const float* bj = VEC + (j == 0 ? V_BQ : (j == 1 ? V_BK : V_BV));
        float* dst = ACT + (j == 0 ? A_Q : (j == 1 ? A_K : A_V));
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int n = lane + 32 * h2;
          const float v = gemv_col<64>(ACT + A_XN, Wj, KLD, n) + bj[n];
          dst[n] = j == 0 ? v * 0.25f : v;                    // q / sqrt(16)
        }
      }
    }
    if (C > 1) cluster_sync_after_cta(); else __syncthreads();      // all 256 K/V rows are in place (in every CTA)
    BASE_TS(4 + 3 * l);

    // ---- phase B ------------------------------------------------------------------------------------
    if (warp < 8) {
      if (l < BL - 1) {
#pragma unroll 1
        for (int t = 0; t < MT; ++t) {
          const int row0 = (warp * MT + t) * 16;
          float c[8][4];
          load_x(X, row0, lane, c);
          uint32_t a[4][4];
          ln_tile(c, VEC + V_LN0S, VEC + V_LN0B, lane, a);
          float q[8][4];
          zero8(q);
          gemm_tile<8, 4>(q, a, sbase + OFF_WQ, KLD, 0, 0, lane);
          add_bias(q, VEC + V_BQ, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) { q[j][0] *= 0.25f; q[j][1] *= 0.25f; q[j][2] *= 0.25f; q[j][3] *= 0.25f; }   // / sqrt(16)
          uint32_t qa[4][4];
          c_to_a(q, qa);
          constexpr float LOG2E = 1.4426950408889634f;
          // Two heads at a time, stage by stage (scores, max, exp, PV): the warp is latency-bound, two independent
          // dependency chains roughly double its issue rate.  Head h's normalised output IS k-step h of the
          // out-projection's A operand, so it is packed straight into `a` (the LayerNorm fragments are dead by now).
#pragma unroll
          for (int hp = 0; hp < BH / 2; ++hp) {
            float m0[2], m1[2], l0[2], l1[2], oh[2][2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              m0[u] = m1[u] = -INFINITY; l0[u] = l1[u] = 0.f;
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[u][0][e] = oh[u][1][e] = 0.f;
            }
#pragma unroll 1
            for (int kc = 0; kc < 4; ++kc) {
              const int key0 = kc * 64;
              float s[2][8][4];
              const int i = lane >> 3;
#pragma unroll
              for (int np = 0; np < 4; ++np) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int h = 2 * hp + u;
                  s[u][2 * np][0] = s[u][2 * np][1] = s[u][2 * np][2] = s[u][2 * np][3] = 0.f;
                  s[u][2 * np + 1][0] = s[u][2 * np + 1][1] = s[u][2 * np + 1][2] = s[u][2 * np + 1][3] = 0.f;
                  uint32_t b0, b1, b2, b3;
                  const uint32_t addr = sK + (uint32_t)(((key0 + (np * 2 + (i >> 1)) * 8 + (lane & 7)) * KLD + h * 16 + (i & 1) * 8) * 2);
                  ldsm_x4(addr, b0, b1, b2, b3);
                  mma_bf16(s[u][2 * np], qa[h], b0, b1);
                  mma_bf16(s[u][2 * np + 1], qa[h], b2, b3);
                }
              }
              float ms0[2], ms1[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                float mx0 = m0[u], mx1 = m1[u];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                  mx0 = fmaxf(mx0, fmaxf(s[u][nt][0], s[u][nt][1]));
                  mx1 = fmaxf(mx1, fmaxf(s[u][nt][2], s[u][nt][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                const float c0 = ex2_fast((m0[u] - mx0) * LOG2E), c1 = ex2_fast((m1[u] - mx1) * LOG2E);
                m0[u] = mx0; m1[u] = mx1;
                l0[u] *= c0; l1[u] *= c1;
                oh[u][0][0] *= c0; oh[u][0][1] *= c0; oh[u][0][2] *= c1; oh[u][0][3] *= c1;
                oh[u][1][0] *= c0; oh[u][1][1] *= c0; oh[u][1][2] *= c1; oh[u][1][3] *= c1;
                ms0[u] = mx0 * LOG2E; ms1[u] = mx1 * LOG2E;
              }
#pragma unroll
              for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  s[u][nt][0] = ex2_fast(fmaf(s[u][nt][0], LOG2E, -ms0[u])); s[u][nt][1] = ex2_fast(fmaf(s[u][nt][1], LOG2E, -ms0[u]));
                  s[u][nt][2] = ex2_fast(fmaf(s[u][nt][2], LOG2E, -ms1[u])); s[u][nt][3] = ex2_fast(fmaf(s[u][nt][3], LOG2E, -ms1[u]));
                  l0[u] += s[u][nt][0] + s[u][nt][1];
                  l1[u] += s[u][nt][2] + s[u][nt][3];
                }
              }
#pragma unroll
              for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int h = 2 * hp + u;
                  uint32_t pa[4];
                  pa[0] = pack2(s[u][2 * kt][0], s[u][2 * kt][1]); pa[1] = pack2(s[u][2 * kt][2], s[u][2 * kt][3]);
                  pa[2] = pack2(s[u][2 * kt + 1][0], s[u][2 * kt + 1][1]); pa[3] = pack2(s[u][2 * kt + 1][2], s[u][2 * kt + 1][3]);
                  uint32_t b0, b1, b2, b3;
                  const uint32_t addr = sV + (uint32_t)(((key0 + 16 * kt + (i & 1) * 8 + (lane & 7)) * KLD + h * 16 + (i >> 1) * 8) * 2);
                  ldsm_x4_t(addr, b0, b1, b2, b3);
                  mma_bf16(oh[u][0], pa, b0, b1);
                  mma_bf16(oh[u][1], pa, b2, b3);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int h = 2 * hp + u;
              float t0 = l0[u], t1 = l1[u];
              t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
              t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
              const float i0 = 1.0f / t0, i1 = 1.0f / t1;
              a[h][0] = pack2(oh[u][0][0] * i0, oh[u][0][1] * i0);
              a[h][1] = pack2(oh[u][0][2] * i1, oh[u][0][3] * i1);
              a[h][2] = pack2(oh[u][1][0] * i0, oh[u][1][1] * i0);
              a[h][3] = pack2(oh[u][1][2] * i1, oh[u][1][3] * i1);
            }
          }
          // out-projection + residual
          float o[8][4];
          zero8(o);
          gemm_tile<8, 4>(o, a, sbase + OFF_WO, KLD, 0, 0, lane);
          add_bias(o, VEC + V_BO, lane);
          load_x(X, row0, lane, c);
#pragma unroll
          for (int j = 0; j < 8; ++j) { c[j][0] += o[j][0]; c[j][1] += o[j][1]; c[j][2] += o[j][2]; c[j][3] += o[j][3]; }
          // MLP (two halves of the 128 hidden units)
          ln_tile(c, VEC + V_LN1S, VEC + V_LN1B, lane, a);
          float y[8][4];
          zero8(y);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float hdn[8][4];
            zero8(hdn);
            gemm_tile<8, 4>(hdn, a, sbase + OFF_W0, W0LD, 0, hf * 64, lane);
            add_bias(hdn, VEC + V_B0 + hf * 64, lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hdn[j][0] = gelu_tanh_fast(hdn[j][0]); hdn[j][1] = gelu_tanh_fast(hdn[j][1]);
              hdn[j][2] = gelu_tanh_fast(hdn[j][2]); hdn[j][3] = gelu_tanh_fast(hdn[j][3]);
            }
            uint32_t ah[4][4];
            c_to_a(hdn, ah);
            gemm_tile<8, 4>(y, ah, sbase + OFF_W1, KLD, hf * 64, 0, lane);
          }
          add_bias(y, VEC + V_B1, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) { c[j][0] += y[j][0]; c[j][1] += y[j][1]; c[j][2] += y[j][2]; c[j][3] += y[j][3]; }
          store_x(X, row0, lane, c);
        }
      }
    } else if (rank == 0) {
      // ---- action token (fp32, CUDA cores): attends to 256 patch keys + itself ----------------------------
#ifdef HVLA_BASE_TS
      const long long tb0_ = clock64();
#endif
      const int h = warp - 8;                                 // one head per action warp
      float* P = ACT + A_P + h * 264;
      {
        float sc[9];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bf16* kp = Ks + (lane + 32 * i) * KLD + h * 16;
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) a = fmaf(ACT[A_Q + h * 16 + d], __bfloat162float(kp[d]), a);
          sc[i] = a;
          mx = fmaxf(mx, a);
        }
        {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) a = fmaf(ACT[A_Q + h * 16 + d], ACT[A_K + h * 16 + d], a);
          sc[8] = a;
          mx = fmaxf(mx, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = expf(sc[i] - mx); sum += sc[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float pself = expf(sc[8] - mx);
        sum += pself;
        const float inv = 1.0f / sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) P[lane + 32 * i] = sc[i] * inv;
        __syncwarp();
        const int d = lane & 15, half = lane >> 4;
        float acc = 0.f;
        const bf16* vp = Vs + h * 16 + d;
#pragma unroll 4
        for (int j = half * 128; j < half * 128 + 128; ++j) acc = fmaf(P[j], __bfloat162float(vp[j * KLD]), acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (half == 0) ACT[A_O + h * 16 + d] = acc + pself * inv * ACT[A_V + h * 16 + d];
      }
      act_bar();                                              // all four heads' outputs are in ACT[A_O..]
      // out-projection + residual: 16 outputs per warp, the K range split over the two half-warps
      {
        const float v = gemv_col_split<64>(ACT + A_O, reinterpret_cast<const bf16*>(smem + OFF_WO), KLD, h * 16, lane);
        if (lane < 16) ACT[A_X + h * 16 + lane] += v + VEC[V_BO + h * 16 + lane];
      }
      act_bar();
      if (warp == 8) warp_ln64(ACT + A_X, VEC + V_LN1S, VEC + V_LN1B, ACT + A_XN, lane);
      act_bar();
      {                                                        // MLP hidden: 32 of the 128 units per warp
        const int n = h * 32 + lane;
        ACT[A_H + n] = gelu_tanh_f(gemv_col<64>(ACT + A_XN, reinterpret_cast<const bf16*>(smem + OFF_W0), W0LD, n) + VEC[V_B0 + n]);
      }
      act_bar();
      {
        const float v = gemv_col_split<128>(ACT + A_H, reinterpret_cast<const bf16*>(smem + OFF_W1), KLD, h * 16, lane);
        if (lane < 16) ACT[A_X + h * 16 + lane] += v + VEC[V_B1 + h * 16 + lane];
      }
#ifdef HVLA_BASE_TS
      if (blockIdx.x == 0 && warp == 8 && lane == 0) printf("base_ts action-token warps, layer %d phase B: %lld cycles\n", l, clock64() - tb0_);
#endif
    }
    if (C > 1) cluster_sync_after_cta(); else __syncthreads();      // nobody reads this layer's K/V any more
    BASE_TS(5 + 3 * l);
  }

  // ------------------------------------------------------------------ encoder_norm + mix head (warp 8)
  if (warp == 8 && rank == 0) {
    const float v0 = ACT[A_X + lane], v1 = ACT[A_X + lane + 32];
    float s = v0 + v1, q = fmaf(v0, v0, v1 * v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float m = s / 64.f, var = fmaxf(0.f, q / 64.f - m * m);
    const float r = 1.0f / sqrtf(var + 1e-6f);
    ACT[A_XN + lane] = (v0 - m) * (r * __bfloat162float(wrow[G::encn_s + lane])) + __bfloat162float(wrow[G::encn_b + lane]);
    ACT[A_XN + lane + 32] = (v1 - m) * (r * __bfloat162float(wrow[G::encn_s + lane + 32])) + __bfloat162float(wrow[G::encn_b + lane + 32]);
    __syncwarp();
    if (lane < NCONT) {
      float a = 0.f;
      for (int k = 0; k < BD; ++k) a = fmaf(ACT[A_XN + k], __bfloat162float(wrow[G::wc + k * NCONT + lane]), a);
      a += __bfloat162float(wrow[G::bc + lane]);
      action[(int64_t)b * (AH * AD) + (lane / 6) * AD + (lane % 6)] = tanhf(a / 5.0f) * 5.0f;
    } else if (lane < NCONT + AH) {
      const int j = lane - NCONT;
      float a = 0.f;
      for (int k = 0; k < BD; ++k) a = fmaf(ACT[A_XN + k], __bfloat162float(wrow[G::wd + k * AH + j]), a);
      a += __bfloat162float(wrow[G::bd + j]);
      action[(int64_t)b * (AH * AD) + j * AD + 6] = a >= 0.f ? 1.0f : 0.0f;
      if (logit_out) logit_out[(int64_t)b * AH + j] = a;
    }
  }
#ifdef HVLA_BASE_TS
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    printf("base_ts:");
    for (int i = 1; i < 15; ++i) printf(" %lld", ts_[i] - ts_[i - 1]);
    printf("\n");
  }
#endif
}

inline int base_act_bf16(cudaStream_t st, const bf16* emb, const bf16* weights, const int* tidx, int B, float* action, float* logit) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(base_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    HVLA_CUDA(cudaFuncSetAttribute(base_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  static const int force = getenv("HVLA_BASE_CLUSTER") ? atoi(getenv("HVLA_BASE_CLUSTER")) : 0;   // experiments: 1 or 2 CTAs per environment
  const int c = force ? force : 2;
  ProfScope ps(st, "base_fused");
  if (c == 2) launch_k(base_fused_kernel<2>, dim3(B * 2), dim3(NT), (size_t)SMEM, st, emb, weights, tidx, action, logit);
  else launch_k(base_fused_kernel<1>, dim3(B), dim3(NT), (size_t)SMEM, st, emb, weights, tidx, action, logit);
  HVLA_LAUNCH_CHECK("base_fused");
  return HVLA_OK;
}

}  // namespace basefused
}  // namespace hvla
