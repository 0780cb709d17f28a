// DINOv2 self-attention on tcgen05 tensor cores (257 tokens, 12 heads x 64, q pre-divided by sqrt(64)).
//
// Per (image, head) item the 257x257 problem is split so that the tensor-core part is a clean 256x256:
//   * query rows 0..255 = two M=128 tiles; keys 0..255 = one N=256 MMA  (S = Q K^T, fp32 in TMEM);
//   * key 256 (the last patch token) is a rank-1 column and query row 256 a single row: both are done by helper warps with
//     warp-level mma.sync matrix-vector products from the staged shared-memory tiles.
// P = exp2(..) goes back to TMEM as bf16 and O = P V runs with A from TMEM (tcgen05.mma TS form) and V as an MN-major
// shared-memory operand.  Measured facts that shaped the design (profiles/README.md): MUFU.EX2 is 16 results/clk/SM whatever the
// operand type, TMEM reads are ~64 B/clk per scheduler, and with four roles on every scheduler the kernel is bound by instruction
// issue and hand-off latency rather than by any single pipe -- hence packed fp32x2 math, mma.sync helpers and TMA stores.
#pragma once
#include "gemm_tc.cuh"
#include "attn_mma.cuh"
#include <stdlib.h>

namespace hvla {
namespace attn_tc {

using namespace tc;

constexpr int S_ = DTOK;                       // 257
constexpr int TILE_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int KV_BYTES = 272 * 128;            // keys 0..271 (256 = last real key, 257.. = padding rows never read as keys)
constexpr int OFF_K = 2 * TILE_BYTES, OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_QT = OFF_V + KV_BYTES;        // query rows 256..271 (only row 256 is used)
constexpr int STAGE_BYTES = OFF_QT + 16 * 128; // Q0 Q1 | K[272] | V[272] | Qtail[16] = 102 KB
static_assert(STAGE_BYTES % 1024 == 0 && OFF_V % 1024 == 0, "UMMA / TMA 128B-swizzle tiles need 1024-byte alignment");
constexpr float LOG2E = 1.4426950408889634f;
#ifndef HVLA_ATTN_NPOLY
#define HVLA_ATTN_NPOLY 0
#endif
constexpr int NPOLY = HVLA_ATTN_NPOLY;            // of every 8 key pairs, this many take the polynomial exp2 (0 = all MUFU, the default:
                                                 // measured with 2 and 3 of 8, the kernel got 10-14 % SLOWER -- the extra 10 instructions per pair
                                                 // cost more issue slots and registers than the freed MUFU slots return; A/B switch only)

__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major B operand (V[key][d], 128-byte rows, SWIZZLE_128B): 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                    // leading byte offset: unused (N = 64 is one swizzle atom wide)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-key groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
constexpr uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }   // B is MN-major

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): x = n + f with n = rint(x) (magic-number add), f in [-0.5, 0.5], 2^f by a degree-3
// minimax polynomial (relative error 7.6e-5, far below the bf16 rounding of P), 2^n by adding n to the exponent field.  Two elements at
// a time on the packed fp32 pipe: 3 FADD2 + 3 FFMA2 + 2 FMNMX + 2 IMAD per pair.  The exponential pass is MUFU-bound (89 % of the
// 16 results / clk / SM while both softmax groups are in it); a share of the keys can go this way instead (HVLA_ATTN_NPOLY).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
  const float2 t = __fadd2_rn(x, magic);                      // low mantissa bits of t = rint(x) (two's complement)
  const float2 n = __fadd2_rn(t, nmagic);
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.05520550534129143f, 0.05520550534129143f), f, make_float2(0.24261397123336792f, 0.24261397123336792f));
  p = __ffma2_rn(p, f, make_float2(0.6932547688484192f, 0.6932547688484192f));
  p = __ffma2_rn(p, f, make_float2(0.9999276995658875f, 0.9999276995658875f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// The K-step offsets are added INSIDE the asm block so that the compiler cannot hoist 24 pre-stepped descriptors out of the item loop
// (they would not fit the issuing warp's 40 registers).
template <bool ACC, int KOFF>
__device__ __forceinline__ void umma_ss2(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .b64 da, db;\n\t"
      "add.u32 ta, %1, %5;\n\t"
      "add.u32 tb, %2, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, %6;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "n"(KOFF), "n"(ACC ? 1 : 0)
      : "memory");
}
template <bool ACC, int AOFF, int BOFF>
__device__ __forceinline__ void umma_ts2(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .b64 db;\n\t"
      "add.u32 ta, %1, %5;\n\t"
      "add.u32 tb, %2, %6;\n\t"
      "mov.b64 db, {tb, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %4, %7;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(bhi), "r"(idesc), "n"(AOFF), "n"(BOFF), "n"(ACC ? 1 : 0)
      : "memory");
}

// ====================================================================================================================
// One CTA per SM, 20 warps, both TMEM halves in flight, softmax warps that never wait.
//   warps 0-7   softmax: TWO threads per query row (warp w: TMEM lane quarter w & 3, key half w >> 2).  A thread pulls its 128
//               scores out of TMEM ONCE, keeps them in registers for both the max and the exp pass; its key half is a softmax unit of
//               its own (own maximum and sum, merged by the epilogue).  All eight warps work on the same tile: while tile n is in
//               softmax, S(n+1) = Q K^T is already sitting in the other TMEM half, and P(n-1) V, its epilogue and the loads of the
//               next item run under it.
//   warps 8-11  epilogue: O out of TMEM, + p_256 v_256, / row sum, bf16, staged (128B-swizzled) in the dead Q tile, one TMA
//               store per warp.
//   warps 12-15 helpers: the key-256 score of all 256 rows and the whole of query row 256 (mma.sync matrix-vector products), one
//               item ahead of the softmax warps.
//   warp 16 TMA producer (two 102 KB stages), warp 17 MMA issuer, warps 18-19 register donors (setmaxnreg: 168 / 72 / 40 / 32: exactly the 96 x 640 registers
//   the CTA is launched with -- the pool setmaxnreg redistributes is the launch allocation, not the 64 K file: asking for more hangs).
// TMEM half s (256 columns): S fp32 [0,256) -> P bf16 of keys 0..127 in [0,64), of keys 128..255 in [128,192); the two key halves are
// INDEPENDENT softmax units (own row maximum, own row sum): O0 = P0 V[0:128] accumulates in [64,128), O1 = P1 V[128:256] in [192,256), and
// the epilogue merges them, (O0 w0 + (O1 + p_256 v_256) w1) / (l0 w0 + l1 w1) with w = exp(m_half - max(m0, m1)).  Nothing couples the
// two softmax groups (warps 0-3 / 4-7) inside a tile: no row-maximum exchange, no 64-thread barrier, and P0 V0 is issued while the second
// group is still in its exponential pass (softmax phase of a tile 3.3 k -> 2.45 k cycles; the tile period is now set by the hand-off of
// the 102 KB input stage at item boundaries and by the O read-out -> Q K^T -> S chain of the two TMEM halves).
// Work is split by TILE: CTA c owns tiles [c T / grid, (c+1) T / grid) (tile = item * 2 + half), so SMs differ by at most one
// tile; an item cut by a range boundary is staged by both neighbours.
// ====================================================================================================================
constexpr int NTHREADS = 20 * 32;
constexpr int W_EPI = 8, W_HELP = 12, W_TMA = 16, W_MMA = 17;
constexpr int OFF_BAR = 2 * STAGE_BYTES;
constexpr int OFF_SX = OFF_BAR + 256;               // float [2 stages][256]: score of every query row against key 256
constexpr int OFF_MX = OFF_SX + 2 * 256 * 4;        // float [2 slots][2 halves][128]: partial row maxima
constexpr int OFF_SUM = OFF_MX + 2 * 2 * 128 * 4;   // float [2 slots][2 halves][128]: partial row sums
constexpr int OFF_PX = OFF_SUM + 2 * 2 * 128 * 4;   // float [2 slots][128]: un-normalised probability of key 256
constexpr int OFF_PS = OFF_PX + 2 * 128 * 4;        // float [264]: un-normalised probabilities of query row 256
constexpr int OFF_RED = OFF_PS + 264 * 4;           // float [8]
constexpr int OFF_OP = OFF_RED + 32;                // float [4][64]: per-warp partial outputs of query row 256
constexpr int OFF_VF = OFF_OP + 4 * 64 * 4;         // float [2 stages][64]: value row of key 256 in fp32
constexpr int OFF_PB = OFF_VF + 2 * 64 * 4;         // bf16 [256]: probabilities of query row 256 as the A operand of its P V
constexpr int SMEM_BYTES = 1024 + OFF_PB + 256 * 2;
static_assert(SMEM_BYTES <= 232448, "shared memory");

// TMEM store without a compiler memory barrier: the exp pipeline below is scheduled across these stores
__device__ __forceinline__ void tmem_st8_nc(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void helper_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// O_half = P_half V_half: eight K16 steps over keys [16 K0, 16 K0 + 128)
template <int K, int K0>
__device__ __forceinline__ void pv_chain(uint32_t tmem_d, uint32_t tmem_a, uint32_t vlo, uint32_t vhi, uint32_t idesc) {
  if constexpr (K < K0 + 8) {
    umma_ts2<K != K0, 8 * K + (K >= 8 ? 64 : 0), K * (2048 >> 4)>(tmem_d, tmem_a, vlo, vhi, idesc);
    pv_chain<K + 1, K0>(tmem_d, tmem_a, vlo, vhi, idesc);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTail, const __grid_constant__ CUtensorMap tmOut,
               int n_tiles, long long* __restrict__ tstamp, bf16* __restrict__ out, int rev) {
  pdl_trigger();
  long long ts_c0 = 0, ts_g0 = 0;
  if (tstamp && threadIdx.x == 0) {
    ts_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_g0));
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + OFF_BAR;
  // [2] each: in_full, in_empty (per stage, part A = Q0 | K | Qtail); s_full, p_full, o_full, o_free (per TMEM half); sx_full (per stage);
  // p_full1 (key half 1); in_fullB, in_emptyB (per stage, part B = Q1 | V).  The stage is handed back in two parts: K and Q0 -- all that
  // S of the next item's first tile needs -- are free as soon as both Q K^T of the item are done and tile 0's output has left its
  // staging area, a whole softmax phase before Q1 and V are.
  const uint32_t in_full = bars, in_empty = bars + 16, s_full = bars + 32, p_full = bars + 48, o_full = bars + 64, o_free = bars + 80;
  const uint32_t sx_full = bars + 96, tmem_slot = bars + 112, p_full1 = bars + 128;      // p_full: key half 0, p_full1: key half 1
  const uint32_t in_fullB = bars + 144, in_emptyB = bars + 160;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(sm + OFF_BAR + 112);
  float* sxs = reinterpret_cast<float*>(sm + OFF_SX);
  float* mxs = reinterpret_cast<float*>(sm + OFF_MX);
  float* sums = reinterpret_cast<float*>(sm + OFF_SUM);
  float* pxs = reinterpret_cast<float*>(sm + OFF_PX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int begin = (int)((long long)blockIdx.x * n_tiles / gridDim.x), end = (int)((long long)(blockIdx.x + 1) * n_tiles / gridDim.x);
  const int item0 = begin >> 1;
  const int last_item = (n_tiles >> 1) - 1;      // rev: the (image, head) items are walked from the last to the first (gemm_tc.cuh, "serpentine tile order")

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTail);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full + 8 * s, 1);
      mbar_init(in_empty + 8 * s, 9);     // MMA commit + 4 helper warps + 4 epilogue warps
      mbar_init(in_fullB + 8 * s, 1);
      mbar_init(in_emptyB + 8 * s, 9);
      mbar_init(s_full + 8 * s, 1);
      mbar_init(p_full + 8 * s, 4);
      mbar_init(p_full1 + 8 * s, 4);
      mbar_init(o_full + 8 * s, 1);
      mbar_init(o_free + 8 * s, 4);
      mbar_init(sx_full + 8 * s, 4);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp >= W_TMA) {
    reg_dec<32>();
    if (warp == W_TMA && lane == 0) {
      // ============================ TMA producer ============================
      const int item1 = (end + 1) >> 1;
      for (int item = item0; item < item1; ++item) {
        const int il = item - item0, s = il & 1;
        const int ir = rev ? last_item - item : item;
        const int b = ir / DH, h = ir % DH;
        const uint32_t st = smem_base + s * STAGE_BYTES;
        const uint32_t par = ((il >> 1) & 1) ^ 1;
        mbar_wait(in_empty + 8 * s, par);                            // part A: Q0 | K (272 rows) | Qtail
        mbar_expect_tx(in_full + 8 * s, TILE_BYTES + KV_BYTES + 16 * 128);
        tma_load_2d(st, &tmQKV, in_full + 8 * s, h * DHD, b * S_);
        tma_load_2d(st + OFF_K, &tmQKV, in_full + 8 * s, DD + h * DHD, b * S_);
        tma_load_2d(st + OFF_K + TILE_BYTES, &tmQKV, in_full + 8 * s, DD + h * DHD, b * S_ + 128);
        tma_load_2d(st + OFF_K + 2 * TILE_BYTES, &tmTail, in_full + 8 * s, DD + h * DHD, b * S_ + 256);
        tma_load_2d(st + OFF_QT, &tmTail, in_full + 8 * s, h * DHD, b * S_ + 256);
        mbar_wait(in_emptyB + 8 * s, par);                           // part B: Q1 | V (272 rows)
        mbar_expect_tx(in_fullB + 8 * s, TILE_BYTES + KV_BYTES);
        tma_load_2d(st + TILE_BYTES, &tmQKV, in_fullB + 8 * s, h * DHD, b * S_ + 128);
        tma_load_2d(st + OFF_V, &tmQKV, in_fullB + 8 * s, 2 * DD + h * DHD, b * S_);
        tma_load_2d(st + OFF_V + TILE_BYTES, &tmQKV, in_fullB + 8 * s, 2 * DD + h * DHD, b * S_ + 128);
        tma_load_2d(st + OFF_V + 2 * TILE_BYTES, &tmTail, in_fullB + 8 * s, 2 * DD + h * DHD, b * S_ + 256);
      }
    } else if (warp == W_MMA && lane == 0) {
      // ============================ MMA issuer ============================
      constexpr uint32_t idesc_s = make_idesc(128, 256);
      constexpr uint32_t idesc_o = make_idesc_bmn(128, 64);
      const uint64_t dk = make_smem_desc(smem_base + OFF_K);
      const uint64_t dv = make_smem_desc_mn(smem_base + OFF_V);
      const uint32_t klo = (uint32_t)dk, khi = (uint32_t)(dk >> 32), vlo = (uint32_t)dv, vhi = (uint32_t)(dv >> 32);
      // S(g) = Q K^T of tile g into TMEM half (g - begin) & 1, as soon as that half's previous O has been read out
      auto issue_qk = [&](int g) {
        const int n = g - begin, slot = n & 1, t = g & 1, il = (g >> 1) - item0, s = il & 1;
        if (t == 0 || g == begin) mbar_wait(in_full + 8 * s, (il >> 1) & 1);
        if (t == 1) mbar_wait(in_fullB + 8 * s, (il >> 1) & 1);        // Q1
        mbar_wait(o_free + 8 * slot, ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t ks = klo + s * (STAGE_BYTES >> 4);
        const uint32_t qs = ks - ((OFF_K - t * TILE_BYTES) >> 4);     // same descriptor fields, Q tile t's address
        const uint32_t d = tmem_base + 256 * slot;
        umma_ss2<false, 0>(d, qs, ks, khi, idesc_s);
        umma_ss2<true, 2>(d, qs, ks, khi, idesc_s);
        umma_ss2<true, 4>(d, qs, ks, khi, idesc_s);
        umma_ss2<true, 6>(d, qs, ks, khi, idesc_s);
        umma_commit(s_full + 8 * slot);
        if (t == 1 || g == end - 1) umma_commit(in_empty + 8 * s);     // the last Q K^T of the item here: K and Q0 are free when it retires
      };
      if (begin < end) issue_qk(begin);
      for (int g = begin; g < end; ++g) {
        if (g + 1 < end) issue_qk(g + 1);
        const int n = g - begin, slot = n & 1, t = g & 1, il = (g >> 1) - item0, s = il & 1;
        const uint32_t d = tmem_base + 256 * slot;
        mbar_wait(in_fullB + 8 * s, (il >> 1) & 1);                    // V
        mbar_wait(p_full + 8 * slot, (n >> 1) & 1);
        tc_fence_after();
        pv_chain<0, 0>(d + 64, d, vlo + s * (STAGE_BYTES >> 4), vhi, idesc_o);
        mbar_wait(p_full1 + 8 * slot, (n >> 1) & 1);
        tc_fence_after();
        pv_chain<8, 8>(d + 192, d, vlo + s * (STAGE_BYTES >> 4), vhi, idesc_o);
        umma_commit(o_full + 8 * slot);
        if (t == 1 || g == end - 1) umma_commit(in_emptyB + 8 * s);   // every MMA reading this stage has been issued
      }
    }
  } else if (warp >= W_HELP) {
    // ============================ helpers: key-256 column for all rows, and query row 256 ============================
    reg_dec<40>();
    // Warp-level mma.sync (m16n8k16, bf16 -> fp32) with the vector in column / row 0 of the B / A fragment: a matrix-vector product
    // costs 1 ldmatrix + 1 mma per 16 x 16 block instead of ~500 CUDA-core instructions, which matters because this kernel is bound by
    // instruction issue, not by any pipe.
    const int hw = warp - W_HELP, tid = hw * 32 + lane, g = lane >> 2, tq = lane & 3, mi = lane >> 3, rr = lane & 7;
    float* ps = reinterpret_cast<float*>(sm + OFF_PS);
    float* red = reinterpret_cast<float*>(sm + OFF_RED);
    float* op = reinterpret_cast<float*>(sm + OFF_OP);
    const uint32_t* pbw = reinterpret_cast<const uint32_t*>(sm + OFF_PB);
    const int item1 = (end + 1) >> 1;
    for (int item = item0; item < item1; ++item) {
      const int il = item - item0, s = il & 1;
      const bool own256 = 2 * item + 1 >= begin && 2 * item + 1 < end;     // the CTA holding tile 1 also produces query row 256
      const int ir = rev ? last_item - item : item;
      const int b = ir / DH, h = ir % DH;
      const uint8_t* st = sm + s * STAGE_BYTES;
      const uint32_t st_u = smem_base + s * STAGE_BYTES;
      mbar_wait(in_full + 8 * s, (il >> 1) & 1);
      mbar_wait(in_fullB + 8 * s, (il >> 1) & 1);
      // ---- key-256 column: rows [64 hw, 64 hw + 64) of Q times k_256 (B fragment column 0 = lanes 0..3) ----
      {
        uint32_t vb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) vb[i] = lane < 4 ? *reinterpret_cast<const uint32_t*>(st + OFF_K + 256 * 128 + 4 * (4 * i + tq)) : 0u;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int row = 64 * hw + 16 * mt + (mi & 1) * 8 + rr;          // this lane's ldmatrix row (row & 7 == rr)
          const uint32_t base = st_u + (row >> 7) * TILE_BYTES + (row & 127) * 128;
          float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            attn::ldsm_x4(base + (((2 * ks + (mi >> 1)) ^ rr) << 4), af[0], af[1], af[2], af[3]);
            attn::mma_bf16(c, af, vb[2 * ks], vb[2 * ks + 1]);
          }
          if (tq == 0) {
            sxs[s * 256 + 64 * hw + 16 * mt + g] = c[0];
            sxs[s * 256 + 64 * hw + 16 * mt + g + 8] = c[2];
          }
        }
      }
      if (hw == 0) {                                       // value row of key 256 in fp32 for the epilogue warps
        const uint32_t v2 = *reinterpret_cast<const uint32_t*>(st + OFF_V + 256 * 128 + lane * 4);
        *reinterpret_cast<float2*>(sm + OFF_VF + s * 256 + lane * 8) = make_float2(bf_lo(v2), bf_hi(v2));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sx_full + 8 * s);
      if (own256) {
        // ---- scores of query 256: keys [64 hw, 64 hw + 64) of K (and key 256 on warp 0) times q_256 ----
        uint32_t vb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) vb[i] = lane < 4 ? *reinterpret_cast<const uint32_t*>(st + OFF_QT + 4 * (4 * i + tq)) : 0u;
#pragma unroll
        for (int mt = 0; mt < 5; ++mt) {
          if (mt == 4 && hw != 0) break;
          const int row = (mt < 4 ? 64 * hw + 16 * mt : 256) + (mi & 1) * 8 + rr;
          const uint32_t base = st_u + OFF_K + row * 128;
          float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            attn::ldsm_x4(base + (((2 * ks + (mi >> 1)) ^ rr) << 4), af[0], af[1], af[2], af[3]);
            attn::mma_bf16(c, af, vb[2 * ks], vb[2 * ks + 1]);
          }
          if (mt < 4) {
            if (tq == 0) {
              ps[64 * hw + 16 * mt + g] = c[0];
              ps[64 * hw + 16 * mt + g + 8] = c[2];
            }
          } else if (lane == 0) {
            ps[256] = c[0];
          }
        }
        helper_sync();
        const float b0 = ps[tid], b1 = ps[128 + tid], a2 = ps[256];
        float m = fmaxf(fmaxf(b0, b1), a2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[hw] = m;
        helper_sync();
        m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        const float p0 = ex2a((b0 - m) * LOG2E), p1 = ex2a((b1 - m) * LOG2E), p2 = ex2a((a2 - m) * LOG2E);
        float l = p0 + p1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        reinterpret_cast<bf16*>(sm + OFF_PB)[tid] = __float2bfloat16_rn(p0);
        reinterpret_cast<bf16*>(sm + OFF_PB)[128 + tid] = __float2bfloat16_rn(p1);
        if (lane == 0) red[4 + hw] = l;
        helper_sync();
        l = red[4] + red[5] + red[6] + red[7] + p2;
        // ---- O_256 = p V over keys [64 hw, 64 hw + 64): p in row 0 of the A fragment (lanes 0..3), V through ldmatrix.trans ----
        uint32_t pa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pa[i] = lane < 4 ? pbw[8 * (4 * hw + (i >> 1)) + 4 * (i & 1) + tq] : 0u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float c[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 16 * (4 * hw + i) + (mi & 1) * 8 + rr;
            const uint32_t af[4] = {pa[2 * i], 0u, pa[2 * i + 1], 0u};
#pragma unroll
            for (int ntp = 0; ntp < 2; ++ntp) {
              uint32_t bf[4];
              attn::ldsm_x4_t(st_u + OFF_V + row * 128 + (((2 * (2 * half + ntp) + (mi >> 1)) ^ rr) << 4), bf[0], bf[1], bf[2], bf[3]);
              attn::mma_bf16(c[2 * ntp], af, bf[0], bf[1]);
              attn::mma_bf16(c[2 * ntp + 1], af, bf[2], bf[3]);
            }
          }
          if (lane < 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(op + hw * 64 + 8 * (4 * half + j) + 2 * tq) = make_float2(c[j][0], c[j][1]);
          }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(in_empty + 8 * s); mbar_arrive(in_emptyB + 8 * s); }        // this warp is done reading the stage
        helper_sync();
        if (hw == 0) {
          const float il2 = 1.0f / l;
          const float2 vx = *reinterpret_cast<const float2*>(sm + OFF_VF + s * 256 + lane * 8);
          const float r0 = (op[2 * lane] + op[64 + 2 * lane]) + (op[128 + 2 * lane] + op[192 + 2 * lane]) + p2 * vx.x;
          const float r1 = (op[2 * lane + 1] + op[64 + 2 * lane + 1]) + (op[128 + 2 * lane + 1] + op[192 + 2 * lane + 1]) + p2 * vx.y;
          *reinterpret_cast<uint32_t*>(out + ((int64_t)b * S_ + 256) * DD + h * DHD + 2 * lane) = pack_bf16(r0 * il2, r1 * il2);
        }
      } else {
        __syncwarp();
        if (lane == 0) { mbar_arrive(in_empty + 8 * s); mbar_arrive(in_emptyB + 8 * s); }
      }
    }
  } else if (warp >= W_EPI) {
    // ============================ epilogue (TMEM lane quarter = warp & 3) ============================
    reg_dec<72>();
    const int q = warp & 3, rl = q * 32 + lane;
    for (int g = begin; g < end; ++g) {
      const int n = g - begin, slot = n & 1, t = g & 1, item = g >> 1, il = item - item0, s = il & 1;
      const uint32_t ph = (n >> 1) & 1;
      const int ir = rev ? last_item - item : item;
      const int b = ir / DH, h = ir % DH;
      uint8_t* st = sm + s * STAGE_BYTES;
      const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + 256 * slot;
      if (t == 0 || g == begin) mbar_wait(sx_full + 8 * s, (il >> 1) & 1);     // the helpers' fp32 copy of V row 256 (and, through them, the stage)
      mbar_wait(p_full + 8 * slot, ph);                                         // ... and both softmax groups' row statistics
      mbar_wait(p_full1 + 8 * slot, ph);
      // merge of the two independent key halves: weights exp(m_half - m), common denominator
      const float m0 = mxs[(slot * 2 + 0) * 128 + rl], m1 = mxs[(slot * 2 + 1) * 128 + rl];
      const float mm = fmaxf(m0, m1);
      const float e0 = ex2a((m0 - mm) * LOG2E), e1 = ex2a((m1 - mm) * LOG2E);
      const float inv = 1.0f / fmaf(sums[(slot * 2 + 0) * 128 + rl], e0, sums[(slot * 2 + 1) * 128 + rl] * e1);
      const float w0 = e0 * inv, w1 = e1 * inv, px = pxs[slot * 128 + rl] * w1;
      const float2 w02 = make_float2(w0, w0), w12 = make_float2(w1, w1), px2 = make_float2(px, px);
      mbar_wait(o_full + 8 * slot, ph);
      tc_fence_after();
      uint8_t* stg = st + t * TILE_BYTES + q * 4096 + lane * 128;     // this warp's 32 x 128 B slice of the dead Q tile
      const float4* vf = reinterpret_cast<const float4*>(sm + OFF_VF + s * 256);
#pragma unroll
      for (int c = 0; c < 4; ++c) {                           // 16 output columns at a time (72 registers): O0 in [64,128), O1 in [192,256)
        uint32_t r0[16], r1[16];
        tmem_ld16(tm + 64 + 16 * c, r0);
        tmem_ld16(tm + 192 + 16 * c, r1);
        tmem_wait_ld();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free + 8 * slot);      // O is in registers now: the next S may overwrite this TMEM half
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {                         // O0 w0 + O1 w1 + p_256 v_256 (w = weight / denominator) on the packed fp32x2 pipe
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4 v4 = vf[c * 4 + i * 2 + e];
            const int k = i * 8 + 4 * e;
            const float2 lo = __ffma2_rn(w02, make_float2(__uint_as_float(r0[k]), __uint_as_float(r0[k + 1])),
                                         __ffma2_rn(w12, make_float2(__uint_as_float(r1[k]), __uint_as_float(r1[k + 1])), __fmul2_rn(px2, make_float2(v4.x, v4.y))));
            const float2 hi = __ffma2_rn(w02, make_float2(__uint_as_float(r0[k + 2]), __uint_as_float(r0[k + 3])),
                                         __ffma2_rn(w12, make_float2(__uint_as_float(r1[k + 2]), __uint_as_float(r1[k + 3])), __fmul2_rn(px2, make_float2(v4.z, v4.w))));
            w[2 * e] = pack_bf16(lo.x, lo.y);
            w[2 * e + 1] = pack_bf16(hi.x, hi.y);
          }
          *reinterpret_cast<uint4*>(stg + (((c * 2 + i) ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmOut, smem_base + s * STAGE_BYTES + t * TILE_BYTES + q * 4096, h * DHD, b * S_ + t * 128 + q * 32);
        bulk_commit();
        const bool first = t == 0 || g == begin, last = t == 1 || g == end - 1;   // of this item's tiles in this CTA
        if (first || last) {                                 // hand the stage back once TMA has read the staging: Q0 | K after the first tile, Q1 | V after the last
          bulk_wait_read0();
          if (first) mbar_arrive(in_empty + 8 * s);
          if (last) mbar_arrive(in_emptyB + 8 * s);
        }
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait0();
  } else {
    // ============================ softmax: 2 threads per row (lane quarter q, key half hf) ============================
    reg_inc<168>();
#define TS(k) do { if (tstamp && blockIdx.x == 0 && threadIdx.x == 0 && n < 6) tstamp[n * 8 + (k)] = clock64(); } while (0)
    const int q = warp & 3, hf = warp >> 2, rl = q * 32 + lane;
    for (int g = begin; g < end; ++g) {
      const int n = g - begin, slot = n & 1, t = g & 1, il = (g >> 1) - item0, s = il & 1;
      const uint32_t ph = (n >> 1) & 1;
      const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + 256 * slot + 128 * hf;
      TS(0);
      mbar_wait(s_full + 8 * slot, ph);
      tc_fence_after();
      TS(1);
      uint32_t r[4][32];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld32(tm + 32 * j, r[j]);
      tmem_wait_ld();
      // four independent chains (the compiler emits 3-input FMNMX3): one chain of 64 dependent FMNMX3 is ~300 cycles of pure latency
      float mq[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 32; i += 2) mq[(i >> 1) & 3] = fmaxf(mq[(i >> 1) & 3], fmaxf(__uint_as_float(r[j][i]), __uint_as_float(r[j][i + 1])));
      float mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
      float sx = 0.f;
      if (hf == 1) {
        if (t == 0 || g == begin) mbar_wait(sx_full + 8 * s, (il >> 1) & 1);     // the helpers' key-256 column of this item
        sx = sxs[s * 256 + t * 128 + rl];
        mx = fmaxf(mx, sx);
      }
      mxs[(slot * 2 + hf) * 128 + rl] = mx;                   // this key half's own maximum: the epilogue merges the halves
      const float nm = -mx * LOG2E;
      TS(2);
      // P = exp2((s - max) log2e) as bf16 into this thread's own (already consumed) half: columns [128 hf, 128 hf + 64).
      // The kernel is instruction-issue bound (four roles share every scheduler), so the scale/shift and the row sum use the packed
      // fp32x2 pipe: per key 0.5 FFMA2 + 1 MUFU + 0.5 FADD2 + 0.5 F2FP.  Blocks of 16 keys are software-pipelined: the MUFU ops of
      // block b + 1 are issued before the adds / packs / TMEM store of block b.
      const float2 l2 = make_float2(LOG2E, LOG2E), nm2 = make_float2(nm, nm);
      float2 acc = make_float2(0.f, 0.f);
      float2 e[2][8];
      auto exp_block = [&](int blk, float2 (&o)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[blk >> 1][16 * (blk & 1) + 2 * i]), __uint_as_float(r[blk >> 1][16 * (blk & 1) + 2 * i + 1])), l2, nm2);
          if (i < NPOLY) o[i] = ex2_poly2(x);                   // this pair on the FMA pipe ...
          else o[i] = make_float2(ex2a(x.x), ex2a(x.y));        // ... the others on the MUFU pipe
        }
      };
      exp_block(0, e[0]);
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {
        if (blk + 1 < 8) exp_block(blk + 1, e[(blk + 1) & 1]);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc = __fadd2_rn(acc, e[blk & 1][i]);
          pk[i] = pack_bf16(e[blk & 1][i].x, e[blk & 1][i].y);
        }
        tmem_st8_nc(tm + 8 * blk, pk);
      }
      float sum = acc.x + acc.y;
      if (hf == 1) {
        const float px = ex2a(fmaf(sx, LOG2E, nm));
        sum += px;
        pxs[slot * 128 + rl] = px;
      }
      sums[(slot * 2 + hf) * 128 + rl] = sum;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive((hf ? p_full1 : p_full) + 8 * slot);
      TS(3);
    }
#undef TS10
  }
  tc_fence_before();
  __syncthreads();
  if (tstamp && threadIdx.x == 0) {
    long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    tstamp[48 + 3 * blockIdx.x] = clock64() - ts_c0;
    tstamp[48 + 3 * blockIdx.x + 1] = ts_g0;
    tstamp[48 + 3 * blockIdx.x + 2] = g1;
  }
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline int dino_attention_tc(cudaStream_t st, const bf16* qkv, bf16* out, int B, int rev = 0) {
  static std::atomic<uint64_t> attr{0};   // per-device one-time setup
  if (device_once(attr)) {
    HVLA_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  CUtensorMap map, tail, omap;
  HVLA_TRY(make_map_bf16(&map, qkv, (int64_t)B * S_, 3 * DD, 128));
  HVLA_TRY(make_map_bf16(&tail, qkv, (int64_t)B * S_, 3 * DD, 16));
  HVLA_TRY(make_map_bf16(&omap, out, (int64_t)B * S_, DD, 32));
  const int n_tiles = 2 * B * DH;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  ProfScope ps(st, "dino_attention");
  static long long* d_ts = nullptr;
  const bool ts = getenv("HVLA_ATTN_TS") != nullptr;            // experiments: phase timestamps of CTA 0, cycle counts of every CTA
  if (ts && !d_ts) { cudaMalloc(&d_ts, 1024 * sizeof(long long)); cudaMemset(d_ts, 0, 1024 * sizeof(long long)); }
  launch_k(attn_tc_kernel, dim3(grid), dim3(NTHREADS), (size_t)SMEM_BYTES, st, map, tail, omap, n_tiles, ts ? d_ts : (long long*)nullptr, out, rev);
  if (ts) {
    long long h[1024];
    cudaMemcpy(h, d_ts, sizeof h, cudaMemcpyDeviceToHost);
    long long cmin = 1LL << 60, cmax = 0, g0 = 1LL << 62, g1 = 0, csum = 0;
    for (int c = 0; c < grid; ++c) {
      const long long cy = h[48 + 3 * c];
      cmin = cy < cmin ? cy : cmin; cmax = cy > cmax ? cy : cmax; csum += cy;
      g0 = h[48 + 3 * c + 1] < g0 ? h[48 + 3 * c + 1] : g0;
      g1 = h[48 + 3 * c + 2] > g1 ? h[48 + 3 * c + 2] : g1;
    }
    fprintf(stderr, "attn_tc CTAs %d: cycles min %lld avg %lld max %lld | wall %lld ns first start -> last end (%.0f MHz)\n", grid, cmin, csum / grid, cmax,
            g1 - g0, 1e3 * cmax / (double)(g1 - g0));
    for (int n = 0; n < 6; ++n) {
      const long long* t = h + n * 8;
      fprintf(stderr, "attn_tc tile %d: start %lld | s_full +%lld max +%lld exp +%lld\n", n, t[0] - h[0], t[1] - t[0], t[2] - t[1], t[3] - t[2]);
    }
  }
  HVLA_LAUNCH_CHECK("attn_tc");
  return HVLA_OK;
}

}  // namespace attn_tc
}  // namespace hvla
