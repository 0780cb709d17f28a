// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi(A[M,K] * Wt[N,K]^T)
//   A, Wt bf16 K-major; fp32 accumulation in tensor memory.
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..9 = epilogue (TMEM -> registers -> global).  128x256x64 tiles, 4-stage smem ring,
// two 256-column TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// Used for the DINOv2 patch-embed / QKV / out-proj / fc1 / fc2 GEMMs (99% of the FLOPs of one
// control step).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace hvla {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ + 4096 /*epilogue bias/ls*/ +
                           1024 + 8 * 2048 /*TMA-store slabs*/;

enum Epi {
  EPI_BIAS_BF16 = 0, EPI_BIAS_GELU_BF16 = 1, EPI_RESIDUAL_F32 = 2, EPI_PATCH_F32 = 3,
  // LayerNorm-free flow over the BLOCKED fp32 stream (see "stream LayerNorm without a LayerNorm kernel"):
  EPI_BIAS_BF16_FOLD = 4, EPI_BIAS_GELU_BF16_FOLD = 5,   // consumers: A = un-normalised bf16 shadow, row statistics applied in the epilogue
  EPI_RESIDUAL_BLK = 6,                                  // producer: x += ls*(acc+bias) in place + bf16 shadow + row statistics
  EPI_PATCH_BLK = 7,                                     // patch embedding: x += acc+bias at the shifted stream rows (no shadow)
  // split-operand ("bf16x3") flow, fp32-class accuracy on the tensor cores (dino_x3.cuh): the fp32 result v = acc + bias
  // (EPI_SPLIT_GELU_BF16: exact erf-GELU of it) leaves as TWO bf16 numbers, hi = bf16(v) and lo = bf16(v - hi), written as the
  // planes [hi | lo] or [hi | lo | hi] of the next GEMM's K-concatenated A operand (plane p at column p * plane_stride)
  EPI_SPLIT_BF16 = 8, EPI_SPLIT_GELU_BF16 = 9
};
constexpr bool epi_split(int e) { return e == EPI_SPLIT_BF16 || e == EPI_SPLIT_GELU_BF16; }
constexpr bool epi_fold(int e) { return e == EPI_BIAS_BF16_FOLD || e == EPI_BIAS_GELU_BF16_FOLD; }
constexpr bool epi_gelu(int e) { return e == EPI_BIAS_GELU_BF16 || e == EPI_BIAS_GELU_BF16_FOLD; }
constexpr bool epi_bf16_out(int e) { return e == EPI_BIAS_BF16 || e == EPI_BIAS_GELU_BF16 || epi_fold(e) || epi_split(e); }
constexpr bool epi_blk(int e) { return e == EPI_RESIDUAL_BLK || e == EPI_PATCH_BLK; }

struct EpiP {
  const float* bias;   // [N]
  void* out;           // bf16 [M,ldo] (EPI 0/1) | fp32 residual stream X (EPI 2/3)
  int ldo;
  const float* ls;     // [N] layer scale (EPI 2)
  const float* pos;    // [257,768] position table (EPI 3); EPI 7: the blocked table of the 256 patch rows (DvecLayout::pos_blk)
  float qscale;        // EPI 0: columns < qcols are multiplied by qscale (query pre-scaling)
  int qcols;
  int debug;           // experiment knob (hvla_gemm_bf16 only): 1 = handshake only, 2 = TMEM loads only, 4 = no global stores
  float* part;         // EPI 2: scratch for split-K partial products (null: never split), see "split-K partial products"
  size_t part_bytes;   //        its size
  int* splits_used;    //        host out: number of K splits of this launch (1 = none)
  int patch_rows;      // EPI 2 used for the patch embedding: GEMM row b*256+p lands in stream row b*257+1+p; ls == null means 1
  // ---- LayerNorm fused away (see "stream LayerNorm without a LayerNorm kernel" below) ----
  int rows;            // valid rows: M (EPI 4/5/7: GEMM rows; EPI 6: stream rows = GEMM rows)
  const float* stats;  // EPI 4/5 consumer: [M][6][2] partial (sum, sum of squares) of the A rows; A is the UN-normalised bf16 shadow
  const float* cs;     //                   [N] column sums of the (gamma-folded, bf16-rounded) weight; bias = folded bias
  bf16* shadow;        // EPI 6 producer: bf16 copy of the updated stream [M,768]; `out` is the BLOCKED fp32 stream
  float* stats_out;    //                 [M][6][2] partial row statistics of the updated stream
  int plane_stride;    // EPI 8/9: columns between the hi / lo / hi planes of the output row (the next GEMM's K)
  int nplanes;         //          2 = [hi | lo], 3 = [hi | lo | hi]
  // ---- L2 locality between consecutive kernels (see "serpentine tile order" below) ----
  int rev;             // 1: the m-tiles are walked from the last row block to the first
  int xhint;           // EPI 6/7: 1 = the blocked fp32 stream is read and written with an L2 evict_last policy
};

// ---- serpentine tile order -----------------------------------------------------------------------------------------------
// Every activation of a DINOv2 layer (25 - 101 MB at 64 images) is written by one kernel and read by the next.  If both walk the
// rows in the same direction the consumer starts with the OLDEST rows of a stream that is as large as the 126 MB L2 -- the rows
// that have just been evicted -- and its own writes keep evicting the rows it is about to read: every read is an HBM read.
// Consecutive kernels therefore walk the row blocks in OPPOSITE directions (ep.rev alternates along the launch chain, the
// attention kernel takes the same flag): the consumer begins with the rows the producer wrote last, which are still in L2.
// The fp32 residual stream is the one long-lived tensor (read by proj and fc2, two kernels apart): its accesses carry an
// evict_last policy (ep.xhint) when it is small enough to stay resident beside the streaming operands.
__device__ __forceinline__ uint64_t l2_policy(bool evict_last) {
  uint64_t p;
  if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldcg_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void stcg_hint(float4* ptr, const float4& v, uint64_t pol) {
  asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// Blocked fp32 residual stream: element (row, col) of the logical [M,768] stream lives at float index
//   ((row >> 5) * 192 + (col >> 2)) * 128 + (row & 31) * 4 + (col & 3)
// i.e. 32-row blocks, inside a block one 512-byte run per group of 4 columns.  A tcgen05 epilogue thread owns one ROW
// (TMEM lane) and a warp 32 consecutive rows, so the warp's 16-byte accesses to one column group are 512 contiguous bytes:
// the residual epilogues read and write the stream in place, fully coalesced, without shared memory or TMA.
__host__ __device__ __forceinline__ int64_t xblk_f4(int row, int col4) {      // float4 index of (row, columns 4*col4 .. 4*col4+3)
  return ((int64_t)(row >> 5) * (DD / 4) + col4) * 32 + (row & 31);
}

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- stream LayerNorm without a LayerNorm kernel ------------------------------------------------------------------------
// The 24 LayerNorms between the residual GEMMs and the q|k|v / fc1 GEMMs cost 13 % of a step as a separate HBM pass
// (50 MB fp32 in, 25 MB bf16 out each).  In the large-batch flow (hvla_capi.cu: dino_bf16, "flow B") they do not exist:
//  * producer (proj / fc2, EPI_RESIDUAL_BLK): x_new = x_old + ls*(acc+bias) is formed in registers; x_old is read from and x_new
//    written to the BLOCKED stream in place (coalesced 16-byte accesses, see xblk_f4), a bf16 shadow of x_new goes out through
//    the per-warp slab + TMA store, and each thread -- it owns one row and 128 columns of the tile -- writes its partial
//    (sum, sum of squares): 6 partials per row (3 n-tiles x 2 halves);
//  * consumer (q|k|v / fc1, EPI_*_FOLD): A is the un-normalised shadow, gamma/beta are folded into W and the bias
//    (params.py), and the row statistics enter after the matrix product:  rstd*(x W' - mean*colsum(W')) + b_f.
// Round 1 had the producer read the old values from a ROW-MAJOR stream (32 different lines per warp instruction) and store
// the new ones through fp32 slabs: proj 27.7 -> 71.8 us.  The blocked layout is what makes the in-place update cheap.
// Small batches (split-K) keep the classic flow: row-major stream, TMA reduce-add, LayerNorm kernels.
//
// ---- split-K partial products -------------------------------------------------------------------------------
// At small batch the residual GEMMs (N = 768) have only a handful of output tiles; their K range is then split
// over several CTA pairs.  Split 0 reduce-adds ls*(acc+bias) into the fp32 residual stream as usual; split s > 0
// writes its ls*acc to block s-1 of a scratch buffer [splits-1][M_pad][N] with a plain TMA store, and the LayerNorm
// that always follows adds the blocks to the stream in a fixed order (simt_kernels.cuh: layernorm768_kernel).
// Deterministic, no inter-CTA waiting.
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps only

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}


// ---- epilogue of one 128x256 accumulator (per CTA), shared by the 1-CTA and 2-CTA kernels --------------
// Called by the 8 epilogue warps.  Bias / LayerScale of the tile's 256 columns are staged in shared
// memory while the MMAs are still running, TMEM loads are double-buffered against the math/stores,
// and (residual epilogue) the fp32 residual values are prefetched one chunk ahead.
constexpr int EPI_SMEM_FLOATS = 2 * 2 * 256;   // [accumulator stage][bias | ls][256]

template <int EPI>
__device__ __forceinline__ void epilogue_tile_direct(const EpiP& ep, float* sepi, uint32_t tfull_bar_addr, uint32_t aph, int as,
                                              uint32_t tmem_base, int m0, int n0, int M, int warp, int lane) {
  const int ew = warp - 2;
  const int quarter = warp & 3;          // TMEM lane quarter this warp may access
  const int half = ew >> 2;              // which 128-column half of the tile
  const int te = threadIdx.x - 64;       // 0..255
  float* sb = sepi + as * 512;
  float* sl = sb + 256;
  sb[te] = __ldg(ep.bias + n0 + te);
  if (EPI == EPI_RESIDUAL_F32) sl[te] = __ldg(ep.ls + n0 + te);
  const int row = m0 + quarter * 32 + lane;
  const bool row_ok = row < M;
  const int colh = n0 + half * 128;
  float4 xr[2][8];
  float* xrow = nullptr;
  if (EPI == EPI_RESIDUAL_F32) {
    xrow = reinterpret_cast<float*>(ep.out) + (int64_t)(row_ok ? row : 0) * ep.ldo + colh;
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[0][j] = *reinterpret_cast<const float4*>(xrow + 4 * j);
  }
  epi_bar_sync();                        // staged bias visible to all epilogue warps
  mbar_wait(tfull_bar_addr, aph);
  tc_fence_after();
  if (ep.debug & 1) { tc_fence_before(); __syncwarp(); return; }
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * 128);
  uint32_t r[2][32];
  tmem_ld32(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_wait_ld();
    if (c < 3) {
      tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
      if (EPI == EPI_RESIDUAL_F32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[(c + 1) & 1][j] = *reinterpret_cast<const float4*>(xrow + (c + 1) * 32 + 4 * j);
      }
    }
    const uint32_t(&rc)[32] = r[c & 1];
    if (ep.debug & 2) { if (rc[0] == 0x7fc12345u && rc[7] == 0x7fc54321u) reinterpret_cast<float*>(ep.out)[0] = 1.f; continue; }
    const int cl = half * 128 + c * 32;  // column inside the tile
    const int col = n0 + cl;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j);
      v[j] = __uint_as_float(rc[j]) + b4.x;
      v[j + 1] = __uint_as_float(rc[j + 1]) + b4.y;
      v[j + 2] = __uint_as_float(rc[j + 2]) + b4.z;
      v[j + 3] = __uint_as_float(rc[j + 3]) + b4.w;
    }
    if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
      if (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_tanhfit(v[j]);
      } else if (col < ep.qcols) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= ep.qscale;
      }
      if (row_ok && !((ep.debug & 4) && v[0] != 123.456f)) {
        bf16* o = reinterpret_cast<bf16*>(ep.out) + (int64_t)row * ep.ldo + col;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 q;
          q.x = pack_bf16(v[j], v[j + 1]);
          q.y = pack_bf16(v[j + 2], v[j + 3]);
          q.z = pack_bf16(v[j + 4], v[j + 5]);
          q.w = pack_bf16(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = q;
        }
      }
    } else if (EPI == EPI_RESIDUAL_F32) {
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(sl + cl + j);
          float4 x4 = xr[c & 1][j >> 2];
          x4.x = x4.x + v[j] * l4.x;
          x4.y = x4.y + v[j + 1] * l4.y;
          x4.z = x4.z + v[j + 2] * l4.z;
          x4.w = x4.w + v[j + 3] * l4.w;
          *reinterpret_cast<float4*>(xrow + c * 32 + j) = x4;
        }
      }
    } else {  // EPI_PATCH_F32: row = b*256+p -> token b*257+1+p, add position table
      if (row_ok) {
        const int b = row >> 8, pidx = row & 255;
        float* x = reinterpret_cast<float*>(ep.out) + ((int64_t)b * DTOK + 1 + pidx) * ep.ldo + col;
        const float* pz = ep.pos + (int64_t)(1 + pidx) * DD + col;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(pz + j));
          *reinterpret_cast<float4*>(x + j) = make_float4(v[j] + p4.x, v[j + 1] + p4.y, v[j + 2] + p4.z, v[j + 3] + p4.w);
        }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
}


// ---- epilogue through shared memory + TMA (EPI 0/1/2) -----------------------------------------------------
// Row-per-thread global stores touch 32 different lines per warp instruction and throttled the whole
// kernel (measured: QKV GEMM 1.03 PFLOP/s with them, 1.50 without).  Here every warp transposes 32x64-byte
// units through a 2 KB shared-memory slab (SWIZZLE_64B, conflict-free 16-byte stores) and one lane hands
// the slab to the TMA engine: a plain tensor store for bf16 outputs, a tensor REDUCE-ADD (fp32) for the
// residual stream, so x += ls * (acc + bias) needs no read of x through the SM at all.
constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 2048;
// called by every epilogue thread once the tile's accumulator is complete and before the tile's first global store
// (gemm_chain.cuh publishes the PREVIOUS tile there: its stores are a whole tile old by then, so the fences cost nothing)
struct NoHook { __device__ __forceinline__ void operator()() const {} };

template <int EPI, int NSLAB = 1, class Hook = NoHook>
__device__ __forceinline__ void epilogue_tile_tma(const EpiP& ep, const CUtensorMap* tmO, float* sepi, uint32_t sstage,
                                                  uint32_t tfull_bar_addr, uint32_t aph, int as, uint32_t tmem_base, int m0,
                                                  int n0, int warp, int lane, int split = 0, const CUtensorMap* tmP = nullptr,
                                                  int part_row0 = 0, const CUtensorMap* tmX = nullptr, Hook hook = Hook()) {
  const bool add_bias = split == 0;
  const int ew = warp - 2;
  const int quarter = warp & 3;
  const int half = ew >> 2;
  const int te = threadIdx.x - 64;
  float* sb = sepi + as * 512;
  float* sl = sb + 256;
  sb[te] = add_bias ? __ldg(ep.bias + n0 + te) : 0.f;      // split-K: only the first K split contributes the bias
  if (EPI == EPI_RESIDUAL_F32) sl[te] = ep.ls ? __ldg(ep.ls + n0 + te) : 1.0f;
  constexpr bool fold = epi_fold(EPI);
  if (fold) sl[te] = __ldg(ep.cs + n0 + te);
  const int row0 = m0 + quarter * 32 + (ep.patch_rows ? m0 / 256 + 1 : 0);
  const int myrow = row0 + lane;                              // the stream / A row this thread's TMEM lane holds
  float f_rstd = 1.f, f_nmr = 0.f;                            // consumer: rstd and -mean*rstd of this row
  if (fold && myrow < ep.rows) {
    const float4* sp = reinterpret_cast<const float4*>(ep.stats + (int64_t)myrow * 12);
    const float4 p0 = __ldcg(sp), p1 = __ldcg(sp + 1), p2 = __ldcg(sp + 2);   // L2, not the read-only path: a GEMM chain rewrites the statistics inside one launch
    const float sum = ((p0.x + p0.z) + (p1.x + p1.z)) + (p2.x + p2.z), sq = ((p0.y + p0.w) + (p1.y + p1.w)) + (p2.y + p2.w);
    const float mean = sum * (1.f / 768.f);
    f_rstd = 1.0f / sqrtf(fmaxf(0.f, sq * (1.f / 768.f) - mean * mean) + 1e-6f);
    f_nmr = -mean * f_rstd;
  }
  // NSLAB 2 KB slabs per warp: with two, the TMA store of one chunk reads its slab while the next chunk is written
  const uint32_t slab0 = sstage + (uint32_t)ew * (2048u * NSLAB);
  const uint32_t my0 = slab0 + (uint32_t)lane * 64u;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);          // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
  epi_bar_sync();
  mbar_wait(tfull_bar_addr, aph);
  tc_fence_after();
  hook();
  if (ep.debug & 1) { tc_fence_before(); __syncwarp(); return; }
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * 128);
  uint32_t r[2][32];
  tmem_ld32(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_wait_ld();
    if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
    const uint32_t(&rc)[32] = r[c & 1];
    const int cl = half * 128 + c * 32;
    const int col = n0 + cl;
    float v[32];
    if (fold) {                                               // rstd*acc + (-mean*rstd)*cs + b_f
      const float2 rs2 = make_float2(f_rstd, f_rstd), nm2 = make_float2(f_nmr, f_nmr);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j);
        const float4 c4 = *reinterpret_cast<const float4*>(sl + cl + j);
        const float2 lo = __ffma2_rn(rs2, make_float2(__uint_as_float(rc[j]), __uint_as_float(rc[j + 1])),
                                     __ffma2_rn(nm2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
        const float2 hi = __ffma2_rn(rs2, make_float2(__uint_as_float(rc[j + 2]), __uint_as_float(rc[j + 3])),
                                     __ffma2_rn(nm2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
        v[j] = lo.x; v[j + 1] = lo.y; v[j + 2] = hi.x; v[j + 3] = hi.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + j);
        const float2 lo = __fadd2_rn(make_float2(__uint_as_float(rc[j]), __uint_as_float(rc[j + 1])), make_float2(b4.x, b4.y));
        const float2 hi = __fadd2_rn(make_float2(__uint_as_float(rc[j + 2]), __uint_as_float(rc[j + 3])), make_float2(b4.z, b4.w));
        v[j] = lo.x; v[j + 1] = lo.y; v[j + 2] = hi.x; v[j + 3] = hi.y;
      }
    }
    if (epi_bf16_out(EPI)) {
      if (epi_gelu(EPI)) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 g = gelu_erf_tanhfit2(make_float2(v[j], v[j + 1]));
          v[j] = g.x; v[j + 1] = g.y;
        }
      } else if (!epi_split(EPI) && col < ep.qcols) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= ep.qscale;
      }
      if (EPI == EPI_SPLIT_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);         // exact erf-GELU: this flow is the fp32-class one
      }
      if (epi_split(EPI)) {                                         // hi plane(s), then the lo plane through the same slab
        float lo[32];
        uint32_t hp[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
          hp[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[j] = v[j] - __low2float(h2);
          lo[j + 1] = v[j + 1] - __high2float(h2);
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my0 + ((((uint32_t)j) ^ sw) << 4)), "r"(hp[4 * j]), "r"(hp[4 * j + 1]),
                       "r"(hp[4 * j + 2]), "r"(hp[4 * j + 3]) : "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmO, slab0, col, row0);
          if (ep.nplanes == 3) tma_store_2d(tmO, slab0, col + 2 * ep.plane_stride, row0);
          bulk_commit();
          bulk_wait_read<0>();
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my0 + ((((uint32_t)j) ^ sw) << 4)), "r"(pack_bf16(lo[8 * j], lo[8 * j + 1])),
                       "r"(pack_bf16(lo[8 * j + 2], lo[8 * j + 3])), "r"(pack_bf16(lo[8 * j + 4], lo[8 * j + 5])),
                       "r"(pack_bf16(lo[8 * j + 6], lo[8 * j + 7])) : "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmO, slab0, col + ep.plane_stride, row0);
          bulk_commit();
        }
        continue;
      }
      const uint32_t slab = slab0 + (uint32_t)((c % NSLAB) * 2048), my = my0 + (uint32_t)((c % NSLAB) * 2048);
      if (lane == 0) bulk_wait_read<NSLAB - 1>();       // the slab's previous TMA store has finished reading it
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = my + ((((uint32_t)j) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(v[8 * j], v[8 * j + 1])),
                     "r"(pack_bf16(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16(v[8 * j + 4], v[8 * j + 5])),
                     "r"(pack_bf16(v[8 * j + 6], v[8 * j + 7]))
                     : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmO, slab, col, row0);
        bulk_commit();
      }
    } else {   // EPI_RESIDUAL_F32: two units of 16 fp32 columns, reduce-added into the residual stream
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t slab = slab0 + (uint32_t)((u % NSLAB) * 2048), my = my0 + (uint32_t)((u % NSLAB) * 2048);
        if (lane == 0) bulk_wait_read<NSLAB - 1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = u * 16 + j * 4;
          const float4 l4 = *reinterpret_cast<const float4*>(sl + cl + e);
          const uint32_t a = my + ((((uint32_t)j) ^ sw) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[e] * l4.x), "f"(v[e + 1] * l4.y),
                       "f"(v[e + 2] * l4.z), "f"(v[e + 3] * l4.w)
                       : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (split == 0) tma_reduce_add_2d(tmO, slab, col + u * 16, row0);
          else tma_store_2d(tmP, slab, col + u * 16, part_row0 + quarter * 32);
          bulk_commit();
        }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
}

// ---- epilogue over the BLOCKED fp32 stream (EPI_RESIDUAL_BLK, EPI_PATCH_BLK) ------------------------------------------------
// x_new = x_old + ls * (acc + bias), in place.  The old values do not depend on the MMAs, and an L2 / HBM round trip is longer than one
// 32-column chunk of epilogue work (two epilogue warps per scheduler hide nothing), so they are prefetched TWO chunks ahead and ACROSS
// tiles: chunks 0 and 1 of the next tile of this CTA are requested while chunks 2 and 3 of the current one are processed (`xo` and
// `primed` live in the caller's tile loop).  EPI_RESIDUAL_BLK also emits the bf16 shadow -- transposed through the warp's 2 KB slab and
// written with plain 16-byte stores (8 rows x 64 contiguous bytes per warp instruction; a TMA store per chunk made the warp wait for
// the TMA engine to drain the slab four times per tile) -- and the thread's partial row statistics; EPI_PATCH_BLK maps GEMM row
// b*256+p to stream row b*257+1+p (ls == 1) and adds to the position row of the patch instead of the stream.
template <int EPI>
struct BlkTile {
  float4* xp;            // this thread's row in the blocked stream, first column group of its 128-column half (+32 per group)
  const float4* xin;     // what the update is added to (the stream, or the blocked position table for the patch embedding)
  int grow;              // GEMM row of this thread's TMEM lane
  bool ok;
  uint64_t pol;          // L2 policy of the stream accesses
  __device__ __forceinline__ BlkTile(const EpiP& ep, int m0, int n0, int quarter, int half, int lane, uint64_t pol_) : pol(pol_) {
    constexpr bool PATCH = EPI == EPI_PATCH_BLK;
    grow = m0 + quarter * 32 + lane;
    const int srow = PATCH ? grow + m0 / 256 + 1 : grow;         // stream row
    ok = grow < ep.rows;
    xp = reinterpret_cast<float4*>(ep.out) + xblk_f4(ok ? srow : 0, (n0 + half * 128) >> 2);
    xin = PATCH ? reinterpret_cast<const float4*>(ep.pos) + xblk_f4(grow & 255, (n0 + half * 128) >> 2) : xp;
  }
  __device__ __forceinline__ void load(float4 (&dst)[8], int chunk) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = ok ? ldcg_hint(xin + 32 * (chunk * 8 + j), pol) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

template <int EPI, class Hook = NoHook>
__device__ __forceinline__ void epilogue_tile_blk(const EpiP& ep, float* sepi, uint32_t sstage, uint32_t tfull_bar_addr, uint32_t aph, int as,
                                                  uint32_t tmem_base, int m0, int n0, int warp, int lane, float4 (&xo)[2][8], bool primed,
                                                  bool has_next, int m0_next, int n0_next, Hook hook = Hook()) {
  constexpr bool PATCH = EPI == EPI_PATCH_BLK;
  const int ew = warp - 2;
  const int quarter = warp & 3;
  const int half = ew >> 2;
  const int te = threadIdx.x - 64;
  float* sb = sepi + as * 512;
  float* sl = sb + 256;
  sb[te] = __ldg(ep.bias + n0 + te);
  sl[te] = (!PATCH && ep.ls) ? __ldg(ep.ls + n0 + te) : 1.0f;
  const uint64_t pol = l2_policy(ep.xhint != 0);
  const BlkTile<EPI> t(ep, m0, n0, quarter, half, lane, pol);
  const BlkTile<EPI> tn(ep, has_next ? m0_next : m0, has_next ? n0_next : n0, quarter, half, lane, pol);
  if (!primed) { t.load(xo[0], 0); t.load(xo[1], 1); }
  const uint32_t slab = sstage + (uint32_t)ew * 2048u;
  const uint32_t my = slab + (uint32_t)lane * 64u;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);          // 16-byte piece index ^= (row >> 1) & 3: conflict-free 16-byte column writes
  float p_sum = 0.f, p_sq = 0.f;
  // the patch embedding emits shadow + statistics too when it is given the buffers (no separate pass in front of the first q|k|v);
  // its GEMM row b*256+p is stream row b*257+1+p
  const bool emit = !PATCH || ep.shadow != nullptr;
  const int roff = PATCH ? m0 / 256 + 1 : 0;
  epi_bar_sync();
  mbar_wait(tfull_bar_addr, aph);
  tc_fence_after();
  hook();
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * 128);
  uint32_t r[2][32];
  tmem_ld32(taddr, r[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_wait_ld();
    if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
    const uint32_t(&rc)[32] = r[c & 1];
    const int cl = half * 128 + c * 32;
    float xn[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + cl + 4 * j);
      const float4 l4 = *reinterpret_cast<const float4*>(sl + cl + 4 * j);
      const float4 x4 = xo[c & 1][j];
      xn[4 * j] = fmaf(__uint_as_float(rc[4 * j]) + b4.x, l4.x, x4.x);
      xn[4 * j + 1] = fmaf(__uint_as_float(rc[4 * j + 1]) + b4.y, l4.y, x4.y);
      xn[4 * j + 2] = fmaf(__uint_as_float(rc[4 * j + 2]) + b4.z, l4.z, x4.z);
      xn[4 * j + 3] = fmaf(__uint_as_float(rc[4 * j + 3]) + b4.w, l4.w, x4.w);
    }
    if (c < 2) t.load(xo[c & 1], c + 2);                      // two chunks ahead ...
    else if (has_next) tn.load(xo[c & 1], c - 2);             // ... also across the tile boundary
    if (t.ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) stcg_hint(t.xp + 32 * (c * 8 + j), make_float4(xn[4 * j], xn[4 * j + 1], xn[4 * j + 2], xn[4 * j + 3]), pol);
    }
    if (emit) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { p_sum += xn[j]; p_sq = fmaf(xn[j], xn[j], p_sq); }
      __syncwarp();                                            // the previous chunk's read-back of the slab is complete
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = my + ((((uint32_t)j) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(xn[8 * j], xn[8 * j + 1])),
                     "r"(pack_bf16(xn[8 * j + 2], xn[8 * j + 3])), "r"(pack_bf16(xn[8 * j + 4], xn[8 * j + 5])),
                     "r"(pack_bf16(xn[8 * j + 6], xn[8 * j + 7]))
                     : "memory");
      }
      __syncwarp();
      const int rbase = m0 + quarter * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {                            // 8 rows x 64 bytes per warp instruction
        const int row = 8 * i + (lane >> 2), pc = lane & 3;
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                     : "r"(slab + (uint32_t)(row * 64) + ((((uint32_t)pc) ^ (uint32_t)((row >> 1) & 3)) << 4)) : "memory");
        if (rbase + row < ep.rows)
          __stcg(reinterpret_cast<uint4*>(ep.shadow + (int64_t)(rbase + row + roff) * DD + n0 + cl + pc * 8), make_uint4(v0, v1, v2, v3));
      }
    }
  }
  if (emit && t.ok)                                            // partial (sum, sumsq) of this row over this tile's 128-column half
    *reinterpret_cast<float2*>(ep.stats_out + (int64_t)(t.grow + roff) * 12 + ((n0 / BN) * 2 + half) * 2) = make_float2(p_sum, p_sq);
  tc_fence_before();
  __syncwarp();
}

// ---- the kernel ---------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, EpiP ep, int M, int N, int K) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = sB + STAGES * B_STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* sepi = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));
  const uint32_t sstage = (tmem_slot + 16 + 4096 + 1023u) & ~1023u;   // per-warp 2 KB TMA-store slabs

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = N / BN;
  const int n_tiles_m = (M + BM - 1) / BM;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int n_kb = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();                  // predecessor's outputs are complete and visible

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          mbar_expect_tx(full_bar + 8 * s, A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, full_bar + 8 * s, kb * BK, m0);
          tma_load_2d(sB + s * B_STAGE_BYTES, &tmB, full_bar + 8 * s, kb * BK, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(sA + s * A_STAGE_BYTES);
          const uint64_t db = make_smem_desc(sB + s * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 B (16 bf16) inside the 128 B swizzle atom: +2 in the (addr>>4) field
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + 8 * s);     // frees the smem stage when these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(tfull_bar + 8 * as);      // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue warps =====================
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      if (EPI == EPI_PATCH_F32) epilogue_tile_direct<EPI>(ep, sepi, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, M, warp, lane);
      else epilogue_tile_tma<EPI>(ep, &tmO, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane);
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
    }
    if (lane == 0) bulk_wait0();            // all TMA stores of this warp have completed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle
inline int make_map_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(HVLA_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HVLA_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return HVLA_OK;
}

// output map for the TMA epilogue: 32-row x 64-byte boxes, SWIZZLE_64B (bf16: 32 columns, fp32: 16 columns)
inline int make_map_out(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(HVLA_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const int es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)(64 / es), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HVLA_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed");
  return HVLA_OK;
}
// rows of the output tensor as the epilogue addresses them (patch epilogue does not use the map)
inline int make_out_map_for(CUtensorMap* mo, int epi, const EpiP& ep, int M) {
  if (epi == EPI_PATCH_F32 && ep.patch_rows) return make_map_out(mo, ep.out, (int64_t)(M / 256) * 257, ep.ldo, true);
  if (epi == EPI_PATCH_F32) { memset(mo, 0, sizeof *mo); return HVLA_OK; }
  return make_map_out(mo, ep.out, M, ep.ldo, epi == EPI_RESIDUAL_F32);
}

template <int EPI>
inline int launch_one(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const EpiP& ep, int M, int N,
                      int K) {
  static std::atomic<uint64_t> attr_set{0};   // per-device one-time setup
  if (device_once(attr_set)) {
    HVLA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ProfScope ps(st, "gemm_tc");
  launch_k(gemm_tc_kernel<EPI>, dim3(grid), dim3(NUM_THREADS), (size_t)SMEM_BYTES, st, ma, mb, mo, ep, M, N, K);
  HVLA_LAUNCH_CHECK("gemm_tc");
  return HVLA_OK;
}

// C = epi(A[M,K] * Wt[N,K]^T);  N % 256 == 0, K % 64 == 0
inline int gemm_tc(cudaStream_t st, const void* A, const void* Wt, int M, int N, int K, int epi, const EpiP& ep) {
  if (N % BN != 0 || K % BK != 0 || M <= 0) return fail(HVLA_ERR_ARG, "gemm_tc: N %% 256 or K %% 64 != 0");
  CUtensorMap ma, mb, mo;
  HVLA_TRY(make_map_bf16(&ma, A, M, K, BM));
  HVLA_TRY(make_map_bf16(&mb, Wt, N, K, BN));
  HVLA_TRY(make_out_map_for(&mo, epi, ep, M));
  switch (epi) {
    case EPI_BIAS_BF16: return launch_one<EPI_BIAS_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_BIAS_GELU_BF16: return launch_one<EPI_BIAS_GELU_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_RESIDUAL_F32: return launch_one<EPI_RESIDUAL_F32>(st, ma, mb, mo, ep, M, N, K);
    case EPI_PATCH_F32: return launch_one<EPI_PATCH_F32>(st, ma, mb, mo, ep, M, N, K);
  }
  return fail(HVLA_ERR_ARG, "gemm_tc: unknown epilogue");
}

}  // namespace tc
}  // namespace hvla
