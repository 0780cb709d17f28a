// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a CTA pair on one TPC computes a 256x256 tile.
// Each CTA stages its own 128 rows of A and its own 128-row half of Wt (32 KB per 64-wide K block
// instead of 48 KB), so the L2->SM operand traffic per FLOP drops by 1/3 and 6 stages fit in smem;
// the leader CTA issues tcgen05.mma.cta_group::2 (M256 N256 K16) for both tensor cores, every CTA
// runs its own TMA producer and its own epilogue over its 128 TMEM lanes.
//   full[s]   (leader)     : 1 arrival (leader producer's expect_tx of BOTH CTAs' bytes) + TMA complete_tx
//   empty[s]  (both CTAs)  : tcgen05.commit multicast from the leader's MMA thread
//   tfull[a]  (both CTAs)  : tcgen05.commit multicast when a tile's accumulator is complete
//   tempty[a] (leader)     : 2 x 8 epilogue-warp arrivals (the peer's arrive remotely)
#pragma once
#include "gemm_tc.cuh"
#include <stdlib.h>

namespace hvla {
namespace tc2 {

using namespace tc;   // PTX wrappers, EpiP, Epi, descriptors

constexpr int BM2 = 256;                         // pair tile rows (128 per CTA)
constexpr int STAGES2 = 6;
constexpr int NSLAB2 = 1;                         // TMA-store slabs per epilogue warp (2 = double-buffered: measured no gain, and costs a pipeline stage)
constexpr int A2_BYTES = 128 * BK * 2;           // 16 KB
constexpr int B2_BYTES = 128 * BK * 2;           // 16 KB
constexpr int SMEM2_BYTES = STAGES2 * (A2_BYTES + B2_BYTES) + 1024 + 256 + 4096 + 1024 + 8 * 2048 * NSLAB2;
static_assert(SMEM2_BYTES <= 232448, "shared memory budget");
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive on the same-offset barrier of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32  remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64  _, [remAddr32];\n\t"
      "}" ::"r"(bar), "r"(cta)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmX, EpiP ep, int M, int N, int K, int splits) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES2 * A2_BYTES;
  const uint32_t bars = sB + STAGES2 * B2_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES2;
  const uint32_t tfull_bar = bars + 16 * STAGES2, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* sepi = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));
  const uint32_t sstage = (tmem_slot + 16 + 4096 + 1023u) & ~1023u;   // per-warp 2 KB TMA-store slabs

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles_n = N / BN;
  const int n_tiles_m = (M + BM2 - 1) / BM2;
  const int n_tiles = n_tiles_m * n_tiles_n * splits;   // work units: (tile, K split); splits > 1 only with the reduce-add epilogue
  const int n_kb = K / BK / splits;                     // K blocks per unit
  const bool rev = ep.rev != 0;                         // serpentine tile order (gemm_tc.cuh): last row block first
  auto tile_m = [&](int tile) { const int tm = tile / n_tiles_n; return rev ? n_tiles_m - 1 - tm : tm; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != EPI_PATCH_BLK) tma_prefetch_desc(&tmO);
    if (splits > 1) tma_prefetch_desc(&tmP);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 2 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();                  // predecessor's outputs (A operand, residual stream) are complete and visible

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = pair; unit < n_tiles; unit += n_pairs) {
        const int tile = unit / splits, kb0 = (unit % splits) * n_kb;
        const int m0 = tile_m(tile) * BM2 + (int)rank * 128;
        const int n0 = (tile % n_tiles_n) * BN + (int)rank * 128;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar + 8 * s, 2 * (A2_BYTES + B2_BYTES));
          tma_load_2d_2sm(sA + s * A2_BYTES, &tmA, full_bar + 8 * s, (kb0 + kb) * BK, m0);
          tma_load_2d_2sm(sB + s * B2_BYTES, &tmB, full_bar + 8 * s, (kb0 + kb) * BK, n0);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM2, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(sA + s * A2_BYTES);
          const uint64_t db = make_smem_desc(sB + s * B2_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(empty_bar + 8 * s);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
        umma_commit_2sm(tfull_bar + 8 * as);
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 TMEM lanes) =====================
    int it = 0;
    float4 xo[epi_blk(EPI) ? 2 : 1][8];               // blocked-stream epilogue: old values prefetched across tiles
    bool primed = false;
    for (int unit = pair; unit < n_tiles; unit += n_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int tile = unit / splits;
      const int m0 = tile_m(tile) * BM2 + (int)rank * 128, n0 = (tile % n_tiles_n) * BN;
      if constexpr (epi_blk(EPI)) {                   // blocked stream: in-place update (+ bf16 shadow, + row statistics); splits == 1
        const int nxt = unit + n_pairs;
        const bool has_next = nxt < n_tiles;
        epilogue_tile_blk<EPI>(ep, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane, xo, primed, has_next,
                               tile_m(nxt) * BM2 + (int)rank * 128, (nxt % n_tiles_n) * BN);
        primed = has_next;
      } else if (EPI == EPI_PATCH_F32 && ep.patch_rows)      // patch embedding accumulated onto the pre-initialised stream rows (TMA reduce-add)
        epilogue_tile_tma<EPI_RESIDUAL_F32, NSLAB2>(ep, &tmO, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane);
      else if (EPI == EPI_PATCH_F32) epilogue_tile_direct<EPI>(ep, sepi, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, M, warp, lane);
      else {
        const int split = unit % splits;          // split s > 0 stores to block s-1 of the partial-product scratch
        epilogue_tile_tma<EPI, NSLAB2>(ep, &tmO, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane, split, &tmP,
                               (split - 1) * n_tiles_m * BM2 + m0, &tmX);
      }
      if (lane == 0) mbar_arrive_remote(tempty_bar + 8 * as, 0);
    }
    if (lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still multicast into this CTA's barriers until it is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
inline int launch_one2(cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const EpiP& ep, int M, int N,
                       int K) {
  static std::atomic<uint64_t> attr_set{0};   // per-device one-time setup
  if (device_once(attr_set)) {
    HVLA_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  }
  const int tiles0 = ((M + BM2 - 1) / BM2) * (N / BN);
  int pairs = num_sms() / 2;
  // Split K over several CTA pairs when there are too few tiles to fill the chip (small batches: 6 tiles at batch 1).
  // Only the residual GEMMs do it (see "split-K partial products" in gemm_tc.cuh): deterministic, but the last bits
  // differ from the unsplit computation of the same rows in a large batch.  HVLA_GEMM_SPLITK=0 turns it off.
  int splits = 1;
  static const bool splitk = !(getenv("HVLA_GEMM_SPLITK") && getenv("HVLA_GEMM_SPLITK")[0] == '0');
  CUtensorMap mp = mo;
  if (EPI == EPI_RESIDUAL_F32 && splitk && ep.part != nullptr) {
    const int nkb = K / BK;
    const size_t block = (size_t)((M + BM2 - 1) / BM2) * BM2 * (size_t)N * 4;
    while (splits < 8 && tiles0 * splits * 2 <= pairs && nkb % (splits * 2) == 0 && block * (splits * 2 - 1) <= ep.part_bytes)
      splits *= 2;
    if (splits > 1) HVLA_TRY(make_map_out(&mp, ep.part, (int64_t)(splits - 1) * ((M + BM2 - 1) / BM2) * BM2, N, true));
  }
  if (ep.splits_used) *ep.splits_used = splits;
  EpiP ep2 = ep;
  ep2.rows = M;
  CUtensorMap mx = mo;
  const int tiles = tiles0 * splits;
  if (const char* e = getenv("HVLA_GEMM_MAX_PAIRS")) { const int v = atoi(e); if (v > 0 && v < pairs) pairs = v; }   // experiment knob
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  ProfScope ps(st, "gemm_tc");
  launch_k(gemm_tc2_kernel<EPI>, dim3(grid), dim3(NUM_THREADS), (size_t)SMEM2_BYTES, st, ma, mb, mo, mp, mx, ep2, M, N, K, splits);
  HVLA_LAUNCH_CHECK("gemm_tc2");
  return HVLA_OK;
}

inline int gemm_tc2(cudaStream_t st, const void* A, const void* Wt, int M, int N, int K, int epi, const EpiP& ep) {
  if (N % BN != 0 || K % BK != 0 || M <= 0) return fail(HVLA_ERR_ARG, "gemm_tc2: N %% 256 or K %% 64 != 0");
  CUtensorMap ma, mb, mo;
  HVLA_TRY(make_map_bf16(&ma, A, M, K, 128));
  HVLA_TRY(make_map_bf16(&mb, Wt, N, K, 128));
  if (epi == EPI_RESIDUAL_BLK) {             // the only TMA output of this epilogue is the bf16 shadow of the stream
    if (!ep.shadow || !ep.stats_out || N != DD) return fail(HVLA_ERR_ARG, "gemm_tc2: EPI_RESIDUAL_BLK needs shadow, stats_out and N == 768");
    HVLA_TRY(make_map_out(&mo, ep.shadow, M, N, false));
  } else if (epi == EPI_PATCH_BLK) {
    if (N != DD || M % 256 != 0) return fail(HVLA_ERR_ARG, "gemm_tc2: EPI_PATCH_BLK needs N == 768 and whole images");
    memset(&mo, 0, sizeof mo);
  } else {
    if (epi_fold(epi) && (!ep.stats || !ep.cs)) return fail(HVLA_ERR_ARG, "gemm_tc2: folded-LayerNorm epilogue needs stats and cs");
    if (epi_split(epi) && !((ep.nplanes == 2 || ep.nplanes == 3) && ep.plane_stride >= N && ep.ldo >= ep.nplanes * ep.plane_stride))
      return fail(HVLA_ERR_ARG, "gemm_tc2: split epilogue needs nplanes in {2,3}, plane_stride >= N and ldo >= nplanes * plane_stride");
    HVLA_TRY(make_out_map_for(&mo, epi, ep, M));
  }
  switch (epi) {
    case EPI_BIAS_BF16_FOLD: return launch_one2<EPI_BIAS_BF16_FOLD>(st, ma, mb, mo, ep, M, N, K);
    case EPI_BIAS_GELU_BF16_FOLD: return launch_one2<EPI_BIAS_GELU_BF16_FOLD>(st, ma, mb, mo, ep, M, N, K);
    case EPI_RESIDUAL_BLK: return launch_one2<EPI_RESIDUAL_BLK>(st, ma, mb, mo, ep, M, N, K);
    case EPI_PATCH_BLK: return launch_one2<EPI_PATCH_BLK>(st, ma, mb, mo, ep, M, N, K);
    case EPI_SPLIT_BF16: return launch_one2<EPI_SPLIT_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_SPLIT_GELU_BF16: return launch_one2<EPI_SPLIT_GELU_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_BIAS_BF16: return launch_one2<EPI_BIAS_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_BIAS_GELU_BF16: return launch_one2<EPI_BIAS_GELU_BF16>(st, ma, mb, mo, ep, M, N, K);
    case EPI_RESIDUAL_F32: return launch_one2<EPI_RESIDUAL_F32>(st, ma, mb, mo, ep, M, N, K);
    case EPI_PATCH_F32: return launch_one2<EPI_PATCH_F32>(st, ma, mb, mo, ep, M, N, K);
  }
  return fail(HVLA_ERR_ARG, "gemm_tc2: unknown epilogue");
}

}  // namespace tc2
}  // namespace hvla
