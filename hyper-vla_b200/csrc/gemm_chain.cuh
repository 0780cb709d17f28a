// A CHAIN of dependent tcgen05 GEMMs in ONE persistent launch (flow B of dino_bf16, 25 to ~290 images):
//     proj (+ residual)  ->  fc1 (+ GELU)  ->  fc2 (+ residual)  ->  q|k|v of the next layer
// instead of four launches.  Same CTA-pair tiles, TMA ring, TMEM double buffering and epilogues as gemm_tc2_kernel; what changes
// is WHO computes WHAT and WHEN:
//   * work units (GEMM g, row block m, column tile n) of all GEMMs form one ordered list; CTA pairs pull the next unit with an
//     atomic counter, so there is no tail (195 tiles on 74 pairs = 2.6 waves cost 3) and no launch / drain / pipeline-fill gap
//     between the GEMMs of a layer;
//   * a unit of GEMM g > 0 needs row block m of GEMM g-1 complete: every CTA bumps done[g][m] when its half tile is in global
//     memory (a PUBLISHER warp does the release, fed lazily by the epilogue warps), and a SCHEDULER warp acquires
//     done[g-1][m] == 2 * n_tiles_n(g-1) BEFORE it posts the unit to the pair: producer, MMA thread and epilogue warps never see an
//     atomic or a counter;
//   * the list is parametrised as a wavefront over row blocks (slot s holds the tiles of proj(s), fc1(s - lag1), fc2(s - lag1 - lag2),
//     qkv(s - lag1 - lag2 - lag3)); the default lags are >= the number of row blocks, i.e. GEMM after GEMM: interleaving the GEMMs so
//     that a row block's activations are re-read out of L2 measured 3-15 % slower.
// No deadlock: every pair takes units in increasing list order, dependencies point to earlier list positions, only running pairs own
// units, and an epilogue warp hands its finished tile over before it blocks on a unit that is not posted yet.
// Bit-identical to one launch per GEMM for every schedule (tools/chain_check.py, tests/test_gpu_chain.py).
#pragma once
#include "gemm_tc2.cuh"

namespace hvla {
namespace chain {

using namespace tc;
using namespace tc2;

constexpr int MAX_G = 4;
constexpr int W_PUB = 2 + NUM_EPI_WARPS;                 // warp 10: publishes finished tiles (its release fence stalls nobody else)
constexpr int W_SCHED = W_PUB + 1;                       // warp 11 (leader CTA): pulls units, waits for their rows, posts them to the pair
constexpr int CHAIN_THREADS = 32 * (W_SCHED + 1);
constexpr int RING = 8;                                  // posted units kept per CTA (a consumer is at most 4 units behind the scheduler)

struct ChainGemm {
  int ntn;       // column tiles (N / 256)
  int nkb;       // K / 64
  int epi;       // tc::Epi
  int lag;       // wavefront lag in slots: slot s holds row block s - lag of this GEMM
  EpiP ep;
};
struct ChainP {
  int n_gemms, n_tiles_m, n_units, n_slots;
  int* next;       // [1]  next unit of the list (zeroed before the launch)
  int* done;       // [MAX_G][n_tiles_m]  half tiles of (g, m) that are complete in global memory (zeroed)
  ChainGemm g[MAX_G];
};
struct __align__(64) ChainMaps { CUtensorMap a[MAX_G], b[MAX_G], o[MAX_G]; };

// ---- the unit list (closed form: no table) ----
// units before slot s
__device__ __host__ __forceinline__ int units_before(const ChainP& p, int s) {
  int n = 0;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g)
    if (g < p.n_gemms) { int r = s - p.g[g].lag; r = r < 0 ? 0 : (r > p.n_tiles_m ? p.n_tiles_m : r); n += r * p.g[g].ntn; }
  return n;
}
struct Unit { int g, m, n; };
// `slot` is a cursor the caller keeps: the units of one pair only increase
__device__ __forceinline__ Unit decode_unit(const ChainP& p, int u, int& slot) {
  while (slot + 1 < p.n_slots && u >= units_before(p, slot + 1)) ++slot;
  int l = u - units_before(p, slot);
  Unit r; r.g = 0; r.m = 0; r.n = 0;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g) {
    if (g < p.n_gemms) {
      const int m = slot - p.g[g].lag;
      if (m >= 0 && m < p.n_tiles_m) {
        if (l >= 0 && l < p.g[g].ntn) { r.g = g; r.m = m; r.n = l; }
        l -= p.g[g].ntn;
      }
    }
  }
  return r;
}
// a posted unit: g | n << 2 | m << 8 (decoded once, by the scheduler); -1 = no more units
__device__ __forceinline__ int pack_unit(const Unit& u) { return u.g | (u.n << 2) | (u.m << 8); }
__device__ __forceinline__ Unit unpack_unit(int v) { Unit u; u.g = v & 3; u.n = (v >> 2) & 63; u.m = v >> 8; return u; }

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// mbarrier wait / arrive with CLUSTER-scope ordering: the ring entries of the peer CTA are written by the leader's scheduler thread
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// store `v` at shared-memory offset `addr` of CTA `cta` of the cluster, then arrive (release, cluster scope) on its barrier `bar`
__device__ __forceinline__ void post_remote(uint32_t addr, uint32_t bar, int v, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra, rb;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %3;\n\t"
      "mapa.shared::cluster.u32 rb, %1, %3;\n\t"
      "st.shared::cluster.s32 [ra], %2;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rb];\n\t"
      "}" ::"r"(addr), "r"(bar), "r"(v), "r"(cta) : "memory");
}

#ifdef HVLA_CHAIN_STATS
__device__ unsigned long long g_chain_stats[8];   // [0] units with a dependency, [1] of those that had to wait, [2] spin iterations, [3] cycles waited
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CHAIN_THREADS, 1)
gemm_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainP p) {
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
  const uint32_t sstage = (tmem_slot + 16 + 4096 + 1023u) & ~1023u;
  // control block in the alignment gap below the slabs (880 bytes): the ring of posted units (entry + "posted" barrier each), the
  // "taken" counter (producer -> scheduler) and the publisher mailbox (two barrier pairs + the done[] index of the tile in each slot)
  const uint32_t ctl = sstage - 256;
  const uint32_t ring_bar = ctl, ring_val = ctl + 8 * RING, taken_bar = ring_val + 4 * RING;
  const uint32_t pub_full = taken_bar + 8, pub_empty = pub_full + 16, pub_idx_a = pub_empty + 16;
  volatile int* ring = reinterpret_cast<volatile int*>(smem_raw + (ring_val - smem_u32(smem_raw)));
  volatile int* pub_idx = reinterpret_cast<volatile int*>(smem_raw + (pub_idx_a - smem_u32(smem_raw)));
  volatile int* taken_cnt = reinterpret_cast<volatile int*>(smem_raw + (taken_bar - smem_u32(smem_raw)));   // units the leader's producer has taken

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.n_gemms; ++g) {
      tma_prefetch_desc(&maps.a[g]);
      tma_prefetch_desc(&maps.b[g]);
      tma_prefetch_desc(&maps.o[g]);
    }
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 2 * NUM_EPI_WARPS);
      mbar_init(pub_full + 8 * s, NUM_EPI_WARPS);
      mbar_init(pub_empty + 8 * s, 1);
    }
    for (int s = 0; s < RING; ++s) mbar_init(ring_bar + 8 * s, 1);
    *taken_cnt = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  // the j-th unit posted to this CTA (blocks until the scheduler has posted it)
  auto posted = [&](int j) {
    mbar_wait_cluster(ring_bar + 8 * (j % RING), (j / RING) & 1);
    return (int)ring[j % RING];
  };

  if (warp == W_SCHED) {
    // ===================== scheduler (leader CTA) =====================
    // Pulls the next unit of the list with an atomic, waits until the rows it reads are complete (acquire), and only then posts it to
    // both CTAs: producers, MMA thread and epilogue warps never look at a dependency counter and never wait for the atomic's round
    // trip -- the pair holds exactly one unit beyond the one being loaded (unit j + 1 is pulled when the producer takes unit j).
    if (rank == 0 && lane == 0) {
      int slot = 0;
      for (int j = 0;; ++j) {
        if (j >= 1) { while (*taken_cnt < j) { } }               // the producer has taken unit j - 1
        const int u = atomicAdd(p.next, 1);
        int v = -1;
        if (u < p.n_units) {
          const Unit un = decode_unit(p, u, slot);
          if (un.g > 0) {
            const int* c = p.done + (un.g - 1) * p.n_tiles_m + un.m;
            const int target = 2 * p.g[un.g - 1].ntn;
#ifdef HVLA_CHAIN_NODEP
            (void)c; (void)target;
#elif defined(HVLA_CHAIN_STATS)
            const long long t0 = clock64();
            unsigned long long spins = 0;
            while (ld_acquire(c) < target) ++spins;
            atomicAdd(&g_chain_stats[0], 1ull);
            if (spins) { atomicAdd(&g_chain_stats[1], 1ull); atomicAdd(&g_chain_stats[2], spins); atomicAdd(&g_chain_stats[3], (unsigned long long)(clock64() - t0)); }
#else
            // the only unbounded wait of the protocol that is not an mbarrier: fail loudly (sticky launch error) instead of hanging if the
            // rows never complete -- which the list order rules out, so this only fires on a broken build or a corrupted workspace
            if (ld_acquire(c) < target) {
              const long long t0 = clock64();
              while (ld_acquire(c) < target) {
                if (clock64() - t0 > (1ll << 33)) __trap();             // ~5 s at 1.7 GHz
              }
            }
#endif
          }
          v = pack_unit(un);
        }
        const int rs = j % RING;
        ring[rs] = v;
        asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(ring_bar + 8 * rs) : "memory");
        post_remote(ring_val + 4 * rs, ring_bar + 8 * rs, v, 1);
        if (v < 0) break;
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0;; ++j) {
        const int v = posted(j);
        if (v < 0) break;
        if (rank == 0) *taken_cnt = j + 1;          // flow control only (no data behind it): the scheduler may pull the next unit
        const Unit un = unpack_unit(v);
#ifndef HVLA_CHAIN_NO_RFENCE
        fence_proxy_async_all();      // generic-proxy writes of other SMs (shadow rows), acquired by the scheduler, before this thread's TMA reads
#endif
        const int nkb = p.g[un.g].nkb;
        const int m0 = un.m * BM2 + (int)rank * 128;
        const int n0 = un.n * BN + (int)rank * 128;
        const CUtensorMap* ma = &maps.a[un.g];
        const CUtensorMap* mb = &maps.b[un.g];
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar + 8 * s, 2 * (A2_BYTES + B2_BYTES));
          tma_load_2d_2sm(sA + s * A2_BYTES, ma, full_bar + 8 * s, kb * BK, m0);
          tma_load_2d_2sm(sB + s * B2_BYTES, mb, full_bar + 8 * s, kb * BK, n0);
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
      for (int it = 0;; ++it) {
        const int v = posted(it);
        if (v < 0) break;
        const int nkb = p.g[v & 3].nkb;
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb) {
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
  } else if (warp == W_PUB) {
    // ===================== publisher: done[g][m] += 1 per finished half tile =====================
    // Every epilogue warp arrives on pub_full when ITS stores of a tile are complete (TMA) or issued (generic); this thread then makes
    // them visible device-wide (the release fence orders everything it observed through the mbarrier) and bumps the counter.  The
    // fence costs most of a microsecond: here it stalls nobody.
    if (lane == 0) {
      for (int t = 0;; ++t) {
        const int ps = t & 1;
        mbar_wait(pub_full + 8 * ps, (t >> 1) & 1);
        const int idx = pub_idx[ps];
        mbar_arrive(pub_empty + 8 * ps);
        if (idx < 0) break;
        red_release_add(p.done + idx, 1);
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 TMEM lanes) =====================
    int pend = -1;                                      // done[] index of the tile whose stores this warp has issued but not yet handed over
    int pt = 0;                                         // tiles handed to the publisher
    // Hand the previous tile to the publisher.  Normally called from inside the next tile's epilogue, after its accumulator is complete
    // and before its first global store (gemm_tc.cuh: Hook): the previous tile's TMA stores are a whole tile old, wait_group returns at
    // once.  Per warp, no CTA barrier.
    auto hand_over = [&](int idx) {
      const int ps = pt & 1;
      if (lane == 0) {
        bulk_wait0();                                   // this warp's TMA stores have completed (lane 0 issued them)
        if (warp == 2) {
          mbar_wait(pub_empty + 8 * ps, ((pt >> 1) & 1) ^ 1);
          pub_idx[ps] = idx;
        }
      }
#ifdef HVLA_CHAIN_WFENCE
      fence_proxy_async_all();
#endif
      __syncwarp();                                     // the other lanes' generic stores are ordered before the arrive (release, CTA scope)
      if (lane == 0) mbar_arrive(pub_full + 8 * ps);
      ++pt;
    };
    auto publish = [&]() {
      if (pend < 0) return;
      hand_over(pend);
      pend = -1;
    };
    for (int it = 0;; ++it) {
      // Liveness: the next unit may be held back because ITS rows are not complete, and the tile this warp has not handed over could be
      // what it (or another pair) is waiting for -- the deferred hand-over sits behind the next unit's MMAs.  Hand over before blocking.
      int ready = 0;
      if (lane == 0) ready = mbar_test_cluster(ring_bar + 8 * (it % RING), (it / RING) & 1) ? 1 : 0;
      if (!__shfl_sync(0xffffffffu, ready, 0)) publish();                 // warp-uniform decision
      const int v = posted(it);
      if (v < 0) break;
      const Unit un = unpack_unit(v);
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m0 = un.m * BM2 + (int)rank * 128, n0 = un.n * BN;
      // the chain is proj -> fc1 -> fc2 -> q|k|v (checked on the host): constant indices keep the epilogue parameters in the constant bank
      float4 xo[2][8];
      switch (un.g) {
        case 0:
          epilogue_tile_blk<EPI_RESIDUAL_BLK>(p.g[0].ep, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane, xo, false, false, m0, n0, publish);
          break;
        case 1:
          epilogue_tile_tma<EPI_BIAS_GELU_BF16_FOLD, NSLAB2>(p.g[1].ep, &maps.o[1], sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane,
                                                              0, nullptr, 0, nullptr, publish);
          break;
        case 2:
          epilogue_tile_blk<EPI_RESIDUAL_BLK>(p.g[2].ep, sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane, xo, false, false, m0, n0, publish);
          break;
        default:
          epilogue_tile_tma<EPI_BIAS_BF16_FOLD, NSLAB2>(p.g[3].ep, &maps.o[3], sepi, sstage, tfull_bar + 8 * as, aph, as, tmem_base, m0, n0, warp, lane,
                                                         0, nullptr, 0, nullptr, publish);
          break;
      }
      if (lane == 0) mbar_arrive_remote(tempty_bar + 8 * as, 0);
      pend = un.g * p.n_tiles_m + un.m;
    }
    publish();
    hand_over(-1);                                      // tells the publisher to stop
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct ChainDesc {          // one GEMM of the chain: C = epi(A[M,K] * Wt[N,K]^T)
  const void* A;
  const void* Wt;
  int N, K, epi;
  EpiP ep;
};

// ints of workspace one chain launch needs for M rows (unit counter, done[])
inline size_t chain_ws_ints(int64_t M) {
  const int64_t ntm = (M + BM2 - 1) / BM2;
  return (size_t)(16 + MAX_G * ntm);
}

inline int gemm_chain(cudaStream_t st, const ChainDesc* d, int n_gemms, int M, int* ws, const int* lags) {
  if (n_gemms < 3 || n_gemms > MAX_G) return fail(HVLA_ERR_ARG, "gemm_chain: 3 or 4 GEMMs");
  static std::atomic<uint64_t> attr_set{0};
  if (device_once(attr_set)) {
    HVLA_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  }
  const int pairs = num_sms() / 2;
  ChainMaps maps;
  ChainP p;
  memset(&maps, 0, sizeof maps);
  memset(&p, 0, sizeof p);
  p.n_gemms = n_gemms;
  p.n_tiles_m = (M + BM2 - 1) / BM2;
  int units = 0, lag_total = 0;
  for (int g = 0; g < n_gemms; ++g) {
    const ChainDesc& c = d[g];
    if (c.N % BN != 0 || c.K % BK != 0) return fail(HVLA_ERR_ARG, "gemm_chain: N %% 256 or K %% 64 != 0");
    static const int pattern[MAX_G] = {EPI_RESIDUAL_BLK, EPI_BIAS_GELU_BF16_FOLD, EPI_RESIDUAL_BLK, EPI_BIAS_BF16_FOLD};
    if (c.epi != pattern[g]) return fail(HVLA_ERR_ARG, "gemm_chain: the chain is proj -> fc1 -> fc2 -> q|k|v");
    HVLA_TRY(make_map_bf16(&maps.a[g], c.A, M, c.K, 128));
    HVLA_TRY(make_map_bf16(&maps.b[g], c.Wt, c.N, c.K, 128));
    if (c.epi == EPI_RESIDUAL_BLK) {
      if (!c.ep.shadow || !c.ep.stats_out || c.N != DD) return fail(HVLA_ERR_ARG, "gemm_chain: EPI_RESIDUAL_BLK needs shadow, stats_out and N == 768");
      HVLA_TRY(make_map_out(&maps.o[g], c.ep.shadow, M, c.N, false));
    } else {
      if (!c.ep.stats || !c.ep.cs) return fail(HVLA_ERR_ARG, "gemm_chain: folded-LayerNorm epilogue needs stats and cs");
      HVLA_TRY(make_map_out(&maps.o[g], c.ep.out, M, c.ep.ldo, false));
    }
    p.g[g].ntn = c.N / BN;
    p.g[g].nkb = c.K / BK;
    p.g[g].epi = c.epi;
    p.g[g].lag = lag_total;
    if (g + 1 < n_gemms) lag_total += lags[g] < p.n_tiles_m ? lags[g] : p.n_tiles_m;      // lag >= row blocks: GEMM g + 1 starts when GEMM g's list is exhausted
    p.g[g].ep = c.ep;
    p.g[g].ep.rows = M;
    units += p.n_tiles_m * p.g[g].ntn;
  }
  if (p.n_tiles_m >= (1 << 22)) return fail(HVLA_ERR_ARG, "gemm_chain: too many row blocks");
  p.n_units = units;
  p.n_slots = p.n_tiles_m + lag_total;
  p.next = ws;
  p.done = ws + 16;
  ProfScope ps(st, "gemm_tc");
  launch_k(gemm_chain_kernel, dim3(2 * pairs), dim3(CHAIN_THREADS), (size_t)SMEM2_BYTES, st, maps, p);
  HVLA_LAUNCH_CHECK("gemm_chain");
  return HVLA_OK;
}

}  // namespace chain
}  // namespace hvla
