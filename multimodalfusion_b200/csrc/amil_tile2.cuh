// amil_tile2.cuh — CTA-pair version of the fused gated attention-MIL tile kernel (sm_100a).
//
// Same math and data flow as amil_tile.cuh (GEMM1 -> H in swizzled smem -> chunked GEMM2 -> gate
// epilogue -> softmax partial / gate backward), re-laid for the hardware limits ncu showed on the
// single-CTA version (tensor pipe 24 % active, the rest load latency and a serial epilogue):
//
//   * a cluster of 2 CTAs (one TPC) owns 256 rows and issues `tcgen05.mma.cta_group::2` (M = 256):
//     each CTA keeps its own 128 rows of X / H / accumulators but stages only HALF of every weight
//     tile, so the bytes a CTA pulls through L2 per k-block drop from 80 KB to 48 KB and the GEMM1
//     ring deepens from 2 to 4 stages (192 KB in flight per SM pair member);
//   * the producer prefetches its whole X tile into L2 up front (cp.async.bulk.prefetch.tensor), so
//     the HBM latency of the only operand that is not L2-resident is paid once, not per k-block;
//   * 8 epilogue warps (two per TMEM lane quadrant, splitting the columns) with TMEM loads issued
//     one piece ahead, and b1 / [ba|bb] / wc / dM staged once in shared memory instead of being
//     re-fetched through L1 for every 32-column piece.
//
// Barrier protocol (every barrier exists at the same offset in both CTAs):
//   full*[s]      leader's copy is used; armed by the leader's producer with 2 x stage bytes; the TMA
//                 loads of BOTH CTAs complete on it (.cta_group::2, peer bit cleared)
//   empty*[s]     each CTA's own copy; the leader's MMA thread commits to both (multicast 0b11)
//   acc1, acc2_full[b]   own copy per CTA; multicast commits
//   h_ready, acc2_empty[b]   leader's copy; one arrive per epilogue warp of either CTA (count 16),
//                 remote arrives go through mapa + mbarrier.arrive.release.cluster
#pragma once
#include "amil_tile.cuh"
#include "amil_head_tail.cuh"

namespace mmf {

template <int L, int D, bool GATED>
struct Amil2Cfg {
  static constexpr int KB1 = 1024 / 64;
  static constexpr int NH1 = L / 256;
  static constexpr int KB2 = L / 64;
  static constexpr int NCH = D / 128;
  static constexpr int CHN = GATED ? 256 : 128;       // MMA N of GEMM2 (per pair)
  static constexpr int KD = GATED ? 2 * D : D;
  static constexpr uint32_t H_BYTES = 128u * L * 2u;
  static constexpr uint32_t STAGE1 = 16384u + NH1 * 16384u;  // x tile + this CTA's half of W1
  static constexpr uint32_t STAGE2 = (CHN / 2) * 128u;       // this CTA's half of the Wab chunk
  static constexpr uint32_t POOL = 208u * 1024u;
  static constexpr uint32_t VEC_BYTES = 16u * 1024u;
#ifndef MMF_STASH_TMA
#define MMF_STASH_TMA 1
#endif
  // per-epilogue-warp staging of the activation stash: MMF_STASH_TMA = 1: one 32 x 32 fp16 block per branch (2 x 2 KB),
  // written out by TMA stores (64-byte swizzle); = 0: one 2 KB transpose scratch read back for st.global.v4
  static constexpr uint32_t XPOSE_WARP = MMF_STASH_TMA ? 4096u : 2048u;
  static constexpr uint32_t XPOSE_BYTES = 8u * XPOSE_WARP;
  static constexpr int NS1 = (POOL / STAGE1) < 6 ? (POOL / STAGE1) : 6;
  // GEMM2 ring: NS2 stages exist (barriers, addresses); the training forward uses only the first NS2_STASH of them — its
  // stash staging overlays the tail of the ring (run-time depth: one kernel instantiation serves both forms)
  static constexpr int NS2 = ((POOL - H_BYTES) / STAGE2) < 8 ? ((POOL - H_BYTES) / STAGE2) : 8;
  static constexpr int NS2_STASH = ((POOL - H_BYTES - XPOSE_BYTES) / STAGE2) < 8 ? ((POOL - H_BYTES - XPOSE_BYTES) / STAGE2) : 8;
  static constexpr uint32_t SMEM_BYTES = POOL + VEC_BYTES + 1024u;
  static constexpr int NCOLS = GATED ? 3 * D : 2 * D;
  // float offsets inside the vector region
  static constexpr int V_B1 = 0, V_BAB = L, V_WC = L + KD, V_DM = L + KD + D, V_S = V_DM + L,
                       V_P = V_S + 256, V_RED = V_P + 128, V_END = V_RED + 16;
  static_assert(V_END * 4 <= (int)VEC_BYTES, "vector region overflow");
  static_assert(NS1 >= 2 && NS2_STASH >= 2, "ring too shallow");
};

// MMF_TILE2_CHUNK_STAMPS = 1 (debug build, tools/phase_chunks.py): stamps 2 / 3 / 4 = "GEMM2 chunk c accumulators seen" and
// 5 / 9 / 15 = "chunk c gate epilogue done" (epilogue thread 0) replace the producer / first-stage / vectors-staged stamps
// MMF_L2_HINTS (mmf_ptx.cuh): >= 1 evict_first on the x stream (forward and wgrad)
// MMF_X_BULK_PREFETCH = 1 (A/B candidate): this CTA's x tile requested with one bulk L2 prefetch BEFORE griddepcontrol.wait
// MMF_DISCARD_DEAD_DU = 1: the training forward drops the dirty L2 lines of the previous step's dU (see EPI1's prologue)
#ifndef MMF_DISCARD_DEAD_DU
#define MMF_DISCARD_DEAD_DU 1
#endif
// MMF_DISCARD_DEAD_STASH: every CTA also drops its own tile's dead H / dG lines (the previous step's) before rewriting them.
//   1: from the epilogue warps' idle window next to the dU discard — 2560 more CCTL per CTA overflow the window into EPI1:
//      SLOWER (91.5 -> 93.2 us per step, profiles/r02_ab_discard_stash.txt);
//   2: from the 32 otherwise idle lanes of warp 3 through the whole GEMM1 + EPI1 phase: 90.85 -> 90.4 us per step in three
//      interleaved A/B pairs on one box (profiles/r02_ab_discard_stash2.txt) — but the forward timed ALONE (bench.py's
//      roofline stage: the same workspace re-written launch after launch) slows from 35.2 to 36.6 us, the discarded lines
//      being re-allocated by the stash stores; 0 (default): off. A 0.5 % step gain is not worth a forward kernel that is
//      slower in isolation.
#ifndef MMF_DISCARD_DEAD_STASH
#define MMF_DISCARD_DEAD_STASH 0
#endif
#ifndef MMF_X_BULK_PREFETCH
#define MMF_X_BULK_PREFETCH 0
#endif
#ifndef MMF_TILE2_CHUNK_STAMPS
#define MMF_TILE2_CHUNK_STAMPS 0
#endif
#define MMF_STAMP_N(a, i) do { if (!MMF_TILE2_CHUNK_STAMPS) MMF_STAMP(a, i); } while (0)
constexpr int AMIL2_THREADS = 384;
constexpr uint32_t AMIL2_EPI_THREADS = 256;

// DROPH / DROPA: train-mode dropout on h / on the attention branches, compile-time: with run-time flags the mask
// selection (shift, and, compare, select per element and branch) was executed even with dropout off and made up
// 10 of the ~22 instructions per element pair of the gate epilogue, which bounds the GEMM2 phase (issue-bound).
// (Round-2 A/B, gpurun_out/ab_r2a.txt: relaying the epilogue's cluster arrives through one thread and starting GEMM2
// chunk 0 half-way through EPI1 both measured within noise of this version (112.2 vs 111.6 / 112.4 us per step) and
// were removed.)

template <int L, int D, bool GATED, int MODE, bool DROPH, bool DROPA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AMIL2_THREADS, 1)
amil_tile2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmWab, const __grid_constant__ CUtensorMap tmH,
                  const __grid_constant__ CUtensorMap tmWk, const __grid_constant__ CUtensorMap tmAGs,
                  const AmilArgs a_in) {
  using C = Amil2Cfg<L, D, GATED>;
  AmilArgs a = a_in;   // (pointers to data the previous kernels wrote are re-derived after griddepcontrol.wait)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full1[C::NS1], bar_empty1[C::NS1];
  __shared__ __align__(8) uint64_t bar_full2[C::NS2], bar_empty2[C::NS2];
  __shared__ __align__(8) uint64_t bar_acc1, bar_h, bar_acc2_full[2], bar_acc2_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pool = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t h_base = pool;
  const uint32_t ring2 = pool + C::H_BYTES;
  float* vec = reinterpret_cast<float*>(smem_raw + (pool - smem_u32(smem_raw)) + C::POOL);
  const int tile = blockIdx.x;
  const long long row0 = (long long)tile * 128;
  const int ns2 = (MODE == AMIL_FWD && a.AG != nullptr) ? C::NS2_STASH : C::NS2;   // GEMM2 ring depth of this launch

  Timeline tl = timeline_start(MODE == AMIL_FWD ? 0 : 4);
  griddep_launch_dependents();   // PDL: the next kernel's prologue may overlap this kernel's tail
  if (threadIdx.x == 0) {
    MMF_STAMP(a, 0);
    if (MMF_X_BULK_PREFETCH && a.x_bulk != nullptr && row0 < a.N) {
      // This CTA's x tile is requested from HBM NOW, before griddepcontrol.wait, with ONE bulk prefetch: the bag is an
      // input of the step, not an output of the preceding kernel (and an L2 prefetch of a line somebody is still writing is
      // harmless — L2 is the point of coherence), so its HBM latency overlaps the predecessor's tail, the launch gap and this
      // prologue (GEMM1 took 21.6k cycles inside the step against 16.4k at the tensor peak, gpurun_out/r2i_instep_fwd.log)
      const long long rows = min((long long)128, a.N - row0);
      const uint32_t row_bytes = (uint32_t)(a.kb1 > 0 ? a.kb1 : C::KB1) * 128u;
      bulk_prefetch_l2_hint(reinterpret_cast<const uint8_t*>(a.x_bulk) + row0 * row_bytes, (uint32_t)rows * row_bytes,
                            l2_policy_evict_first());
    }
    for (int s = 0; s < C::NS1; ++s) { mbar_init(smem_u32(&bar_full1[s]), 1); mbar_init(smem_u32(&bar_empty1[s]), 1); }
    for (int s = 0; s < C::NS2; ++s) { mbar_init(smem_u32(&bar_full2[s]), 1); mbar_init(smem_u32(&bar_empty2[s]), 1); }
    mbar_init(smem_u32(&bar_acc1), 1);
    mbar_init(smem_u32(&bar_h), 16);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_acc2_full[b]), 1);
      mbar_init(smem_u32(&bar_acc2_empty[b]), 16);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmWab);
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  griddep_wait();                // PDL: everything above overlapped the previous kernel; its outputs are visible now
  a.b1 = pdl_fresh(a.b1); a.bab = pdl_fresh(a.bab); a.wc = pdl_fresh(a.wc); a.bc = pdl_fresh(a.bc);   // (optimizer step)
  a.head_wk = pdl_fresh(a.head_wk); a.tile_valid = pdl_fresh(a.tile_valid);
  if (DROPH || DROPA) a.seed = seed_resolve(a.seed);   // (device-resident seed of a graph-captured step)
  if (MODE == AMIL_BWD_GATE) {
    a.A_raw = pdl_fresh(a.A_raw); a.ml = pdl_fresh(a.ml); a.M = pdl_fresh(a.M); a.dM = pdl_fresh(a.dM); a.dA_raw = pdl_fresh(a.dA_raw);
  }
  timeline_wait_done(tl);
  if (threadIdx.x == 0) MMF_STAMP(a, 1);

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer (both CTAs) ==========================
    const int kb1 = a.kb1 > 0 ? a.kb1 : C::KB1;   // 16 k-blocks of x, or 48 in the split-precision form (MMF_PRECISE_FC)
    // x streams through L2 once per kernel (32 MB per 16k bag): evict_first keeps it from displacing the step's stash
    // (H, [a|g] -> dG, dU: 56 MB that the backward re-reads and the next step overwrites IN PLACE — as long as those lines
    // stay in L2 they are never written back; with default priority the x stream evicted them dirty: 63 MB of DRAM
    // writes per step in bursts on the kernels' critical paths, profiles/r02a_ncu_step_summary.md)
    const uint64_t pol_x = l2_policy_evict_first();
    for (int kb = 0; kb < kb1; ++kb) {
      // the ring's first NS1 loads go out first; only then is the rest of this CTA's x tile prefetched
      // into L2 (issuing all 16 prefetches up front queued the first real load behind them:
      // first stage landed 8k cycles after the cluster sync)
      if (kb == C::NS1 && !(MMF_X_BULK_PREFETCH && a.x_bulk != nullptr))
        for (int kp = C::NS1; kp < kb1; ++kp) {
          if (MMF_L2_HINTS) tma_prefetch_l2_2d_hint(&tmX, kp * 64, (int)row0, pol_x);
          else tma_prefetch_l2_2d(&tmX, kp * 64, (int)row0);
        }
      const int s = kb % C::NS1;
      const uint32_t ph = (kb / C::NS1) & 1;
      mbar_wait(smem_u32(&bar_empty1[s]), ph ^ 1);
      const uint32_t full = smem_u32(&bar_full1[s]);
      const uint32_t dst = pool + s * C::STAGE1;
      if (leader) mbar_arrive_expect_tx(full, 2 * C::STAGE1);
      if (MMF_L2_HINTS) tma_load_2d_pair_hint(dst, &tmX, full, kb * 64, (int)row0, pol_x);
      else tma_load_2d_pair(dst, &tmX, full, kb * 64, (int)row0);
#pragma unroll
      for (int j = 0; j < C::NH1; ++j)
        tma_load_2d_pair(dst + 16384 + j * 16384, &tmW1, full, kb * 64, j * 256 + 128 * (int)rank);
    }
    MMF_STAMP_N(a, 2);
    mbar_wait(smem_u32(&bar_acc1), 0);   // GEMM1 retired: its ring (overlaying H / ring2) is free
    MMF_STAMP_N(a, 3);
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < C::NCH; ++c) {
      for (int kb = 0; kb < C::KB2; ++kb) {
        mbar_wait(smem_u32(&bar_empty2[s]), ph ^ 1);
        const uint32_t full = smem_u32(&bar_full2[s]);
        if (leader) mbar_arrive_expect_tx(full, 2 * C::STAGE2);
        tma_load_2d_pair(ring2 + s * C::STAGE2, &tmWab, full, kb * 64, c * C::CHN + (C::CHN / 2) * (int)rank);
        if (++s == ns2) { s = 0; ph ^= 1u; }
      }
    }
    if (MODE == AMIL_FWD && a.z_out != nullptr) {
      // side product z = Wk h (head-projected backward): this CTA's 8 rows of the bf16 hi / lo split of the
      // classifier, all KB2 k-blocks in ONE ring stage (8 rows x 128 B = one swizzle atom per k-block)
      mbar_wait(smem_u32(&bar_empty2[s]), ph ^ 1);
      const uint32_t full = smem_u32(&bar_full2[s]);
      if (leader) mbar_arrive_expect_tx(full, 2 * C::KB2 * 1024);
      for (int kb = 0; kb < C::KB2; ++kb)
        tma_load_2d_pair(ring2 + s * C::STAGE2 + kb * 1024, &tmWk, full, kb * 64, 8 * (int)rank);
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // =============================== MMA issuer (leader CTA) ===========================
    constexpr uint32_t idesc1 = umma_idesc_bf16(256, 256, 0, 0);
    constexpr uint32_t idesc2 = umma_idesc_bf16(256, C::CHN, 0, 0);
    const int kb1 = a.kb1 > 0 ? a.kb1 : C::KB1;
    for (int kb = 0; kb < kb1; ++kb) {
      const int s = kb % C::NS1;
      const uint32_t ph = (kb / C::NS1) & 1;
      mbar_wait(smem_u32(&bar_full1[s]), ph);
      tc_fence_after();
      if (kb == 0) MMF_STAMP_N(a, 5);
      const uint32_t xs = pool + s * C::STAGE1;
      const uint32_t ws = xs + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_desc_sw128(xs + k * 32, 16, 1024);
#pragma unroll
        for (int j = 0; j < C::NH1; ++j)
          umma_bf16_ss_pair(tmem + j * 256, ad, umma_desc_sw128(ws + j * 16384 + k * 32, 16, 1024), idesc1,
                            (kb | k) != 0);
      }
      umma_commit_pair_mc(smem_u32(&bar_empty1[s]), 3);
    }
    umma_commit_pair_mc(smem_u32(&bar_acc1), 3);
    MMF_STAMP(a, 6);

    mbar_wait_cluster(smem_u32(&bar_h), 0);   // both CTAs' H tiles written, GEMM1 TMEM columns drained
    tc_fence_after();
    MMF_STAMP(a, 7);
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < C::NCH; ++c) {
      const int buf = c & 1;
      mbar_wait_cluster(smem_u32(&bar_acc2_empty[buf]), ((c >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < C::KB2; ++kb) {
        mbar_wait(smem_u32(&bar_full2[s]), ph);
        tc_fence_after();
        const uint32_t hs = h_base + kb * 16384;
        const uint32_t bs = ring2 + s * C::STAGE2;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_pair(tmem + buf * C::CHN, umma_desc_sw128(hs + k * 32, 16, 1024),
                            umma_desc_sw128(bs + k * 32, 16, 1024), idesc2, (kb | k) != 0);
        umma_commit_pair_mc(smem_u32(&bar_empty2[s]), 3);
        if (++s == ns2) { s = 0; ph ^= 1u; }
      }
      umma_commit_pair_mc(smem_u32(&bar_acc2_full[buf]), 3);
    }
    if (MODE == AMIL_FWD && a.z_out != nullptr) {
      // z[256, 16] = H_pair · [Wk_hi | Wk_lo]^T into the chunk buffer the last GEMM2 chunk does not use
      constexpr uint32_t idescz = umma_idesc_bf16(256, 16, 0, 0);
      constexpr int c = C::NCH, buf = c & 1;
      mbar_wait_cluster(smem_u32(&bar_acc2_empty[buf]), ((c >> 1) & 1) ^ 1);
      mbar_wait(smem_u32(&bar_full2[s]), ph);
      tc_fence_after();
      for (int kb = 0; kb < C::KB2; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_pair(tmem + buf * C::CHN, umma_desc_sw128(h_base + kb * 16384 + k * 32, 16, 1024),
                            umma_desc_sw128(ring2 + s * C::STAGE2 + kb * 1024 + k * 32, 16, 1024), idescz, (kb | k) != 0);
      }
      umma_commit_pair_mc(smem_u32(&bar_empty2[s]), 3);
      umma_commit_pair_mc(smem_u32(&bar_acc2_full[buf]), 3);
    }
    MMF_STAMP(a, 8);
  } else if (MODE == AMIL_FWD && warp == 3 && (lane == 0 || (MMF_DISCARD_DEAD_STASH == 2 && a.AG != nullptr))) {
#if MMF_DISCARD_DEAD_STASH == 2
    if (a.h_stash != nullptr && a.AG != nullptr) {
      // the otherwise idle lanes of this warp drop the dead lines of THIS tile's H / dG rows (the previous
      // step's; this CTA rewrites exactly these rows) while GEMM1 and EPI1 run. Order: generic-proxy discards, proxy
      // fence, then (a) this warp's own H stores in program order and (b) the epilogue warps' [a|g] stores behind named
      // barrier 5 (they wait on it once, before their first stash store).
      if (row0 < a.N) {
        const long long rows = min((long long)128, a.N - row0);
        uint8_t* hb = a.h_stash + row0 * (L * 2);
        for (long long i = lane; i < rows * (L * 2 / 128); i += 32)
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(hb + i * 128) : "memory");
        uint8_t* gb = reinterpret_cast<uint8_t*>(a.AG) + row0 * (a.ldag * 2);
        for (long long i = lane; i < rows * (a.ldag * 2 / 128); i += 32)
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(gb + i * 128) : "memory");
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      named_bar_arrive(5, AMIL2_EPI_THREADS + 32);
    }
    if (lane == 0)
#endif
    // =============================== H stash store (training forward, each CTA) =========
    // Issued in NCH batches, batch c once GEMM2 chunk c has retired: the 128 KB read of the H tile then shares the
    // shared-memory / TMA pipes with chunks 1.. (hidden behind the gate epilogue) and the pooling tail. Issued right after
    // EPI1 (round 1) it slowed the EXPOSED first chunk from 4.9k to 8.0k cycles; issued in one piece after the last chunk it
    // held up that chunk's stash stores instead (gate epilogue 5.3k -> 7.7k cycles, gpurun_out/r2c / r2d_chunks.log).
    if (a.store_h) {
      for (int c = 0; c < C::NCH; ++c) {
        mbar_wait(smem_u32(&bar_acc2_full[c & 1]), (c >> 1) & 1);   // (=> bar_h completed: both CTAs' H tiles are written and fenced)
        for (int kb = c * C::KB2 / C::NCH; kb < (c + 1) * C::KB2 / C::NCH; ++kb)
          tma_store_2d_keep(&tmH, h_base + kb * 16384, kb * 64, (int)row0);
        tma_store_commit();
      }
      tma_store_wait_exit();
    }
  } else if (warp >= 4) {
    // =============================== epilogue warps (both CTAs) ========================
    const uint32_t q = warp & 3;
    const uint32_t half = (warp - 4) >> 2;      // column half handled by this warp
    const uint32_t r = q * 32 + lane;           // row within the tile == TMEM lane
    const uint32_t e = threadIdx.x - 128;       // 0..255 epilogue thread index
    const long long row = row0 + r;
    // rows of this tile that belong to the bag: the tail of a single bag, or the per-tile count of a packed
    // multi-bag buffer (varlen cohort inference: bags start on 128-row boundaries, padding rows are zero)
    const int valid = (row0 >= a.N) ? 0
                      : (MODE == AMIL_FWD && a.tile_valid != nullptr) ? __ldg(a.tile_valid + tile)
                      : (int)min((long long)128, a.N - row0);
    const bool row_ok = (int)r < valid;
    const uint32_t tq = tmem + ((q * 32u) << 16);
    constexpr bool drop_h = DROPH;
    constexpr bool drop_attn = DROPA;
    const uint32_t rs_h = drop_row_state(a.seed, 0, (uint32_t)row);
    const uint32_t h_ready_leader = mapa_cluster(smem_u32(&bar_h), 0);
    float* sS = vec + C::V_S;     // [2][128] per-half partial row sums (t_i in bwd, score in fwd)
    float* sP = vec + C::V_P;
    float* sRed = vec + C::V_RED;
    // stage the per-column vectors once (epilogue warps only: the producer / MMA threads start at once).
    // The inverted-dropout scale of h is folded into the bias: relu(u + b) * s == relu(s*u + s*b).
    constexpr float h_scale = DROPH ? (1.0f / 0.75f) : 1.0f;
    for (int i = e; i < L; i += AMIL2_EPI_THREADS) {
      vec[C::V_B1 + i] = __ldg(a.b1 + i) * h_scale;
      if (MODE == AMIL_BWD_GATE) vec[C::V_DM + i] = __ldg(a.dM + i);
    }
    // (the sigmoid branch bias is staged pre-halved: sigmoid(z + bb) = 0.5 tanh(0.5 z + 0.5 bb) + 0.5, one FFMA feeds the MUFU)
    for (int i = e; i < C::KD; i += AMIL2_EPI_THREADS) vec[C::V_BAB + i] = __ldg(a.bab + i) * ((GATED && i >= D) ? 0.5f : 1.0f);
    for (int i = e; i < D; i += AMIL2_EPI_THREADS) vec[C::V_WC + i] = __ldg(a.wc + i);
    named_bar_sync(4, AMIL2_EPI_THREADS);
    if (e == 0) MMF_STAMP_N(a, 9);
    if (MODE == AMIL_FWD && a.zero_ptr != nullptr) {
      // fused zero_grad: the epilogue warps have nothing to do until GEMM1 retires (~16k cycles); they clear the
      // step's gradient accumulators (a separate fill kernel costs ~8 us per step with its two kernel boundaries)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long i = (long long)blockIdx.x * AMIL2_EPI_THREADS + e; i < a.zero_n4;
           i += (long long)gridDim.x * AMIL2_EPI_THREADS)
        a.zero_ptr[i] = z;
    }

#if MMF_DISCARD_DEAD_DU
    if (MODE == AMIL_FWD && a.discard_ptr != nullptr) {
      // The previous step's dU (16 MB at the metric shape) is dead — its wgrad has read it, this step's hidden-gradient
      // kernel rewrites every line — but its lines sit DIRTY in L2, and while this kernel streams the bag and allocates the
      // stash they are evicted: 17.5 MB of DRAM write-backs of data nobody will read, inside the forward
      // (profiles/r02d_ncu_step_summary.md §2). discard.global.L2 drops the lines without a write-back.
      for (long long i = (long long)blockIdx.x * AMIL2_EPI_THREADS + e; i < a.discard_n128;
           i += (long long)gridDim.x * AMIL2_EPI_THREADS)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.discard_ptr + i * 128) : "memory");
    }
#endif
#if MMF_DISCARD_DEAD_STASH == 1
    if (MODE == AMIL_FWD && a.h_stash != nullptr && a.AG != nullptr && valid > 0) {
      // the previous step's H and dG rows of THIS tile (dead; this CTA rewrites exactly these rows at its end)
      const long long rows = min((long long)128, a.N - row0);
      uint8_t* hb = a.h_stash + row0 * (L * 2);
      for (long long i = e; i < rows * (L * 2 / 128); i += AMIL2_EPI_THREADS)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(hb + i * 128) : "memory");
      uint8_t* gb = reinterpret_cast<uint8_t*>(a.AG) + row0 * (a.ldag * 2);
      for (long long i = e; i < rows * (a.ldag * 2 / 128); i += AMIL2_EPI_THREADS)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(gb + i * 128) : "memory");
      // (these lines are rewritten later in this kernel by TMA stores = the async proxy: order the generic-proxy discards
      //  before them; the stores are issued behind several CTA barriers by other warps)
      asm volatile("fence.proxy.async;" ::: "memory");
    }
#endif

    // ---------------- EPI1: H = dropout(relu(U + b1)) -> swizzled smem -----------------
    constexpr int PIECES1 = L / 64;  // 32-column pieces per half
    const int cb0 = half * PIECES1;
    // Dropout(0.25) on h: the keep bits of this thread's row and column half (one 32-bit word per piece) are hashed
    // NOW, while GEMM1 runs and the epilogue warps are idle — inside EPI1 the two hashes per piece were a third of its
    // instructions (EPI1: 5.6k cycles without dropout, 8.6k with, gpurun_out/r2c_chunks.log). The words are consumed
    // in order through a register shift (the piece loop is not unrolled).
    uint32_t mw_acc[PIECES1] = {};   // ReLU mask words of this thread's row and column half (training forward)
    static_assert(PIECES1 % 4 == 0, "mask words are stored as 16-byte vectors");
    uint32_t keepw[PIECES1];
#pragma unroll
    for (int ii = 0; ii < PIECES1; ++ii)
      keepw[ii] = DROPH ? (drop_keep_mask16(drop_bits16(rs_h, (uint32_t)((cb0 + ii) * 2))) |
                           (drop_keep_mask16(drop_bits16(rs_h, (uint32_t)((cb0 + ii) * 2 + 1))) << 16))
                        : 0xFFFFFFFFu;
    mbar_wait(smem_u32(&bar_acc1), 0);
    tc_fence_after();
    if (e == 0) MMF_STAMP(a, 10);
    float t_i = 0.f;
    float v[2][32];
    tmem_ld32(tq + cb0 * 32, v[0]);
    // two pieces per iteration so the TMEM double-buffer indices stay static; NOT fully unrolled: the
    // fully unrolled body (8 x ~500 SASS instructions) thrashed the instruction cache
    // (ncu: stall_no_inst 18 % of samples in the backward kernel)
#pragma unroll 1
    for (int i2 = 0; i2 < PIECES1; i2 += 2) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int ii = i2 + par;
        const int cb = cb0 + ii;
        tmem_ld_wait();
        if (ii + 1 < PIECES1) tmem_ld32(tq + (cb + 1) * 32, v[par ^ 1]);
        float (&u)[32] = v[par];
        const float4* b4p = reinterpret_cast<const float4*>(vec + C::V_B1 + cb * 32);
        const uint32_t keep = keepw[par];   // bit i = column cb * 32 + i kept (all ones when dropout is off)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = b4p[i >> 2];
          const float r0 = fmaxf(fmaf(u[i], h_scale, b4.x), 0.f), r1 = fmaxf(fmaf(u[i + 1], h_scale, b4.y), 0.f);
          const float r2 = fmaxf(fmaf(u[i + 2], h_scale, b4.z), 0.f), r3 = fmaxf(fmaf(u[i + 3], h_scale, b4.w), 0.f);
          u[i] = (!DROPH || ((keep >> i) & 1u)) ? r0 : 0.f;
          u[i + 1] = (!DROPH || ((keep >> (i + 1)) & 1u)) ? r1 : 0.f;
          u[i + 2] = (!DROPH || ((keep >> (i + 2)) & 1u)) ? r2 : 0.f;
          u[i + 3] = (!DROPH || ((keep >> (i + 3)) & 1u)) ? r3 : 0.f;
        }
        const uint32_t kb_base = h_base + (cb >> 1) * 16384;
        uint32_t mword = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t p0 = pack_bf16x2(u[8 * j], u[8 * j + 1]), p1 = pack_bf16x2(u[8 * j + 2], u[8 * j + 3]);
          const uint32_t p2 = pack_bf16x2(u[8 * j + 4], u[8 * j + 5]), p3 = pack_bf16x2(u[8 * j + 6], u[8 * j + 7]);
          st_shared_v4(kb_base + sw128_offset(r, (cb & 1) * 4 + j), p0, p1, p2, p3);
          if (MODE == AMIL_FWD) mword |= relu_mask_byte(p0, p1, p2, p3) << (8 * j);
          if (MODE == AMIL_BWD_GATE) {
            const uint32_t pk[4] = {p0, p1, p2, p3};
            const float4* dm4 = reinterpret_cast<const float4*>(vec + C::V_DM + cb * 32 + 8 * j);
            const float4 d0 = dm4[0], d1 = dm4[1];
            const float dmv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int x2 = 0; x2 < 4; ++x2) {
              const float2 hf = unpack_bf16x2(pk[x2]);
              t_i = fmaf(hf.x, dmv[2 * x2], t_i);
              t_i = fmaf(hf.y, dmv[2 * x2 + 1], t_i);
            }
          }
        }
        // training forward: [h > 0] of the bf16 activations the backward will see, 1 bit per element (the dU epilogue
        // of the hidden-gradient kernel reads 64 B per row instead of re-deriving the bits from a re-read of H)
        if (MODE == AMIL_FWD) mw_acc[PIECES1 - 2 + par] = mword;
      }
#pragma unroll
      for (int j = 0; j + 2 < PIECES1; ++j) keepw[j] = keepw[j + 2];
      if (MODE == AMIL_FWD && i2 + 2 < PIECES1) {
#pragma unroll
        for (int j = 0; j + 2 < PIECES1; ++j) mw_acc[j] = mw_acc[j + 2];   // (register shift: the piece loop is not unrolled)
      }
    }
    if (MODE == AMIL_FWD && a.mask_out != nullptr && row_ok) {
      // this thread's PIECES1 consecutive words of the row: 16-byte stores (one word per piece and store instruction was
      // 32 scattered 4-byte writes per warp)
      uint4* mp = reinterpret_cast<uint4*>(a.mask_out + row * (L / 32) + cb0);
#pragma unroll
      for (int j = 0; j < PIECES1 / 4; ++j) mp[j] = make_uint4(mw_acc[4 * j], mw_acc[4 * j + 1], mw_acc[4 * j + 2], mw_acc[4 * j + 3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(h_ready_leader);
    if (e == 0) MMF_STAMP(a, 11);

    if (MODE == AMIL_BWD_GATE) sS[half * 128 + r] = t_i;
    if (MODE == AMIL_BWD_GATE) {
      named_bar_sync(1, AMIL2_EPI_THREADS);   // all H writes of this CTA fenced (+ t_i halves visible)
      if (e == 0) {
        for (int kb = 0; kb < C::KB2; ++kb) tma_store_2d(&tmH, h_base + kb * 16384, kb * 64, (int)row0);
        tma_store_commit();
      }
    }
    float ds = 0.f;
    if (MODE == AMIL_BWD_GATE) {
      t_i = sS[r] + sS[128 + r];
      float dot = 0.f;
      for (int c = lane; c < L; c += 32) dot = fmaf(vec[C::V_DM + c], __ldg(a.M + c), dot);
      dot = warp_sum(dot);
      if (row_ok) {
        const float p = __expf(__ldg(a.A_raw + row) - __ldg(a.ml)) / __ldg(a.ml + 1);
        ds = p * (t_i - dot);
        if (a.dA_raw) ds += __ldg(a.dA_raw + row);
      }
    }

    // ---------------- EPI2: gate + score (fwd) / gate backward (bwd) -------------------
    float s_acc = 0.f;
    constexpr float attn_scale = DROPA ? (1.0f / 0.75f) : 1.0f;
    const uint32_t rs_a = drop_row_state(a.seed, 1, (uint32_t)row);
    const uint32_t rs_g = drop_row_state(a.seed, 2, (uint32_t)row);
#pragma unroll 1
    for (int c = 0; c < C::NCH; ++c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(&bar_acc2_full[buf]), (c >> 1) & 1);
      tc_fence_after();
      if (MMF_TILE2_CHUNK_STAMPS && e == 0 && c < 3) MMF_STAMP(a, 2 + c);
#pragma unroll 1
      for (int pp = 0; pp < 2; ++pp) {
        const int pc = half * 2 + pp;
        const int d0 = c * 128 + pc * 32;
        float va[32], vg[32];
        tmem_ld32(tq + buf * C::CHN + pc * 32, va);
        if (GATED) tmem_ld32(tq + buf * C::CHN + 128 + pc * 32, vg);
        tmem_ld_wait();
        float dwc_v[32];
        uint32_t ab0 = 0xFFFFFFFFu, ab1 = 0xFFFFFFFFu, gb0 = 0xFFFFFFFFu, gb1 = 0xFFFFFFFFu;
        if (drop_attn) {
          ab0 = drop_bits16(rs_a, (uint32_t)(d0 >> 4)); ab1 = drop_bits16(rs_a, (uint32_t)(d0 >> 4) + 1);
          gb0 = drop_bits16(rs_g, (uint32_t)(d0 >> 4)); gb1 = drop_bits16(rs_g, (uint32_t)(d0 >> 4) + 1);
        }
        const float4* ba4p = reinterpret_cast<const float4*>(vec + C::V_BAB + d0);
        const float4* bb4p = reinterpret_cast<const float4*>(vec + C::V_BAB + D + d0);
        const float4* wc4p = reinterpret_cast<const float4*>(vec + C::V_WC + d0);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 ba4 = ba4p[i >> 2], wc4 = wc4p[i >> 2];
          float4 bb4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (GATED) bb4 = bb4p[i >> 2];
          const float bav[4] = {ba4.x, ba4.y, ba4.z, ba4.w};
          const float bbv[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
          const float wcv[4] = {wc4.x, wc4.y, wc4.z, wc4.w};
#pragma unroll
          for (int x2 = 0; x2 < 4; ++x2) {
            const float av = tanh_fast(va[i + x2] + bav[x2]);
            const float gv = GATED ? fmaf(0.5f, tanh_fast(fmaf(0.5f, vg[i + x2], bbv[x2])), 0.5f) : 1.f;
            const uint32_t ab = (i < 16) ? ab0 : ab1, gb = (i < 16) ? gb0 : gb1;
            const float ka = (!DROPA || drop_keep(ab, (i & 15) + x2)) ? attn_scale : 0.f;
            const float kg = (!GATED || !DROPA || drop_keep(gb, (i & 15) + x2)) ? (GATED ? attn_scale : 1.f) : 0.f;
            const float ad = av * ka, gd = gv * kg;
            if (MODE == AMIL_FWD) {
              s_acc = fmaf(wcv[x2], ad * gd, s_acc);
              va[i + x2] = av;               // kept for the optional activation stash
              if (GATED) vg[i + x2] = gv;
            } else {
              const float dq = ds * wcv[x2];
              dwc_v[i + x2] = ds * ad * gd;
              va[i + x2] = dq * gd * ka * (1.f - av * av);
              if (GATED) vg[i + x2] = dq * ad * kg * gv * (1.f - gv);
            }
          }
        }
        if (MODE == AMIL_FWD && a.AG != nullptr) {
          // training forward: stash the pre-dropout branch activations (fp16) so that the backward needs
          // neither GEMM again (amil_hidden_fused.cuh consumes them in place). Each thread owns one ROW of the
          // warp's 32 x 32 block; storing that directly is 32 scattered 16-byte writes per instruction
          // (+10 us per 16k bag).
#if MMF_STASH_TMA
          // The block of each branch is staged in shared memory (row = 64 bytes, 16-byte chunk index XOR (row >> 1) & 3
          // = the TMA's 64-byte swizzle: conflict-free v4 writes) and leaves through ONE TMA store per branch: no read-back,
          // no st.global in the epilogue warps, and the cluster-scope release arrive that ends the chunk has no generic-proxy
          // global stores to drain. The staging is reused one piece (~500 instructions) later: the read-completion wait is free.
          const uint32_t scratch = pool + C::POOL - C::XPOSE_BYTES + (warp - 4) * C::XPOSE_WARP;
#if MMF_DISCARD_DEAD_STASH == 2
          if (c == 0 && pp == 0 && a.h_stash != nullptr) named_bar_sync(5, AMIL2_EPI_THREADS + 32);   // warp 3's discards are fenced
#endif
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int br = 0; br < (GATED ? 2 : 1); ++br) {
            const float (&src)[32] = br == 0 ? va : vg;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(scratch + br * 2048u + lane * 64u + ((j ^ ((lane >> 1) & 3u)) << 4),
                           pack_f16x2(src[8 * j], src[8 * j + 1]), pack_f16x2(src[8 * j + 2], src[8 * j + 3]),
                           pack_f16x2(src[8 * j + 4], src[8 * j + 5]), pack_f16x2(src[8 * j + 6], src[8 * j + 7]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_keep(&tmAGs, scratch, d0, (int)(row0 + q * 32));
            if (GATED) tma_store_2d_keep(&tmAGs, scratch + 2048u, D + d0, (int)(row0 + q * 32));
            tma_store_commit();
          }
#else
          // The block is transposed through a 2 KB per-warp scratch so that every
          // st.global.v4 covers 8 rows x 64 contiguous bytes (full 32-byte sectors).
          const uint32_t scratch = pool + C::POOL - C::XPOSE_BYTES + (warp - 4) * 2048u;
          const uint32_t orow = lane >> 2, ochunk = lane & 3u;
#pragma unroll
          for (int br = 0; br < (GATED ? 2 : 1); ++br) {
            const float (&src)[32] = br == 0 ? va : vg;
            __syncwarp();   // the previous block's reads of the scratch are done
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(scratch + lane * 64u + ((j ^ ((lane >> 1) & 3u)) << 4),
                           pack_f16x2(src[8 * j], src[8 * j + 1]), pack_f16x2(src[8 * j + 2], src[8 * j + 3]),
                           pack_f16x2(src[8 * j + 4], src[8 * j + 5]), pack_f16x2(src[8 * j + 6], src[8 * j + 7]));
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t rr = 8u * t + orow;
              const uint4 v4 = ld_shared_v4(scratch + rr * 64u + ((ochunk ^ ((rr >> 1) & 3u)) << 4));
              const long long grow = row0 + q * 32 + rr;
              if (grow < a.N)
                *reinterpret_cast<uint4*>(a.AG + grow * a.ldag + br * D + d0 + ochunk * 8) = v4;
            }
          }
#endif
        }
        if (MODE == AMIL_BWD_GATE) {
          if (row_ok) {
            uint4* dst_a = reinterpret_cast<uint4*>(a.dG + row * a.lddg + d0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst_a[j] = make_uint4(pack_bf16x2(va[8 * j], va[8 * j + 1]), pack_bf16x2(va[8 * j + 2], va[8 * j + 3]),
                                    pack_bf16x2(va[8 * j + 4], va[8 * j + 5]), pack_bf16x2(va[8 * j + 6], va[8 * j + 7]));
            if (GATED) {
              uint4* dst_g = reinterpret_cast<uint4*>(a.dG + row * a.lddg + D + d0);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst_g[j] = make_uint4(pack_bf16x2(vg[8 * j], vg[8 * j + 1]), pack_bf16x2(vg[8 * j + 2], vg[8 * j + 3]),
                                      pack_bf16x2(vg[8 * j + 4], vg[8 * j + 5]), pack_bf16x2(vg[8 * j + 6], vg[8 * j + 7]));
            }
          }
          float* wsrow = a.colsum_ws + ((long long)tile * 4 + q) * C::NCOLS;
          const float s0 = warp_colsum32(dwc_v);
          wsrow[d0 + lane] = s0;
          const float s1 = warp_colsum32(va);
          wsrow[D + d0 + lane] = s1;
          if (GATED) {
            const float s2 = warp_colsum32(vg);
            wsrow[2 * D + d0 + lane] = s2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&bar_acc2_empty[buf]), 0));
      if (MMF_TILE2_CHUNK_STAMPS && e == 0 && c < 3) MMF_STAMP(a, c == 0 ? 5 : c == 1 ? 9 : 15);
    }

    if (MODE == AMIL_FWD && a.z_out != nullptr) {
      constexpr int c = C::NCH, buf = c & 1;
      mbar_wait(smem_u32(&bar_acc2_full[buf]), (c >> 1) & 1);
      tc_fence_after();
      if (half == 0) {
        float zv[16];
        tmem_ld16(tq + buf * C::CHN, zv);
        tmem_ld_wait();
        if (row_ok) {   // columns 0..7 = Wk_hi · h, 8..15 = Wk_lo · h
          float4* zp = reinterpret_cast<float4*>(a.z_out + row * a.zld);
          zp[0] = make_float4(zv[0] + zv[8], zv[1] + zv[9], zv[2] + zv[10], zv[3] + zv[11]);
          if (a.zld == 8) zp[1] = make_float4(zv[4] + zv[12], zv[5] + zv[13], zv[6] + zv[14], zv[7] + zv[15]);
        }
      }
      tc_fence_before();
    }
    if (e == 0) MMF_STAMP(a, 12);
    if (MODE == AMIL_BWD_GATE) {
      if (half == 0) {
        const float dsum = warp_sum(ds);
        if (lane == 0) a.dbc_ws[(long long)tile * 4 + q] = dsum;
      }
      if (e == 0) tma_store_wait_exit();
    } else {
      // ---------------- FWD: scores out + tile softmax partial ------------------------
      // head row of the tile (training step): this thread's two classifier columns, requested before the softmax barriers
      float2 hwk[HEAD_MAX_K];
      const bool head_row = a.tile_head != nullptr;
#pragma unroll
      for (int k = 0; k < HEAD_MAX_K; ++k)
        hwk[k] = (head_row && k < a.head_k && e < (uint32_t)L / 2)
                     ? __ldg(reinterpret_cast<const float2*>(a.head_wk + (long long)k * L + 2 * e)) : make_float2(0.f, 0.f);
      float m_keep = 0.f, l_keep = 0.f;
      sS[half * 128 + r] = s_acc;
      named_bar_sync(2, AMIL2_EPI_THREADS);
      const bool tile_ok = valid > 0 || (a.tile_valid != nullptr && row0 < a.N);   // (empty varlen tiles still write m = -inf)
      if (half == 0) {
        const float s = row_ok ? sS[r] + sS[128 + r] + __ldg(a.bc) : -INFINITY;
        if (row_ok) a.A_raw[row] = s;
        const float wm = warp_max(s);
        if (lane == 0) sRed[q] = wm;
        named_bar_sync(3, 128);
        const float m_t = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
        const float p = row_ok ? __expf(s - m_t) : 0.f;
        sP[r] = p;
        const float wsum = warp_sum(p);
        if (lane == 0) sRed[4 + q] = wsum;
        named_bar_sync(3, 128);
        m_keep = m_t; l_keep = sRed[4] + sRed[5] + sRed[6] + sRed[7];
        if (r == 0 && tile_ok) {
          float* prow = a.partials + (long long)tile * (L + 2);
          prow[0] = m_t;
          prow[1] = l_keep;
        }
      }
      named_bar_sync(2, AMIL2_EPI_THREADS);   // sP complete
      if (tile_ok) {
        float* prow = a.partials + (long long)tile * (L + 2);
        for (uint32_t cp = e; cp < (uint32_t)L / 2; cp += AMIL2_EPI_THREADS) {
          const uint32_t col = 2u * cp;
          const uint32_t kb = col >> 6, chunk = (col & 63u) >> 3, inb = (col & 7u) * 2u;
          const uint32_t blk = h_base + kb * 16384u;
          // 8 rows per step: two 16-byte loads of p (broadcast), eight 4-byte loads of the bf16 pairs (row r of the
          // 128B-swizzled block: r * 128 + ((chunk ^ (r & 7)) << 4)), two independent accumulator pairs
          float2 acc_e = make_float2(0.f, 0.f), acc_o = make_float2(0.f, 0.f);
#pragma unroll 2
          for (uint32_t r8 = 0; r8 < 128; r8 += 8) {
            const float4 pa = *reinterpret_cast<const float4*>(sP + r8), pb = *reinterpret_cast<const float4*>(sP + r8 + 4);
            const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
            uint32_t hw[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) hw[j] = ld_shared_b32(blk + (r8 + j) * 128u + ((chunk ^ j) << 4) + inb);
#pragma unroll
            for (uint32_t j = 0; j < 8; j += 2) {
              const float2 he = unpack_bf16x2(hw[j]), ho = unpack_bf16x2(hw[j + 1]);
              acc_e.x = fmaf(pv[j], he.x, acc_e.x); acc_e.y = fmaf(pv[j], he.y, acc_e.y);
              acc_o.x = fmaf(pv[j + 1], ho.x, acc_o.x); acc_o.y = fmaf(pv[j + 1], ho.y, acc_o.y);
            }
          }
          const float c0 = acc_e.x + acc_o.x, c1 = acc_e.y + acc_o.y;
          prow[2 + col] = c0;
          prow[2 + col + 1] = c1;
#pragma unroll
          for (int k = 0; k < HEAD_MAX_K; ++k) hwk[k].x = fmaf(hwk[k].x, c0, hwk[k].y * c1);   // (L / 2 <= 256: one pass)
        }
      }
      static_assert(L / 2 <= (int)AMIL2_EPI_THREADS, "one column pair per epilogue thread");
      if (head_row) {
        // Wk · acc_t: the head of the backward then merges 12 floats per tile instead of L + 2
        // (logits - bk = sum_t e^{m_t - m} (Wk·acc_t) / l, exact fp32 classifier weights)
        float* sZ = sS;   // (the score halves are dead)
#pragma unroll
        for (int k = 0; k < HEAD_MAX_K; ++k) {
          const float v = warp_sum((tile_ok && e < (uint32_t)L / 2) ? hwk[k].x : 0.f);
          if (lane == 0) sZ[(warp - 4) * HEAD_MAX_K + k] = v;
        }
        named_bar_sync(2, AMIL2_EPI_THREADS);
        if (tile_ok && e < 12u) {
          float v = 0.f;
          if (e >= 4u) {
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sZ[w * HEAD_MAX_K + (e - 4u)];
          } else if (e == 0u) v = m_keep;
          else if (e == 1u) v = l_keep;
          a.tile_head[(long long)tile * 12 + e] = v;
        }
      }
      if (MMF_STASH_TMA && a.AG != nullptr && lane == 0) tma_store_wait_exit();   // this warp's stash stores
    }
    if (e == 0) MMF_STAMP(a, 13);
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's smem / TMEM through the pair MMA
  if (threadIdx.x == 0) MMF_STAMP(a, 14);
  timeline_end(MODE == AMIL_FWD ? 0 : 4, tl);
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}

}  // namespace mmf
