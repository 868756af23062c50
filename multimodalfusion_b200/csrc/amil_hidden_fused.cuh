// amil_hidden_fused.cuh — stash-mode backward, stages 1 + 2 in ONE kernel (sm_100a, CTA pair).
//
// Fuses the gate backward (an elementwise pass over the stashed activations) into the hidden-gradient GEMM as the PRODUCER of its A
// operand, so dG makes no round trip through L2 between the two and the step loses one launch:
//
//   phase A  (worker warps, while the producer prefetches Wab): per row i of the 128-row tile
//            p_i = e^{s_i-m}/l, ds_i = p_i (t_i - dM·M) + dA_i with
//              HEADPROJ: t_i = dlogits·z_i, z_i = Wk h_i emitted by the forward's tensor cores (K FMAs per row;
//                        a linear classifier sits directly on M: dM = Wk^T dlogits — path / radio models), or
//              general : t_i = dM·h_i from coalesced warp-per-row loads of the stashed H tile (dM from a fusion head);
//            the ReLU mask [h_i > 0] comes as 1 bit per element from the training forward (mask words);
//   mainloop per 64-wide slice c of D ("k-pair"): TMA brings the stashed fp16 tiles a[:, c], g[:, c] into
//            an A stage; the epilogue warps rewrite them IN PLACE as the bf16 tiles
//            dG_a = ds wc g (1-a²), dG_g = ds wc a g (1-g)  (column-stationary: thread = 4 columns x 8
//            rows, so dwc / dba / dbb sums stay in registers), publish them to the MMA warp, and one
//            thread TMA-stores them to the dG buffer (the wgrad GEMM needs dG);
//            tcgen05.mma.cta_group::2 (M = 256, N = 256 x L/256):  acc += dG_a Wa[c] + dG_g Wb[c];
//   epilogue dU = (acc + p_i dM) ⊙ [h > 0] · 1/(1-p) -> bf16 -> swizzled staging -> TMA store; db1 column
//            sums from the staged tile (as gemm2_tc.cuh EPI_DU, mask words now from shared memory).
//
// Math: SURVEY.md App. A.2 = backward of models/model_modules.py:105-110 (gated attention) and
// models/model_attention_mil_path.py:20-21,29,53-56 (fc + ReLU + dropout, softmax pooling).
//
// Shared memory (dynamic, 1024-aligned pool):
//   B ring  NSB stages x [64 k-rows][L/2 cols] (this CTA's half of a Wab k-block, MN-major boxes)
//   A ring  NSA stages x (a tile 16 KB + g tile 16 KB)   [128 rows][64 cols], K-major, 128B swizzle
//   after the mainloop both rings are dead and hold the staged dU tile (128 x L bf16)
//   vectors: dM[L], wc[D], ds[128], p[128], mask[128][L/32], colsum scratch [3D]
// Barriers: b_full[s] leader copy (pair TMA), b_empty[s] / a_empty[s] / acc per CTA (multicast commits;
// a_empty also counts the dG store's read-completion), a_full[s] per CTA (local TMA), a_ready[s] leader
// copy, one arrive per CTA relayed by its dG-store thread (MMF_HIDDEN_RELAY) or one per worker warp of either CTA.
#pragma once
#include <cuda_fp16.h>

#include "amil_tile.cuh"
#include "gemm_tc.cuh"

namespace mmf {

struct HiddenFusedArgs {
  long long N;
  const __nv_bfloat16* H;   // stash [N, L]
  const float* A_raw;       // [N]
  const float* ml;          // (m, l)
  const float* M;           // [L]  (general phase A only)
  const float* dM;          // [L]
  const uint32_t* mask;     // [N, L/32] ReLU mask words written by the training forward
  const float* z;           // HEADPROJ: [N, zld] = Wk h_i
  int zld;                  // 4 or 8
  const float* partials;    // HEADPROJ: the forward's tile partials [n_tiles, L + 2] (pooled embedding M, off the critical path)
  int n_tiles;              //           (<= HEAD_MAX_TILES)
  const float* tile_head;   // HEADPROJ: the forward's head rows [n_tiles, 12] = (m_t, l_t, -, -, Wk·acc_t [8])
  HeadTail head;            // HEADPROJ: the step's head, run in this kernel's prologue (amil_head_tail.cuh)
  const float* dA_raw;      // [N] or null
  const float* wc;          // [D]
  float* dwc;               // [D]  accumulated
  float* dbab;              // [KD] accumulated
  float* dbc;               // [1]  accumulated
  float* db1;               // [L]  accumulated
  float du_scale;           // 1 or 1/(1-p)
  // varlen window (general phase A only): the rows are a packed buffer of several bags that start on 128-row boundaries
  const int* tile_bag;      // [tiles] bag of every 128-row tile: M / dM are [bags, L], ml is [bags, 2]
  const int* tile_valid;    // [tiles] rows of the tile that belong to its bag (the rest is zero padding)
  unsigned long long seed;
  unsigned long long* dbg;
};

template <int L, int D, bool GATED>
struct HiddenFusedCfg {
  static constexpr int KD = GATED ? 2 * D : D;
  static constexpr int NKP = D / 64;                       // k-pairs (64-wide slices of D)
  static constexpr int NH = L / 256;                       // N = 256 MMAs per k-step
  static constexpr uint32_t B_STAGE = NH * 16384u;         // this CTA's half of one Wab k-block
  static constexpr uint32_t A_STAGE = GATED ? 32768u : 16384u;
  // Ring depths (both stages are 32 KB at L = 512, gated). The mainloop is a dependency cycle per A stage — tile
  // landed -> transform (until the slowest of 32 worker warps has arrived, ~3k cycles) -> MMAs (~2k) -> commit ->
  // TMA reload (~2k) — of ~7.3k cycles; with 2 A stages a slice took 3.7-4.4k cycles (measured, tools/phase_gemm.py)
  // against ~2k of MMA. 3 A stages bring the cycle under the transform / MMA time; the L2-hot Wab stream needs
  // less cover (one stage is consumed every ~1.2k cycles, reload ~1k), so it gives up one stage.
#ifndef MMF_HIDDEN_NSB
#define MMF_HIDDEN_NSB 3
#endif
#ifndef MMF_HIDDEN_NSA
#define MMF_HIDDEN_NSA 3
#endif
  static constexpr int NSB = MMF_HIDDEN_NSB;
  static constexpr int NSA = MMF_HIDDEN_NSA;
  static constexpr uint32_t RING_BYTES = NSB * B_STAGE + NSA * A_STAGE;
  static constexpr uint32_t STAGING = 128u * L * 2u;       // dU tile
  static constexpr uint32_t POOL = RING_BYTES > STAGING ? RING_BYTES : STAGING;
  // vector region (floats): dM | wc | ds | p | head scratch: merged accumulators of the row groups [RG][L],
  // their (m, l) [RG][2] (8 floats), per-warp partial logits [16][8]
  static constexpr int V_DM = 0, V_WC = L, V_DS = L + D, V_P = V_DS + 128, V_HACC = V_P + 128,
                       V_HML = V_HACC + 1024, V_HRED = V_HML + 8, V_END = V_HRED + 128;
  static constexpr uint32_t VEC_BYTES = ((V_END * 4u + 1023u) / 1024u) * 1024u;
  static constexpr uint32_t SMEM_BYTES = POOL + VEC_BYTES + 1024u;
};

// MMF_HIDDEN_RELAY = 1: the worker warps publish a transformed stage with ONE local arrive (bar_astore); the CTA's
// dG-store thread, which waits on that barrier anyway, relays a single cluster-scope arrive to the leader's bar_aready.
// = 0: every worker warp of either CTA arrives on the leader's barrier itself (32 release.cluster arrives per slice,
// each draining the warp's shared-memory writes at cluster scope on the transform's critical path).
#ifndef MMF_HIDDEN_RELAY
#define MMF_HIDDEN_RELAY 1
#endif
#ifndef MMF_HIDDEN_PACKED
#define MMF_HIDDEN_PACKED 1   // transform in packed f32x2 arithmetic + halving-butterfly column sums (0: the scalar round-1 form)
#endif
#ifndef MMF_HEAD_STAMPS
#define MMF_HEAD_STAMPS 0    // 1: stamps 8-13 time the head prologue instead of the first three slices
#endif
#define MMF_HS(i) do { if (MMF_HEAD_STAMPS && e == 0) MMF_STAMP(a, i); } while (0)
constexpr int HIDDEN_EW = 16;                         // worker (phase A / transform / epilogue) warps: the CUDA-core
                                                     // phases are latency-bound, 16 warps hide ~2x what 8 did
constexpr int HIDDEN_ET = HIDDEN_EW * 32;             // worker threads
constexpr int HIDDEN_THREADS = 128 + HIDDEN_ET;

template <int L, int D, bool GATED, bool DROP, bool HEADPROJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HIDDEN_THREADS, 1)
amil_hidden_fused_kernel(const __grid_constant__ CUtensorMap tmAG,   // stash fp16 [N, KD], box [128][64]
                         const __grid_constant__ CUtensorMap tmDG,   // same memory viewed as bf16 dG (store)
                         const __grid_constant__ CUtensorMap tmWab,  // bf16 [KD, L], box [64][64]
                         const __grid_constant__ CUtensorMap tmDU,   // bf16 [N, L], box [128][64] (store)
                         const HiddenFusedArgs a_in) {
  using C = HiddenFusedCfg<L, D, GATED>;
  HiddenFusedArgs a = a_in;   // (the pointers to the previous kernel's outputs are re-derived after griddepcontrol.wait)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_bfull[C::NSB], bar_bempty[C::NSB];
  __shared__ __align__(8) uint64_t bar_afull[C::NSA], bar_aready[C::NSA], bar_aempty[C::NSA];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ __align__(8) uint64_t bar_astore[C::NSA];   // local: one arrive per epilogue warp once a stage holds dG
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_dbc;
  __shared__ float s_dl_pub[HEAD_MAX_K];          // HEADPROJ: dlogits for the warp that forms M / dWk (valid once s_dl_flag != 0)
  __shared__ volatile uint32_t s_dl_flag;
  // per-warp private column-sum slots (no atomics: shared-memory fp32 atomicAdd is a CAS loop): warp (column group
  // ctg = w & 3, row quarter rq = w >> 2) owns, per slice, 3 sums (dwc | dba | dbb) x 16 columns
  __shared__ float s_part[HIDDEN_EW][C::NKP * 48];

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pool = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* pool_ptr = smem_raw + (pool - smem_u32(smem_raw));
  const uint32_t b_ring = pool, a_ring = pool + C::NSB * C::B_STAGE;
  float* vec = reinterpret_cast<float*>(pool_ptr + C::POOL);
  const int pair_m = blockIdx.x >> 1;
  const long long row0 = (long long)pair_m * 256 + 128 * (int)rank;   // first row of this CTA

  Timeline tl = timeline_start(2);
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    MMF_STAMP(a, 0);
    for (int s = 0; s < C::NSB; ++s) { mbar_init(smem_u32(&bar_bfull[s]), 1); mbar_init(smem_u32(&bar_bempty[s]), 1); }
    for (int s = 0; s < C::NSA; ++s) {
      mbar_init(smem_u32(&bar_afull[s]), 1);
      mbar_init(smem_u32(&bar_aready[s]), MMF_HIDDEN_RELAY ? 2 : 2 * HIDDEN_EW);
      mbar_init(smem_u32(&bar_aempty[s]), 2);   // MMA commit + the dG store's read-completion
      mbar_init(smem_u32(&bar_astore[s]), HIDDEN_EW);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmAG); tma_prefetch_desc(&tmWab); tma_prefetch_desc(&tmDG); tma_prefetch_desc(&tmDU);
    s_dbc = 0.f;
    s_dl_flag = 0u;
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(&tmem_base_slot), L);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  griddep_wait();
  a.H = pdl_fresh(a.H); a.A_raw = pdl_fresh(a.A_raw); a.ml = pdl_fresh(a.ml); a.M = pdl_fresh(a.M); a.dM = pdl_fresh(a.dM);
  a.mask = pdl_fresh(a.mask); a.z = pdl_fresh(a.z); a.partials = pdl_fresh(a.partials); a.tile_head = pdl_fresh(a.tile_head);
  a.dA_raw = pdl_fresh(a.dA_raw);
  if (DROP) a.seed = seed_resolve(a.seed);   // (device-resident seed of a graph-captured step)
  if (!HEADPROJ && a.tile_bag != nullptr && row0 < a.N) {
    // packed window of bags: this CTA's tile belongs to ONE bag — its softmax statistics, pooled embedding and dM — and
    // only its first tile_valid rows exist (padding rows get p = ds = 0 and a zero mask: dG = dU = 0, nothing reaches
    // the weight gradients)
    const long long tile = row0 >> 7;
    const long long bag = pdl_fresh(a.tile_bag)[tile];
    a.M += bag * L; a.dM += bag * L; a.ml += 2 * bag;
    a.N = row0 + pdl_fresh(a.tile_valid)[tile];
  }
  timeline_wait_done(tl);
  if (threadIdx.x == 0) MMF_STAMP(a, 1);

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer, B stream (both CTAs) =================
    // Wab k-block order = [a rows of slice 0, g rows of slice 0, a rows of slice 1, ...]. The A stream has its
    // own thread (warp 3): a wait for a free B slot must never delay the next slice's activation tiles.
    constexpr int NKB = C::NKP * (GATED ? 2 : 1);
    for (int i = 0; i < NKB; ++i) {
      const int s = i % C::NSB;
      mbar_wait(smem_u32(&bar_bempty[s]), ((i / C::NSB) & 1) ^ 1);
      const uint32_t full = smem_u32(&bar_bfull[s]);
      if (leader) mbar_arrive_expect_tx(full, 2 * C::B_STAGE);
      const int kp = GATED ? (i >> 1) : i;
      const int krow = (GATED && (i & 1)) ? D + kp * 64 : kp * 64;   // row of Wab [KD, L]
#pragma unroll
      for (int h = 0; h < C::NH; ++h)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_2d_pair(b_ring + s * C::B_STAGE + h * 16384 + j * 8192, &tmWab, full,
                           256 * h + 128 * (int)rank + j * 64, krow);
    }
  } else if (warp == 3 && lane == 0) {
    // =============================== TMA producer, A stream (each CTA its own rows) =====
    for (int kp = 0; kp < C::NKP; ++kp) {
      const int s = kp % C::NSA;
      mbar_wait(smem_u32(&bar_aempty[s]), ((kp / C::NSA) & 1) ^ 1);
      const uint32_t full = smem_u32(&bar_afull[s]);
      mbar_arrive_expect_tx(full, C::A_STAGE);
      tma_load_2d(a_ring + s * C::A_STAGE, &tmAG, full, kp * 64, (int)row0);
      if (GATED) tma_load_2d(a_ring + s * C::A_STAGE + 16384, &tmAG, full, D + kp * 64, (int)row0);
    }
  } else if (warp == 2 && lane == 0) {
    // =============================== dG store thread (each CTA) =========================
    // dG -> global for the wgrad GEMM, straight from the transformed A stage (same swizzled tile layout)
    const uint32_t a_ready_leader = mapa_cluster(smem_u32(&bar_aready[0]), 0);
    for (int kp = 0; kp < C::NKP; ++kp) {
      const int s = kp % C::NSA;
      mbar_wait(smem_u32(&bar_astore[s]), (kp / C::NSA) & 1);
      if (MMF_HIDDEN_RELAY) mbar_arrive_cluster(a_ready_leader + s * 8u);   // this CTA's half of the stage is ready
      tma_store_2d_keep(&tmDG, a_ring + s * C::A_STAGE, kp * 64, (int)row0);
      if (GATED) tma_store_2d_keep(&tmDG, a_ring + s * C::A_STAGE + 16384, D + kp * 64, (int)row0);
      tma_store_commit();
      tma_store_wait_read();                       // the stage may be overwritten once the store has read it
      mbar_arrive(smem_u32(&bar_aempty[s]));
    }
    tma_store_wait_exit();
  } else if (warp == 1 && lane == 0 && leader) {
    // =============================== MMA issuer (leader CTA) ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 1);
    int bi = 0;
    for (int kp = 0; kp < C::NKP; ++kp) {
      const int sa = kp % C::NSA;
      mbar_wait_cluster(smem_u32(&bar_aready[sa]), (kp / C::NSA) & 1);
      tc_fence_after();
      if (kp == 0) MMF_STAMP(a, 2);
#pragma unroll
      for (int br = 0; br < (GATED ? 2 : 1); ++br, ++bi) {
        const int sb = bi % C::NSB;
        mbar_wait(smem_u32(&bar_bfull[sb]), (bi / C::NSB) & 1);
        tc_fence_after();
        const uint32_t a_src = a_ring + sa * C::A_STAGE + br * 16384;
        const uint32_t b_src = b_ring + sb * C::B_STAGE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = umma_desc_sw128(a_src + k * 32, 16, 1024);
#pragma unroll
          for (int h = 0; h < C::NH; ++h)
            umma_bf16_ss_pair(tmem + h * 256, ad, umma_desc_sw128(b_src + h * 16384 + k * 2048, 8192, 1024), idesc,
                              (kp | br | k) != 0);
        }
        umma_commit_pair_mc(smem_u32(&bar_bempty[sb]), 3);
      }
      umma_commit_pair_mc(smem_u32(&bar_aempty[sa]), 3);
    }
    umma_commit_pair_mc(smem_u32(&bar_acc), 3);
    MMF_STAMP(a, 3);
  } else if (HEADPROJ && warp == 1 && !leader) {
    // =============================== pooled embedding M and dWk (idle warp of the non-leader CTAs) =====
    // The step's head no longer needs M (it merges Wk·acc_t per tile); M itself is an output and feeds dWk += dlogits (x) M.
    // Pair p forms the 8-column groups p, p + pairs, ...: lane = (row group lane >> 3, column lane & 7), tiles strided by 4.
    const HeadTail& h = a.head;
    const int pairs = (int)(gridDim.x >> 1);
    float m = -CUDART_INF_F;
    for (int t = (int)lane; t < a.n_tiles; t += 32) m = fmaxf(m, __ldg(a.tile_head + (long long)t * 12));
    m = warp_max(m);
    float l = 0.f;
    for (int t = (int)lane; t < a.n_tiles; t += 32) {
      const float2 mlv = __ldg(reinterpret_cast<const float2*>(a.tile_head + (long long)t * 12));
      l = fmaf(mlv.y, (mlv.x > -CUDART_INF_F) ? __expf(mlv.x - m) : 0.f, l);
    }
    l = warp_sum(l);
    const float inv_l = 1.f / l;
    const int c8 = (int)(lane & 7u), rgp = (int)(lane >> 3);
    bool have_dl = false;
    float dlv[HEAD_MAX_K];
    for (int cg = pair_m; cg < L / 8; cg += pairs) {
      const int col = cg * 8 + c8;
      float acc = 0.f;
#pragma unroll 1
      for (int t0 = rgp; t0 < a.n_tiles; t0 += 32) {
        float mt[8], v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + 4 * u;
          const bool ok = t < a.n_tiles;
          mt[u] = ok ? __ldg(a.partials + (long long)t * (L + 2)) : -CUDART_INF_F;
          v[u] = ok ? __ldg(a.partials + (long long)t * (L + 2) + 2 + col) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(v[u], (mt[u] > -CUDART_INF_F) ? __expf(mt[u] - m) : 0.f, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      const float Mc = acc * inv_l;
      if (!have_dl) {   // the workers of this CTA publish dlogits after their head; long done by now
        uint32_t spins = 0;
        while (s_dl_flag == 0u) {
          if (++spins > (1u << 26)) { printf("mmf: dlogits never published (block %d)\n", (int)blockIdx.x); __trap(); }
        }
        __threadfence_block();
#pragma unroll
        for (int j = 0; j < HEAD_MAX_K; ++j) dlv[j] = s_dl_pub[j];
        have_dl = true;
      }
      if (rgp == 0) {
        h.M[col] = Mc;
        if (h.dWk) {
#pragma unroll
          for (int j = 0; j < HEAD_MAX_K; ++j)
            if (j < h.K) { float* dw = h.dWk + (long long)j * L + col; *dw = fmaf(dlv[j], Mc, *dw); }
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue / transform warps ========================
    const uint32_t q = warp & 3;
    const uint32_t part = (warp - 4) >> 2;         // column part handled in the epilogue (L / (EW/4) columns)
    const uint32_t e = threadIdx.x - 128;          // 0..HIDDEN_ET-1
    const uint32_t ew = warp - 4;                  // 0..EW-1
    constexpr int ROWS_PER_WARP = 128 / HIDDEN_EW;
    if (!HEADPROJ) {
      for (int i = e; i < L; i += HIDDEN_ET) vec[C::V_DM + i] = __ldg(a.dM + i);
      for (int i = e; i < D; i += HIDDEN_ET) vec[C::V_WC + i] = __ldg(a.wc + i);
    }
    // (HEADPROJ: wc is staged after the head's loads have been requested — the head is a chain of L2 round trips on
    //  the kernel's critical path, nothing may sit in front of its first request)


    // ---------------- phase A: ds_i, p_i -----------------------------------------------------------------
    if (HEADPROJ) {
      // ---- the step's head, redundantly in every CTA (amil_head_tail.cuh); CTA 0 writes its outputs ----------
      const HeadTail& h = a.head;
      constexpr int CP = L / 2, RG = HIDDEN_ET / CP;     // column pairs; row groups (2 at L = 512, 4 at L = 256)
      static_assert(RG * L <= 1024 && RG <= 4, "head scratch");
      const int cp = (int)e % CP, rg = (int)e / CP;
      const int K = h.K;
      // every global input of the prologue is requested before the first use: one L2 round trip for all of them
      float2 wk[HEAD_MAX_K];
#pragma unroll
      for (int j = 0; j < HEAD_MAX_K; ++j)
        wk[j] = (j < K) ? __ldg(reinterpret_cast<const float2*>(h.Wk + (long long)j * L + 2 * cp)) : make_float2(0.f, 0.f);
      const float bk_l = ((int)lane < K) ? __ldg(h.bk + lane) : 0.f;   // lane j <-> class j in the scalar section
      const int y = (int)__ldg(h.Y);
      const float cb = __ldg(h.c);
      const long long prow = row0 + e;               // phase-A row of threads e < 128
      const bool prow_ok = e < 128 && prow < a.N;
      float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
      float s_raw = 0.f, dA = 0.f;
      if (prow_ok) {
        z0 = __ldg(reinterpret_cast<const float4*>(a.z + prow * a.zld));
        if (a.zld == 8) z1 = __ldg(reinterpret_cast<const float4*>(a.z + prow * a.zld + 4));
        s_raw = __ldg(a.A_raw + prow);
        if (a.dA_raw) dA = __ldg(a.dA_raw + prow);
      }
      float wc_stage = 0.f;
      if ((int)e < D) wc_stage = __ldg(a.wc + e);
      static_assert(D <= HIDDEN_ET, "wc staged one element per thread");
      float* s_hacc = vec + C::V_HACC;
      float* s_hred = vec + C::V_HRED;
      // Merge of the forward's per-tile head rows (m_t, l_t, Wk·acc_t): ONE 48-byte row per thread, one L2 round trip,
      // three CTA barriers. (Round 2a merged the (L + 2)-float partials in two levels with group flags: 14.5k cycles of
      // the kernel's 45k sat in front of the first MMA, gpurun_out/r2_phase9.log; the pooled embedding itself is only
      // an output and an operand of dWk: the idle warp of the non-leader CTAs forms it off the critical path.)
      float4 th0 = make_float4(-CUDART_INF_F, 0.f, 0.f, 0.f), th1 = make_float4(0.f, 0.f, 0.f, 0.f), th2 = th1;
      if ((int)e < a.n_tiles) {
        const float4* th = reinterpret_cast<const float4*>(a.tile_head + (long long)e * 12);
        th0 = __ldg(th); th1 = __ldg(th + 1); th2 = __ldg(th + 2);
      }
      static_assert(HEAD_MAX_TILES <= HIDDEN_ET, "one head row per worker thread");
      MMF_HS(8);
      {
        const float wm = warp_max(th0.x);
        if (lane == 0) s_hred[ew] = wm;
      }
      named_bar_sync(1, HIDDEN_ET);
      float m = s_hred[0];
#pragma unroll
      for (int w = 1; w < HIDDEN_EW; ++w) m = fmaxf(m, s_hred[w]);
      MMF_HS(9);
      {
        const float wt = (th0.x > -CUDART_INF_F) ? __expf(th0.x - m) : 0.f;
        float vals[9] = {th0.y * wt, th1.x * wt, th1.y * wt, th1.z * wt, th1.w * wt, th2.x * wt, th2.y * wt, th2.z * wt, th2.w * wt};
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const float v = warp_sum(vals[j]);
          if (lane == 0) s_hacc[ew * 16 + j] = v;
        }
      }
      named_bar_sync(1, HIDDEN_ET);
      MMF_HS(10);
      if ((int)e < D) vec[C::V_WC + e] = wc_stage;
      // hazards, survival function, nll_surv and its gradient (closed form, amil_head_tail.cuh) by ONE warp, lane j =
      // class j; the K dlogits, dM·M and l go to the other warps through shared memory
      float* s_dl = s_hacc + 512;   // [HEAD_MAX_K] dlogits | dM·M | l
      if (ew == 0) {
        const int j = (int)lane;
        float tot = 0.f;           // lane 0: l, lane 1 + k: sum_t w_t (Wk·acc_t)_k
        if (j < 9)
#pragma unroll
          for (int w = 0; w < HIDDEN_EW; ++w) tot += s_hacc[w * 16 + j];
        const float l = __shfl_sync(0xffffffffu, tot, 0);
        const float zsum = __shfl_sync(0xffffffffu, tot, (j + 1) & 31);
        const float lgm = (j < K) ? zsum / l : 0.f;     // logit_j - bk_j
        const float lg = lgm + bk_l;
        const float hz = (j < K) ? 1.f / (1.f + expf(-lg)) : 0.f;
        float sv = 1.f - hz;             // inclusive product scan over the classes: S(j)
#pragma unroll
        for (int o = 1; o < HEAD_MAX_K; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, sv, o);
          if (j >= o) sv *= t;
        }
        const float sp_y = (y > 0) ? __shfl_sync(0xffffffffu, sv, y > 0 ? y - 1 : 0) : 1.f;   // S(y-1)
        const float h_y = __shfl_sync(0xffffffffu, hz, y), sp_y1 = __shfl_sync(0xffffffffu, sv, y);
        const float alpha = h.alpha, eps = h.eps;
        const float c1 = (y > 0 && sp_y >= eps) ? (1.f - cb) : 0.f;          // h_j, j <= y-1
        const float c2 = (h_y >= eps) ? (1.f - cb) : 0.f;                    // -(1 - h_y), j == y
        const float c3 = (sp_y1 >= eps) ? (1.f - alpha) * cb : 0.f;          // h_j, j <= y
        float dl = hz * ((j <= y - 1 ? c1 : 0.f) + (j <= y ? c3 : 0.f)) - (j == y ? c2 * (1.f - hz) : 0.f);
        dl = (j < K) ? dl * h.loss_scale : 0.f;
        const float dot = warp_sum(dl * lgm);          // dM·M = sum_j dlogit_j (logit_j - bk_j)
        float bv = (j < K) ? lg : -CUDART_INF_F;       // argmax, first maximum (torch.topk on the logits)
        int bi = j;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (j < HEAD_MAX_K) { s_dl[j] = dl; s_dl_pub[j] = dl; }
        if (j == 0) { s_dl[HEAD_MAX_K] = dot; s_dl[HEAD_MAX_K + 1] = l; }
        __threadfence_block();
        __syncwarp();
        if (j == 0) s_dl_flag = 1u;      // the M / dWk warp (warp 1 of the non-leader CTAs) may read s_dl_pub
        if (blockIdx.x == 0) {
          if (j < K) {
            h.hazards[j] = hz; h.S[j] = sv;
            if (h.dbk) h.dbk[j] += dl;
          }
          if (j < HEAD_MAX_K) h.hs[j] = dl;
          if (j == 0) {
            h.ml[0] = m; h.ml[1] = l;
            h.hs[HS_DOT] = dot;
            if (h.Y_hat) *h.Y_hat = bi;
            const float unc = -(1.f - cb) * (logf(fmaxf(sp_y, eps)) + logf(fmaxf(h_y, eps)));
            const float cen = -cb * logf(fmaxf(sp_y1, eps));
            *h.loss = (1.f - alpha) * (cen + unc) + alpha * unc;
          }
        }
      }
      named_bar_sync(1, HIDDEN_ET);
      float dlv[HEAD_MAX_K];
#pragma unroll
      for (int j = 0; j < HEAD_MAX_K; ++j) dlv[j] = s_dl[j];
      const float dot = s_dl[HEAD_MAX_K];
      const float inv_l = 1.f / s_dl[HEAD_MAX_K + 1];
      float d0 = 0.f, d1 = 0.f;      // dM = Wk^T dlogits, this thread's two columns
#pragma unroll
      for (int j = 0; j < HEAD_MAX_K; ++j) {
        d0 = fmaf(dlv[j], wk[j].x, d0);
        d1 = fmaf(dlv[j], wk[j].y, d1);
      }
      if (rg == 0) *reinterpret_cast<float2*>(vec + C::V_DM + 2 * cp) = make_float2(d0, d1);
      if (blockIdx.x == 0 && rg == 0) *reinterpret_cast<float2*>(h.dM + 2 * cp) = make_float2(d0, d1);
      MMF_HS(13);
      // ---- phase A: t_i = dlogits · z_i, p_i from the global (m, l) just formed -----------------------------
      if (e < 128) {
        float ds = 0.f, p = 0.f;
        if (prow_ok) {
          float t = dlv[0] * z0.x;
          t = fmaf(dlv[1], z0.y, t); t = fmaf(dlv[2], z0.z, t); t = fmaf(dlv[3], z0.w, t);
          t = fmaf(dlv[4], z1.x, t); t = fmaf(dlv[5], z1.y, t); t = fmaf(dlv[6], z1.z, t); t = fmaf(dlv[7], z1.w, t);
          p = __expf(s_raw - m) * inv_l;
          ds = p * (t - dot) + dA;
        }
        vec[C::V_DS + e] = ds;
        vec[C::V_P + e] = p;
        const float sum = warp_sum(ds);
        if (lane == 0) atomicAdd(&s_dbc, sum);
      }
    } else {
      // one warp per row, rows software-pipelined two at a time: t_i = dM · h_i from the stashed H tile
      constexpr int HJ = L / 256;
      float dmv[HJ][8];
      float dotMM = 0.f;
#pragma unroll
      for (int j = 0; j < HJ; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          dmv[j][k] = __ldg(a.dM + 8 * lane + 256 * j + k);
          dotMM = fmaf(dmv[j][k], __ldg(a.M + 8 * lane + 256 * j + k), dotMM);
        }
      dotMM = warp_sum(dotMM);
      const float m = __ldg(a.ml), inv_l = 1.0f / __ldg(a.ml + 1);
      float acc_ds = 0.f;
      // the warp's raw scores / incoming score gradients, fetched once (lane l <-> l-th row of this warp): a load
      // inside the row loop would put one L2 round trip per row on lane 0's critical path
      float s_raw_l = 0.f, dA_l = 0.f;
      {
        const long long row = row0 + (long long)ew * ROWS_PER_WARP + lane;
        if (lane < ROWS_PER_WARP && row < a.N) {
          s_raw_l = __ldg(a.A_raw + row);
          dA_l = a.dA_raw ? __ldg(a.dA_raw + row) : 0.f;
        }
      }
      constexpr int PA_STEP = 2, PA_STEPS = ROWS_PER_WARP / PA_STEP;
      static_assert(ROWS_PER_WARP % PA_STEP == 0, "phase A step");
      uint4 hv[2][PA_STEP][HJ];
      auto pa_load = [&](int step, int buf) {
#pragma unroll
        for (int u = 0; u < PA_STEP; ++u) {
          const long long row = row0 + (long long)ew * ROWS_PER_WARP + step * PA_STEP + u;
          const uint4* hp = reinterpret_cast<const uint4*>(a.H + (row < a.N ? row : 0) * L);
#pragma unroll
          for (int j = 0; j < HJ; ++j) hv[buf][u][j] = __ldg(hp + lane + 32 * j);
        }
      };
      pa_load(0, 0);
#pragma unroll
      for (int step = 0; step < PA_STEPS; ++step) {
        const int buf = step & 1;
        if (step + 1 < PA_STEPS) pa_load(step + 1, buf ^ 1);
#pragma unroll
        for (int u = 0; u < PA_STEP; ++u) {
          const int r = (int)ew * ROWS_PER_WARP + step * PA_STEP + u;
          const bool ok = row0 + r < a.N;
          float t0 = 0.f, t1 = 0.f;   // two partial sums: halves the dependent FMA chain
#pragma unroll
          for (int j = 0; j < HJ; ++j) {
            const uint32_t w[4] = {hv[buf][u][j].x, hv[buf][u][j].y, hv[buf][u][j].z, hv[buf][u][j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = unpack_bf16x2(w[k]);
              t0 = fmaf(f.x, dmv[j][2 * k], t0);
              t1 = fmaf(f.y, dmv[j][2 * k + 1], t1);
            }
          }
          const float t = warp_sum(t0 + t1);
          const float s_raw = __shfl_sync(0xffffffffu, s_raw_l, step * PA_STEP + u);
          const float dA = __shfl_sync(0xffffffffu, dA_l, step * PA_STEP + u);
          if (lane == 0) {
            float p = 0.f, ds = 0.f;
            if (ok) {
              p = __expf(s_raw - m) * inv_l;
              ds = p * (t - dotMM) + dA;
            }
            vec[C::V_DS + r] = ds;
            vec[C::V_P + r] = p;
            acc_ds += ds;
          }
        }
      }
      if (lane == 0) atomicAdd(&s_dbc, acc_ds);
    }
    named_bar_sync(1, HIDDEN_ET);   // ds / p / vectors visible to all worker threads
    if (e == 0) MMF_STAMP(a, 4);

    // ---------------- mainloop: transform the stashed activations into dG, in place -------------------------
    // thread map of a slice: lane = (column quad low bits, row) -> a warp covers 4 column quads x 8 consecutive rows
    // per load (conflict-free under the 128B swizzle), warps = 4 column-quad groups x 4 row quarters. The column
    // sums are then reduced over the warp's 8 rows with shuffles and only lanes 0-3 touch shared memory (4-way
    // contention between the row quarters). Shared-memory fp32 atomicAdd is a CAS loop: the first version had all
    // 512 threads add into 192 addresses per slice (32-way contention) and spent half the kernel retrying.
    static_assert(HIDDEN_EW == 16, "slice thread map assumes 16 worker warps");
    const uint32_t ct = (ew & 3u) * 4u + (lane & 3u);     // column quad 0..15
    const uint32_t rbase = (ew >> 2) * 32u + (lane >> 2); // first row; rows rbase + 8 u
    constexpr int RPT = 4;                                 // rows per thread and slice
    const uint32_t a_ready_leader = mapa_cluster(smem_u32(&bar_aready[0]), 0);
    constexpr float attn_scale = DROP ? (1.0f / 0.75f) : 1.0f;
#pragma unroll 1
    for (int kp = 0; kp < C::NKP; ++kp) {
      const int s = kp % C::NSA;
      const int d0 = kp * 64 + 4 * (int)ct;
      const float4 wc4 = *reinterpret_cast<const float4*>(vec + C::V_WC + d0);
      mbar_wait(smem_u32(&bar_afull[s]), (kp / C::NSA) & 1);
      if (!MMF_HEAD_STAMPS && e == 0 && kp < 3) MMF_STAMP(a, 8 + 2 * kp);
      uint8_t* ta = pool_ptr + (a_ring - pool) + s * C::A_STAGE;
#if MMF_HIDDEN_PACKED
      // packed f32x2 arithmetic (FMUL2 / FFMA2 / FADD2 process a column pair per instruction): the transform is bound by
      // the worker warps' instruction issue (~60 scalar instructions per row-step of 4 columns x 2 branches)
      const float2 wc01 = make_float2(wc4.x, wc4.y), wc23 = make_float2(wc4.z, wc4.w);
      float2 s_wc[2] = {}, s_a[2] = {}, s_g[2] = {};     // column sums: [pair 0 | pair 1] of dwc, dba, dbb
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        const uint32_t r = rbase + 8u * u;
        const uint32_t off = sw128_offset(r, ct >> 1) + (ct & 1u) * 8u;
        uint2* pa = reinterpret_cast<uint2*>(ta + off);
        uint2* pg = reinterpret_cast<uint2*>(ta + 16384 + off);
        const uint2 av = *pa;
        uint2 gv = make_uint2(0u, 0u);
        if (GATED) gv = *pg;
        const float ds = vec[C::V_DS + r];
        const float2 ds2 = make_float2(ds, ds);
        float2 ka[2] = {make_float2(attn_scale, attn_scale), make_float2(attn_scale, attn_scale)};
        float2 kg[2] = {make_float2(GATED && DROP ? attn_scale : 1.f, GATED && DROP ? attn_scale : 1.f),
                        make_float2(GATED && DROP ? attn_scale : 1.f, GATED && DROP ? attn_scale : 1.f)};
        if (DROP) {
          const uint32_t ab = drop_bits16(drop_row_state(a.seed, 1, (uint32_t)(row0 + r)), (uint32_t)(d0 >> 4));
          const uint32_t gb = drop_bits16(drop_row_state(a.seed, 2, (uint32_t)(row0 + r)), (uint32_t)(d0 >> 4));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            ka[k].x = drop_keep(ab, (d0 & 15) + 2 * k) ? attn_scale : 0.f;
            ka[k].y = drop_keep(ab, (d0 & 15) + 2 * k + 1) ? attn_scale : 0.f;
            if (GATED) {
              kg[k].x = drop_keep(gb, (d0 & 15) + 2 * k) ? attn_scale : 0.f;
              kg[k].y = drop_keep(gb, (d0 & 15) + 2 * k + 1) ? attn_scale : 0.f;
            }
          }
        }
        const float2 aa[2] = {__half22float2(*reinterpret_cast<const __half2*>(&av.x)),
                              __half22float2(*reinterpret_cast<const __half2*>(&av.y))};
        float2 gg[2] = {make_float2(1.f, 1.f), make_float2(1.f, 1.f)};
        if (GATED) {
          gg[0] = __half22float2(*reinterpret_cast<const __half2*>(&gv.x));
          gg[1] = __half22float2(*reinterpret_cast<const __half2*>(&gv.y));
        }
        const float2 dsw[2] = {__fmul2_rn(ds2, wc01), __fmul2_rn(ds2, wc23)};
        uint32_t oa[2], og[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          // same factoring as the scalar form: q = ds wc gd, dG_a = ka (q - (q a) a), dG_g = (q ad) - (q ad) g
          const float2 gd = (GATED && DROP) ? __fmul2_rn(gg[k], kg[k]) : gg[k];
          const float2 ad = DROP ? __fmul2_rn(aa[k], ka[k]) : aa[k];
          s_wc[k] = __ffma2_rn(ds2, __fmul2_rn(ad, gd), s_wc[k]);
          const float2 q = __fmul2_rn(dsw[k], gd);
          const float2 qa = __fmul2_rn(q, aa[k]);
          const float2 qad = DROP ? __fmul2_rn(q, ad) : qa;
          float2 da = __ffma2_rn(make_float2(-qa.x, -qa.y), aa[k], q);
          if (DROP) da = __fmul2_rn(da, ka[k]);
          const float2 dg = GATED ? __ffma2_rn(make_float2(-qad.x, -qad.y), gg[k], qad) : make_float2(0.f, 0.f);
          s_a[k] = __fadd2_rn(s_a[k], da);
          s_g[k] = __fadd2_rn(s_g[k], dg);
          oa[k] = pack_bf16x2(da.x, da.y);
          og[k] = pack_bf16x2(dg.x, dg.y);
        }
        *pa = make_uint2(oa[0], oa[1]);
        if (GATED) *pg = make_uint2(og[0], og[1]);
      }
      float acc_wc[4] = {s_wc[0].x, s_wc[0].y, s_wc[1].x, s_wc[1].y};
      float acc_a[4] = {s_a[0].x, s_a[0].y, s_a[1].x, s_a[1].y};
      float acc_g[4] = {s_g[0].x, s_g[0].y, s_g[1].x, s_g[1].y};
#else
      const float wcv[4] = {wc4.x, wc4.y, wc4.z, wc4.w};
      float acc_wc[4] = {}, acc_a[4] = {}, acc_g[4] = {};
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        const uint32_t r = rbase + 8u * u;
        const uint32_t off = sw128_offset(r, ct >> 1) + (ct & 1u) * 8u;
        uint2* pa = reinterpret_cast<uint2*>(ta + off);
        uint2* pg = reinterpret_cast<uint2*>(ta + 16384 + off);
        const uint2 av = *pa;
        uint2 gv = make_uint2(0u, 0u);
        if (GATED) gv = *pg;
        const float ds = vec[C::V_DS + r];
        uint32_t ab = 0xFFFFFFFFu, gb = 0xFFFFFFFFu;
        if (DROP) {
          ab = drop_bits16(drop_row_state(a.seed, 1, (uint32_t)(row0 + r)), (uint32_t)(d0 >> 4));
          gb = drop_bits16(drop_row_state(a.seed, 2, (uint32_t)(row0 + r)), (uint32_t)(d0 >> 4));
        }
        const float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&av.x));
        const float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&av.y));
        float2 g01 = make_float2(1.f, 1.f), g23 = make_float2(1.f, 1.f);
        if (GATED) {
          g01 = __half22float2(*reinterpret_cast<const __half2*>(&gv.x));
          g23 = __half22float2(*reinterpret_cast<const __half2*>(&gv.y));
        }
        const float aa[4] = {a01.x, a01.y, a23.x, a23.y};
        const float gg[4] = {g01.x, g01.y, g23.x, g23.y};
        float da[4], dg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float ka = (!DROP || drop_keep(ab, (d0 & 15) + k)) ? attn_scale : 0.f;
          const float kg = (GATED && DROP) ? (drop_keep(gb, (d0 & 15) + k) ? attn_scale : 0.f) : 1.f;
          const float ad = aa[k] * ka, gd = gg[k] * kg;
          acc_wc[k] = fmaf(ds, ad * gd, acc_wc[k]);
          // dG_a = dq gd ka (1 - a^2) and dG_g = dq ad kg g (1 - g), factored through q = dq gd (gd = g kg):
          //   dG_a = ka (q - (q a) a),  dG_g = (q ad) - (q ad) g        (9 instead of 12 operations per element pair)
          const float q = ds * wcv[k] * gd;
          const float qa = q * aa[k];
          const float qad = DROP ? q * ad : qa;
          da[k] = ka * fmaf(-qa, aa[k], q);
          dg[k] = GATED ? fmaf(-qad, gg[k], qad) : 0.f;
          acc_a[k] += da[k];
          acc_g[k] += dg[k];
        }
        *pa = make_uint2(pack_bf16x2(da[0], da[1]), pack_bf16x2(da[2], da[3]));
        if (GATED) *pg = make_uint2(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3]));
      }
#endif
      fence_proxy_async_smem();
      __syncwarp();
      if (!MMF_HEAD_STAMPS && e == 0 && kp < 3) MMF_STAMP(a, 9 + 2 * kp);
      if (lane == 0) {
        if (!MMF_HIDDEN_RELAY) mbar_arrive_cluster(a_ready_leader + s * 8u);   // bar_aready[s] of the leader
        mbar_arrive(smem_u32(&bar_astore[s]));   // the store thread (warp 2) relays to the MMA and writes the stage out as dG
      }
      // column sums of this slice over the warp's 8 row lanes (lane bits 2-4), then the 4 row quarters meet in shared memory
#if MMF_HIDDEN_PACKED
      // halving butterfly: 12 sums per thread (dwc | dba | dbb x 4 columns). At each of the 3 levels a lane keeps half of
      // its values and hands the other half to its partner: 6 + 3 + 2 shuffles instead of 3 x 12. Afterwards the row
      // lane rl = lane >> 2 holds: level 1 (bit 4) split wc/a/g halves ... the owner lane of sum j is given by owner_of().
      {
        float v12[12] = {acc_wc[0], acc_wc[1], acc_wc[2], acc_wc[3], acc_a[0], acc_a[1], acc_a[2], acc_a[3],
                         acc_g[0], acc_g[1], acc_g[2], acc_g[3]};
        const bool up4 = (lane & 16u) != 0u, up3 = (lane & 8u) != 0u, up2 = (lane & 4u) != 0u;
        float v6[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {          // level 1 (xor 16): lower half keeps 0..5, upper half keeps 6..11
          const float send = up4 ? v12[j] : v12[j + 6];
          const float got = __shfl_xor_sync(0xffffffffu, send, 16);
          v6[j] = (up4 ? v12[j + 6] : v12[j]) + got;
        }
        float v3[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {          // level 2 (xor 8): keeps 0..2 / 3..5 of its six
          const float send = up3 ? v6[j] : v6[j + 3];
          const float got = __shfl_xor_sync(0xffffffffu, send, 8);
          v3[j] = (up3 ? v6[j + 3] : v6[j]) + got;
        }
        // level 3 (xor 4): three values, both lanes end with all three sums
#pragma unroll
        for (int j = 0; j < 3; ++j) v3[j] += __shfl_xor_sync(0xffffffffu, v3[j], 4);
        // lane (up4, up3, *) holds sums 6 up4 + 3 up3 + {0,1,2}; one writer per (up4, up3): the lanes with bit 2 clear
        if (!up2) {
          const int base = 6 * (int)up4 + 3 * (int)up3;
          const uint32_t cq = lane & 3u;       // column quad within the warp's 16 columns
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int idx = base + j;          // 0..11 = which * 4 + k
            s_part[ew][kp * 48 + (idx >> 2) * 16 + cq * 4 + (idx & 3)] = v3[j];
          }
        }
      }
#else
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
          acc_wc[k] += __shfl_xor_sync(0xffffffffu, acc_wc[k], o);
          acc_a[k] += __shfl_xor_sync(0xffffffffu, acc_a[k], o);
          if (GATED) acc_g[k] += __shfl_xor_sync(0xffffffffu, acc_g[k], o);
        }
      }
      if (lane < 4) {
        float* slot = &s_part[ew][kp * 48 + lane * 4];
        *reinterpret_cast<float4*>(slot) = make_float4(acc_wc[0], acc_wc[1], acc_wc[2], acc_wc[3]);
        *reinterpret_cast<float4*>(slot + 16) = make_float4(acc_a[0], acc_a[1], acc_a[2], acc_a[3]);
        if (GATED) *reinterpret_cast<float4*>(slot + 32) = make_float4(acc_g[0], acc_g[1], acc_g[2], acc_g[3]);
      }
#endif
    }
    if (e == 0) MMF_STAMP(a, 5);

    // ---------------- epilogue: dU tile -------------------------------------------------------------------
    const uint32_t r = q * 32 + lane;
    constexpr int PIECES = L / (8 * HIDDEN_EW);   // 32-column pieces per thread (L / 32 pieces over EW / 4 parts)
    const float p_row = vec[C::V_P + r];
    // ReLU mask words of this thread's row and column part (written by the training forward; rows past N: 0)
    uint32_t mw[PIECES];
#pragma unroll
    for (int ii = 0; ii < PIECES; ++ii)
      mw[ii] = (row0 + r < a.N) ? __ldg(a.mask + (row0 + r) * (L / 32) + part * PIECES + ii) : 0u;
    mbar_wait(smem_u32(&bar_acc), 0);   // every MMA retired: accumulator complete, both rings idle
    tc_fence_after();
    if (e == 0) MMF_STAMP(a, 6);
    // (the dG stores of warp 2 have finished reading the A ring: bar_aempty completed before the last MMAs could
    //  be issued ... except the final NSA slices: wait for their store thread explicitly)
#pragma unroll
    for (int j = 1; j <= C::NSA; ++j)
      if (C::NKP >= j) mbar_wait(smem_u32(&bar_aempty[(C::NKP - j) % C::NSA]), ((C::NKP - j) / C::NSA) & 1);
    named_bar_sync(1, HIDDEN_ET);
    // Each thread owns one row x PIECES 32-column pieces. The four quadrant warps of a column part complete one
    // 64-column block of the staged tile every two pieces: they meet on their own named barrier and one of them
    // TMA-stores the block while the others go on (the first version staged the whole 128 KB tile, then issued all
    // stores, then walked the staged tile again for db1: 9.3k cycles, gpurun_out/r2_phase9.log). db1 comes from the
    // registers: 31-shuffle transpose-reduce per piece, the four quadrants meet in shared memory.
    float* s_db1 = reinterpret_cast<float*>(pool_ptr + C::STAGING);   // [4 quadrants][L], behind the staged tile
    static_assert(C::POOL >= C::STAGING + 4u * L * 4u, "db1 scratch must fit behind the staged tile");
    static_assert(PIECES == 2 || PIECES == 4, "a column part is one or two 64-column blocks");
    float v[2][32];
    tmem_ld32(tmem + ((q * 32u) << 16) + part * PIECES * 32, v[0]);
#pragma unroll
    for (int ii = 0; ii < PIECES; ++ii) {
      const int cb = part * PIECES + ii;
      tmem_ld_wait();
      if (ii + 1 < PIECES) tmem_ld32(tmem + ((q * 32u) << 16) + (cb + 1) * 32, v[(ii + 1) & 1]);
      float (&u)[32] = v[ii & 1];
      const uint32_t bits = mw[ii];
      const float4* dm4 = reinterpret_cast<const float4*>(vec + C::V_DM + cb * 32);
      // dU = s (acc + p_i dM) = fma(acc, s, (s p_i) dM): two packed f32x2 operations per element PAIR (FMUL2 + FFMA2)
      const float2 s2 = make_float2(a.du_scale, a.du_scale), sp2 = make_float2(a.du_scale * p_row, a.du_scale * p_row);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 d = dm4[i >> 2];
        const float2 o01 = __ffma2_rn(make_float2(u[i], u[i + 1]), s2, __fmul2_rn(sp2, make_float2(d.x, d.y)));
        const float2 o23 = __ffma2_rn(make_float2(u[i + 2], u[i + 3]), s2, __fmul2_rn(sp2, make_float2(d.z, d.w)));
        u[i] = (bits >> i) & 1u ? o01.x : 0.f;
        u[i + 1] = (bits >> (i + 1)) & 1u ? o01.y : 0.f;
        u[i + 2] = (bits >> (i + 2)) & 1u ? o23.x : 0.f;
        u[i + 3] = (bits >> (i + 3)) & 1u ? o23.y : 0.f;
      }
      const uint32_t blk = pool + (cb >> 1) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(blk + sw128_offset(r, (cb & 1) * 4 + j), pack_bf16x2(u[8 * j], u[8 * j + 1]),
                     pack_bf16x2(u[8 * j + 2], u[8 * j + 3]), pack_bf16x2(u[8 * j + 4], u[8 * j + 5]),
                     pack_bf16x2(u[8 * j + 6], u[8 * j + 7]));
      if (ii & 1) {   // this warp's half of block cb >> 1 is staged
        fence_proxy_async_smem();
        named_bar_sync(2 + part, 128);
        if (q == 0 && lane == 0) {
          tma_store_2d_keep(&tmDU, blk, (cb >> 1) * 64, (int)row0);
          tma_store_commit();
        }
      }
      s_db1[q * L + cb * 32 + lane] = warp_colsum32(u);   // (rows past N are zeros: their mask words are 0)
    }
    tc_fence_before();
    named_bar_sync(1, HIDDEN_ET);     // every quadrant's column sums are in shared memory
    for (uint32_t c0 = e; c0 < (uint32_t)L; c0 += HIDDEN_ET)
      atomicAdd(a.db1 + c0, s_db1[c0] + s_db1[L + c0] + s_db1[2 * L + c0] + s_db1[3 * L + c0]);
    // dwc / dbab / dbc of this CTA's rows
    // column d of branch `which` lives in slice d / 64, column group (d % 64) / 16, slot which * 16 + d % 16 of the
    // four row-quarter warps of that column group
    for (int i = e; i < D + C::KD; i += HIDDEN_ET) {
      const int which = i / D, d = i - which * D;
      const int kp = d >> 6, ctg = (d & 63) >> 4, off = kp * 48 + which * 16 + (d & 15);
      const float v = s_part[ctg][off] + s_part[ctg + 4][off] + s_part[ctg + 8][off] + s_part[ctg + 12][off];
      atomicAdd(which == 0 ? a.dwc + d : a.dbab + (which - 1) * D + d, v);
    }
    if (e == 0) atomicAdd(a.dbc, s_dbc);
    if (q == 0 && lane == 0) tma_store_wait_exit();   // the threads that committed the dU block stores
    if (e == 0) MMF_STAMP(a, 7);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  timeline_end(2, tl);
  if (warp == 2) tmem_dealloc_pair(tmem, L);
}

}  // namespace mmf
