// amil_gate_ew.cuh — gate backward from the forward's activation stash (sm_100a, HBM/L2-bound).
//
// The training forward (mmf_amil_fwd_train) leaves in the backward workspace
//   H  bf16 [N, L]    post-ReLU / post-dropout hidden activations
//   AG fp16 [N, KD]   pre-dropout branch outputs  [tanh(Wa h + ba) | sigmoid(Wb h + bb)]
// (fp16, not bf16: a, g live in [-1, 1] and the backward forms 1 - a^2 and 1 - g, where bf16's 8
// mantissa bits would cost up to 0.4 % of the largest element through cancellation).
// This kernel turns AG into dG IN PLACE (same [N, KD] 2-byte footprint, now bf16) without touching
// the tensor cores:
//   t_i  = dM · h_i,   p_i = e^{s_i - m} / l,   ds_i = p_i (t_i - dM·M) + dA_raw_i
//   dG_i = [ ds_i wc ⊙ g ⊙ (1 - a²) | ds_i wc ⊙ a ⊙ g ⊙ (1 - g) ]          
//   dwc += Σ ds_i (a ⊙ g)_i,  dba/dbb += Σ dG_i,  dbc += Σ ds_i   (per-block partials -> workspace)
// i.e. the backward of models/model_modules.py:105-110 + the softmax pooling of
// models/model_attention_mil_path.py:53-56, restricted to what needs no GEMM.
//
// Streaming design (HBM-bound: 4104 B moved per row, ~25 instructions per element pair): a persistent
// grid of one CTA per SM walks 16-row chunks. Chunks are contiguous spans of H and AG, so each is
// fetched with two 1-D bulk copies (cp.async.bulk -> mbarrier) into a 4-stage shared-memory ring —
// three chunks (120 KB) in flight per SM — transformed in place in shared memory, and written back with
// one bulk store. Phase 1 of a chunk: one warp per row reduces t_i over the H row; phase 2:
// column-stationary (thread = 4 columns of each branch x 4 rows), so the dwc / dba / dbb column sums
// are 12 registers per thread for the whole kernel and are added to the outputs with one atomicAdd
// per (CTA, column) at the end. Earlier versions: warp-per-row with 36 accumulators per lane (spilled,
// 34 us); column-stationary with direct global loads (latency-bound bursts, 18 us); see profiles/.
// Algorithmic bytes per row: 2L (H) + 2·2KD (AG in, dG out) + 8 = 4104 B (big preset) -> 67 MB per
// 16k bag.
#pragma once
#include <cuda_fp16.h>

#include "amil_tile.cuh"

namespace mmf {

struct GateEwArgs {
  long long N;
  const __nv_bfloat16* H;   // [N, L]
  void* AG;                 // in: fp16 [N, KD]; out: bf16 dG [N, KD]
  const float* A_raw;       // [N]
  const float* ml;          // (m, l)
  const float* M;           // [L]
  const float* dM;          // [L]
  const float* dA_raw;      // [N] or null
  const float* wc;          // [D]
  float* dwc;               // [D]   accumulated (atomicAdd)
  float* dbab;              // [KD]  accumulated
  float* dbc;               // [1]   accumulated
  uint32_t* mask;           // out: [N, L/32] words, bit j of word w = (h[row, 32w + j] > 0), for the dU GEMM
  int flags;
  unsigned long long seed;
};

template <int L, int D, bool GATED, bool DROP = false>
struct GateEwCfg {
  static constexpr int KD = GATED ? 2 * D : D;
  static constexpr int ROWS = 16;                     // rows per chunk
  static constexpr int STAGES = 4;
  static constexpr int CT = D / 4;                    // column threads per row group (4 columns each)
  static constexpr int RG = 8;                        // row groups (2 rows each per chunk)
  static constexpr int THREADS = CT * RG;             // 768 (D = 384) or 512 (D = 256): 24 / 16 warps hide the
                                                      // smem / conversion latencies of the column phase
  static constexpr uint32_t H_BYTES = ROWS * L * 2u;
  static constexpr uint32_t AG_BYTES = ROWS * KD * 2u;
  static constexpr uint32_t STAGE_BYTES = H_BYTES + AG_BYTES;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 128;
};

template <int L, int D, bool GATED, bool DROP>
__global__ void __launch_bounds__(GateEwCfg<L, D, GATED>::THREADS, 1) amil_gate_ew_kernel(const GateEwArgs a) {
  using C = GateEwCfg<L, D, GATED>;
  static_assert(C::THREADS / 32 >= C::ROWS, "phase 1 is one pass: a warp per row");
  constexpr int KD = C::KD;
  constexpr int HJ = L / 256;   // uint4 (8 bf16) pieces of an H row per lane
  constexpr int WARPS = C::THREADS / 32;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES];
  __shared__ float s_cols[3 * D];
  __shared__ float s_ds[C::ROWS];
  __shared__ float s_dbc;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pool = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* pool_ptr = smem_raw + (pool - smem_u32(smem_raw));
  const long long n_chunks = (a.N + C::ROWS - 1) / C::ROWS;
  // this CTA's chunks: blockIdx.x, + gridDim.x, ...
  const long long my_chunks = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto issue_load = [&](long long k) {   // k-th chunk of this CTA -> stage k % STAGES (one thread)
    const long long ch = blockIdx.x + k * gridDim.x;
    const long long r0 = ch * C::ROWS;
    const uint32_t nr = (uint32_t)min((long long)C::ROWS, a.N - r0);
    const uint32_t st = (uint32_t)(k % C::STAGES);
    const uint32_t bar = smem_u32(&bar_full[st]);
    const uint32_t dst = pool + st * C::STAGE_BYTES;
    mbar_arrive_expect_tx(bar, nr * (uint32_t)(2 * L + 2 * KD));
    bulk_load_1d(dst, a.H + r0 * L, nr * 2u * L, bar);
    bulk_load_1d(dst + C::H_BYTES, reinterpret_cast<const uint16_t*>(a.AG) + r0 * KD, nr * 2u * KD, bar);
  };

  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) mbar_init(smem_u32(&bar_full[s]), 1);
    fence_barrier_init();
    s_dbc = 0.f;
  }
  for (int i = threadIdx.x; i < 3 * D; i += C::THREADS) s_cols[i] = 0.f;
  __syncthreads();
  griddep_wait();
  if (threadIdx.x == 0)
    for (long long k = 0; k < my_chunks && k < C::STAGES - 1; ++k) issue_load(k);

  // phase-1 constants: this lane's slice of dM, and dM·M
  float dmv[HJ][8];
  float dotMM = 0.f;
#pragma unroll
  for (int j = 0; j < HJ; ++j) {
    const int c0 = 8 * lane + 256 * j;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dmv[j][k] = __ldg(a.dM + c0 + k);
      dotMM = fmaf(dmv[j][k], __ldg(a.M + c0 + k), dotMM);
    }
  }
  dotMM = warp_sum(dotMM);
  const float m = __ldg(a.ml), inv_l = 1.0f / __ldg(a.ml + 1);
  // phase-2 constants: this thread's 4 columns
  const int ct = threadIdx.x % C::CT, rg = threadIdx.x / C::CT;
  const int d0 = 4 * ct;
  const float4 wc4 = __ldg(reinterpret_cast<const float4*>(a.wc + d0));
  const float wcv[4] = {wc4.x, wc4.y, wc4.z, wc4.w};
  constexpr bool drop_attn = DROP;   // MMF_DROPOUT_ATTN: masks regenerated from the counter hash
  constexpr float attn_scale = DROP ? (1.0f / 0.75f) : 1.0f;
  float acc_wc[4] = {}, acc_a[4] = {}, acc_g[4] = {};
  float acc_ds = 0.f;

  for (long long k = 0; k < my_chunks; ++k) {
    const long long ch = blockIdx.x + k * gridDim.x;
    const long long r0 = ch * C::ROWS;
    const int nr = (int)min((long long)C::ROWS, a.N - r0);
    const uint32_t st = (uint32_t)(k % C::STAGES);
    // per-row scalars of phase 1 are fetched before the wait (independent of the chunk data)
    float s_raw = 0.f, dA = 0.f;
    if (lane == 0 && (int)warp < nr) {
      s_raw = __ldg(a.A_raw + r0 + warp);
      dA = a.dA_raw ? __ldg(a.dA_raw + r0 + warp) : 0.f;
    }
    mbar_wait(smem_u32(&bar_full[st]), (uint32_t)((k / C::STAGES) & 1));
    const uint8_t* hs = pool_ptr + st * C::STAGE_BYTES;
    uint8_t* ags = pool_ptr + st * C::STAGE_BYTES + C::H_BYTES;
    // ---- phase 1: ds_i, one warp per row ----
    if ((int)warp < nr) {
      const uint4* hp = reinterpret_cast<const uint4*>(hs + warp * (2 * L));
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < HJ; ++j) {
        const uint4 hv = hp[lane + 32 * j];
        const uint32_t w[4] = {hv.x, hv.y, hv.z, hv.w};
        uint32_t byte = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w[q]);
          t = fmaf(f.x, dmv[j][2 * q], t);
          t = fmaf(f.y, dmv[j][2 * q + 1], t);
          byte |= (uint32_t)(f.x > 0.f) << (2 * q) | (uint32_t)(f.y > 0.f) << (2 * q + 1);
        }
        // this lane's 8 columns are byte (lane % 4) of mask word lane / 4 + 8 j: OR the 4-lane group together
        uint32_t word = byte << (8 * (lane & 3));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        if ((lane & 3) == 0) a.mask[(r0 + warp) * (L / 32) + (lane >> 2) + 8 * j] = word;
      }
      t = warp_sum(t);
      if (lane == 0) {
        const float ds = __expf(s_raw - m) * inv_l * (t - dotMM) + dA;
        s_ds[warp] = ds;
        acc_ds += ds;
      }
    }
    __syncthreads();
    // ---- phase 2: AG -> dG in place in shared memory; thread = 4 columns x rows rg, rg + 8 ----
#pragma unroll
    for (int u = 0; u < C::ROWS / C::RG; ++u) {
      const int r = rg + u * C::RG;
      if (r < nr) {
        uint2* pa = reinterpret_cast<uint2*>(ags + r * (2 * KD)) + ct;
        const uint2 av = pa[0];
        uint2 gv = make_uint2(0u, 0u);
        if (GATED) gv = pa[D / 4];
        const float ds = s_ds[r];
        uint32_t ab = 0xFFFFFFFFu, gb = 0xFFFFFFFFu;
        if (drop_attn) {
          ab = drop_bits16(drop_row_state(a.seed, 1, (uint32_t)(r0 + r)), (uint32_t)(d0 >> 4));
          gb = drop_bits16(drop_row_state(a.seed, 2, (uint32_t)(r0 + r)), (uint32_t)(d0 >> 4));
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
        for (int q = 0; q < 4; ++q) {
          const float ka = (!DROP || drop_keep(ab, (d0 & 15) + q)) ? attn_scale : 0.f;
          const float kg = (GATED && DROP) ? (drop_keep(gb, (d0 & 15) + q) ? attn_scale : 0.f) : 1.f;
          const float ad = aa[q] * ka, gd = gg[q] * kg;
          const float dq = ds * wcv[q];
          acc_wc[q] = fmaf(ds, ad * gd, acc_wc[q]);
          da[q] = dq * gd * ka * (1.f - aa[q] * aa[q]);
          dg[q] = GATED ? dq * ad * kg * gg[q] * (1.f - gg[q]) : 0.f;
          acc_a[q] += da[q];
          acc_g[q] += dg[q];
        }
        pa[0] = make_uint2(pack_bf16x2(da[0], da[1]), pack_bf16x2(da[2], da[3]));
        if (GATED) pa[D / 4] = make_uint2(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3]));
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store_1d(reinterpret_cast<uint16_t*>(a.AG) + r0 * KD, pool + st * C::STAGE_BYTES + C::H_BYTES,
                    (uint32_t)nr * 2u * KD);
      tma_store_commit();
      // chunk k + STAGES - 1 reuses the stage of chunk k - 1: its store must have finished reading smem
      if (k + C::STAGES - 1 < my_chunks) {
        bulk_store_wait_read<1>();
        issue_load(k + C::STAGES - 1);
      }
    }
  }

  // CTA-level reduction over the row groups, then one atomicAdd per column into the outputs
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    atomicAdd(&s_cols[d0 + q], acc_wc[q]);
    atomicAdd(&s_cols[D + d0 + q], acc_a[q]);
    if (GATED) atomicAdd(&s_cols[2 * D + d0 + q], acc_g[q]);
  }
  if (lane == 0 && acc_ds != 0.f) atomicAdd(&s_dbc, acc_ds);
  __syncthreads();
  if (my_chunks > 0) {
    for (int i = threadIdx.x; i < D; i += C::THREADS) atomicAdd(a.dwc + i, s_cols[i]);
    for (int i = threadIdx.x; i < KD; i += C::THREADS) atomicAdd(a.dbab + i, s_cols[D + i]);
    if (threadIdx.x == 0) atomicAdd(a.dbc, s_dbc);
  }
  if (threadIdx.x == 0) tma_store_wait_all();
}

}  // namespace mmf
