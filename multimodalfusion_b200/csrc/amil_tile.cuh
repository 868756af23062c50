// amil_tile.cuh — the fused gated attention-MIL tile kernel (sm_100a).
//
// One CTA owns 128 instances (rows) of a bag and runs, without the [N,L] hidden activations or
// the [N,2D] attention branches ever leaving the SM:
//
//   GEMM1  U[128,L]   = X_tile[128,1024] · W1^T          TMA ring -> tcgen05.mma -> TMEM
//   EPI1   H          = dropout(relu(U + b1)) -> bf16, written into shared memory directly in
//                       the UMMA K-major / 128B-swizzled layout (it is GEMM2's A operand)
//   GEMM2  [A|G]_c    = H · [Wa;Wb]_c^T  per 128-wide chunk c of D (N = 256: 128 tanh columns
//                       and the matching 128 sigmoid columns), double-buffered in TMEM
//   EPI2   s_i        = Σ_d wc_d · tanh(A+ba) · sigmoid(G+bb)  (+ bc)      per row, in registers
//   FWD :  tile softmax partial (m, l, Σ e^{s-m} h) from the resident H tile; A_raw[N] out
//   BWD :  ds_i = p_i (dM·h_i − dM·M) + dA_raw_i ; dG = [dq g (1−a²) | dq a g (1−g)] -> HBM (bf16),
//          H tile -> HBM (TMA store), column sums for dwc / dba / dbb / dbc
//
// Math follows SURVEY.md Appendix A.1/A.2, i.e. the reference ops
//   models/model_attention_mil_path.py:20-21,29,52-56 and models/model_modules.py:84-85,105-110.
//
// Shared-memory pool (224 KB, 1024-byte aligned):
//   [0, H_BYTES)            H tile: L/64 k-blocks of [128 rows][128 B]
//   [H_BYTES, POOL)         GEMM2 ring: NS2 stages of [CHN rows][128 B] (Wab chunk k-block)
//   GEMM1 ring (NS1 stages of x[128][128B] + W1[L][128B]) overlays the pool from offset 0: H does
//   not exist before GEMM1 has drained, and Wab loads start only after GEMM1's last MMA retired.
#pragma once
#include "amil_head_tail.cuh"
#include "mmf_ptx.cuh"

namespace mmf {

enum { AMIL_FWD = 0, AMIL_BWD_GATE = 1 };

template <int L, int D, bool GATED>
struct AmilCfg {
  static_assert(L == 256 || L == 512, "fc width must be 256 (small) or 512 (big)");
  static_assert(D == 256 || D == 384, "attention width must be 256 (small) or 384 (big)");
  static constexpr int KB1 = 1024 / 64;
  static constexpr int NH1 = L / 256;
  static constexpr int KB2 = L / 64;
  static constexpr int NCH = D / 128;
  static constexpr int CHN = GATED ? 256 : 128;
  static constexpr uint32_t H_BYTES = 128u * L * 2u;
  static constexpr uint32_t STAGE1 = 16384u + L * 128u;
  static constexpr uint32_t STAGE2 = CHN * 128u;
  static constexpr uint32_t POOL = 224u * 1024u;
  static constexpr int NS1 = (POOL / STAGE1) < 4 ? (POOL / STAGE1) : 4;
  static constexpr int NS2 = ((POOL - H_BYTES) / STAGE2) < 4 ? ((POOL - H_BYTES) / STAGE2) : 4;
  static constexpr uint32_t SMEM_BYTES = POOL + 1024u;
  static constexpr int NCOLS = GATED ? 3 * D : 2 * D;  // colsum workspace row: dwc | dba | dbb
};

struct AmilArgs {
  long long N;
  const float* b1;
  const float* bab;
  const float* wc;
  const float* bc;
  float* A_raw;     // fwd: out [N]; bwd: in
  float* partials;  // fwd: [tiles, L+2]
  int store_h;      // fwd: also TMA-store the H tile (stash). bwd: always stored.
  uint16_t* AG;     // fwd stash: pre-dropout [tanh | sigmoid] branch outputs, fp16 [N, ldag] (or null)
  long long ldag;
  const int* tile_valid;  // fwd, varlen inference (optional): valid rows (0..128) of every 128-row tile of a packed
                          // multi-bag buffer; null = one bag of N rows
  float4* zero_ptr;   // fwd (optional): buffer the epilogue warps clear while GEMM1 runs ("zero_grad" of the step)
  long long zero_n4;  // its length in float4
  uint32_t* mask_out; // fwd train (optional): ReLU mask words [N, L/32], bit j of word w = (h[row][32 w + j] > 0)
  float* z_out;       // fwd train (optional): z_i = Wk h_i as fp32 [N, zld] from the N = 16 side MMA (needs tmWk)
  int zld;            // 4 or 8
  unsigned int* gflags; // fwd train (optional): [HEAD_MAX_GROUPS] flags of the backward's two-level head merge, cleared here
  int flags;
  unsigned long long seed;
  // backward only
  const float* ml;
  const float* M;
  const float* dM;
  const float* dA_raw;
  __nv_bfloat16* dG;  // [N, lddg] natural column order (a-branch cols, then g-branch cols)
  long long lddg;
  float* colsum_ws;   // [tiles*4, NCOLS]
  float* dbc_ws;      // [tiles*4]
  unsigned long long* dbg;  // optional [gridDim.x, 16] clock64 phase stamps (mmf_debug_set_timing_buffer)
};

#define MMF_STAMP(a, i) do { if ((a).dbg) (a).dbg[(long long)blockIdx.x * 16 + (i)] = clock64(); } while (0)

// ---- counter-based dropout bits (shared with the oracle: oracle/dropout_mask.py) -------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// per-row state; stream: 0 = h, 1 = tanh branch, 2 = sigmoid branch
__host__ __device__ __forceinline__ uint32_t drop_row_state(unsigned long long seed, uint32_t stream,
                                                            uint32_t row) {
  uint32_t x = mix32((uint32_t)seed ^ (stream * 0x9E3779B9U));
  return mix32(x + row * 0x85EBCA6BU + (uint32_t)(seed >> 32));
}
// One 32-bit hash covers 16 consecutive columns, 2 bits each: an element is DROPPED iff its 2-bit
// field is 0 (p_drop = 1/4 exactly).
__host__ __device__ __forceinline__ uint32_t drop_bits16(uint32_t row_state, uint32_t cg16) {
  return mix32(row_state ^ (cg16 * 0xC2B2AE35U));
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t bits16, uint32_t idx) {
  return ((bits16 >> (2u * idx)) & 3u) != 0u;
}
// 4-column view used by the single-CTA kernel: the 8 bits of columns 4*cg4 .. 4*cg4+3
__host__ __device__ __forceinline__ uint32_t drop_bits4(uint32_t row_state, uint32_t cg4) {
  return (drop_bits16(row_state, cg4 >> 2) >> (8u * (cg4 & 3u))) & 0xFFu;
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t bits4, uint32_t j) {
  return ((bits4 >> (2u * j)) & 3u) != 0u ? (1.0f / 0.75f) : 0.0f;
}

template <int L, int D, bool GATED, int MODE>
__global__ void __launch_bounds__(256, 1)
amil_tile_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmWab, const __grid_constant__ CUtensorMap tmH,
                 const AmilArgs a) {
  using C = AmilCfg<L, D, GATED>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full1[C::NS1], bar_empty1[C::NS1];
  __shared__ __align__(8) uint64_t bar_full2[C::NS2], bar_empty2[C::NS2];
  __shared__ __align__(8) uint64_t bar_acc1, bar_h, bar_acc2_full[2], bar_acc2_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float sP[128];
  __shared__ float sRed[8];

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pool = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t h_base = pool;
  const uint32_t ring2 = pool + C::H_BYTES;
  const int tile = blockIdx.x;
  const long long row0 = (long long)tile * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NS1; ++s) { mbar_init(smem_u32(&bar_full1[s]), 1); mbar_init(smem_u32(&bar_empty1[s]), 1); }
    for (int s = 0; s < C::NS2; ++s) { mbar_init(smem_u32(&bar_full2[s]), 1); mbar_init(smem_u32(&bar_empty2[s]), 1); }
    mbar_init(smem_u32(&bar_acc1), 1);
    mbar_init(smem_u32(&bar_h), 128);
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_acc2_full[b]), 1); mbar_init(smem_u32(&bar_acc2_empty[b]), 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmWab);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer =====================================
    for (int kb = 0; kb < C::KB1; ++kb) {
      const int s = kb % C::NS1;
      const uint32_t ph = (kb / C::NS1) & 1;
      mbar_wait(smem_u32(&bar_empty1[s]), ph ^ 1);
      const uint32_t full = smem_u32(&bar_full1[s]);
      const uint32_t dst = pool + s * C::STAGE1;
      mbar_arrive_expect_tx(full, C::STAGE1);
      tma_load_2d(dst, &tmX, full, kb * 64, (int)row0);
#pragma unroll
      for (int j = 0; j < C::NH1; ++j)
        tma_load_2d(dst + 16384 + j * 32768, &tmW1, full, kb * 64, j * 256);
    }
    // GEMM1's stages overlay H and the GEMM2 ring: wait until its last MMA has retired.
    mbar_wait(smem_u32(&bar_acc1), 0);
    for (int c = 0; c < C::NCH; ++c) {
      for (int kb = 0; kb < C::KB2; ++kb) {
        const int it = c * C::KB2 + kb;
        const int s = it % C::NS2;
        const uint32_t ph = (it / C::NS2) & 1;
        mbar_wait(smem_u32(&bar_empty2[s]), ph ^ 1);
        const uint32_t full = smem_u32(&bar_full2[s]);
        mbar_arrive_expect_tx(full, C::STAGE2);
        tma_load_2d(ring2 + s * C::STAGE2, &tmWab, full, kb * 64, c * C::CHN);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // =============================== MMA issuer =======================================
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256, 0, 0);
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, C::CHN, 0, 0);
    for (int kb = 0; kb < C::KB1; ++kb) {
      const int s = kb % C::NS1;
      const uint32_t ph = (kb / C::NS1) & 1;
      mbar_wait(smem_u32(&bar_full1[s]), ph);
      tc_fence_after();
      const uint32_t xs = pool + s * C::STAGE1;
      const uint32_t ws = xs + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_desc_sw128(xs + k * 32, 16, 1024);
#pragma unroll
        for (int j = 0; j < C::NH1; ++j) {
          const uint64_t bd = umma_desc_sw128(ws + j * 32768 + k * 32, 16, 1024);
          umma_bf16_ss(tmem + j * 256, ad, bd, idesc1, (kb | k) != 0);
        }
      }
      umma_commit(smem_u32(&bar_empty1[s]));
    }
    umma_commit(smem_u32(&bar_acc1));

    // H tile complete in smem (and TMEM columns of GEMM1 drained)
    mbar_wait(smem_u32(&bar_h), 0);
    tc_fence_after();
    if (MODE == AMIL_BWD_GATE || a.store_h) {
      for (int kb = 0; kb < C::KB2; ++kb) tma_store_2d(&tmH, h_base + kb * 16384, kb * 64, (int)row0);
      tma_store_commit();
    }
    for (int c = 0; c < C::NCH; ++c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(&bar_acc2_empty[buf]), ((c >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < C::KB2; ++kb) {
        const int it = c * C::KB2 + kb;
        const int s = it % C::NS2;
        const uint32_t ph = (it / C::NS2) & 1;
        mbar_wait(smem_u32(&bar_full2[s]), ph);
        tc_fence_after();
        const uint32_t hs = h_base + kb * 16384;
        const uint32_t bs = ring2 + s * C::STAGE2;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem + buf * C::CHN, umma_desc_sw128(hs + k * 32, 16, 1024),
                       umma_desc_sw128(bs + k * 32, 16, 1024), idesc2, (kb | k) != 0);
        umma_commit(smem_u32(&bar_empty2[s]));
      }
      umma_commit(smem_u32(&bar_acc2_full[buf]));
    }
    if (MODE == AMIL_BWD_GATE || a.store_h) tma_store_wait_all();
  } else if (warp >= 4) {
    // =============================== epilogue warps ===================================
    const uint32_t q = warp & 3;
    const uint32_t r = q * 32 + lane;          // row within the tile == TMEM lane
    const long long row = row0 + r;
    const bool row_ok = row < a.N;
    const uint32_t tq = tmem + ((q * 32u) << 16);
    const bool drop_h = (a.flags & MMF_DROPOUT_H) != 0;
    const bool drop_attn = (a.flags & MMF_DROPOUT_ATTN) != 0;
    const uint32_t rs_h = drop_row_state(a.seed, 0, (uint32_t)row);

    // ---------------- EPI1: H = dropout(relu(U + b1)) -> swizzled smem -----------------
    mbar_wait(smem_u32(&bar_acc1), 0);
    tc_fence_after();
    float t_i = 0.f;  // BWD: dM · h_i
#pragma unroll 1
    for (int cb = 0; cb < L / 32; ++cb) {
      float v[32];
      tmem_ld32(tq + cb * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b1 + cb * 32 + i));
        v[i] = fmaxf(v[i] + b4.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + b4.y, 0.f);
        v[i + 2] = fmaxf(v[i + 2] + b4.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + b4.w, 0.f);
        if (drop_h) {
          const uint32_t bits = drop_bits4(rs_h, (uint32_t)(cb * 8 + (i >> 2)));
          v[i] *= drop_scale(bits, 0); v[i + 1] *= drop_scale(bits, 1);
          v[i + 2] *= drop_scale(bits, 2); v[i + 3] *= drop_scale(bits, 3);
        }
      }
      const uint32_t kb_base = h_base + (cb >> 1) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t p0 = pack_bf16x2(v[8 * j], v[8 * j + 1]), p1 = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        const uint32_t p2 = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), p3 = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        st_shared_v4(kb_base + sw128_offset(r, (cb & 1) * 4 + j), p0, p1, p2, p3);
        if (MODE == AMIL_BWD_GATE) {
          const uint32_t pk[4] = {p0, p1, p2, p3};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = unpack_bf16x2(pk[e]);
            const float2 dm = __ldg(reinterpret_cast<const float2*>(a.dM + cb * 32 + 8 * j + 2 * e));
            t_i = fmaf(hf.x, dm.x, t_i);
            t_i = fmaf(hf.y, dm.y, t_i);
          }
        }
      }
    }
    fence_proxy_async_smem();   // generic-proxy H writes -> visible to UMMA / TMA store
    tc_fence_before();
    mbar_arrive(smem_u32(&bar_h));

    // BWD: ds_i = p_i (t_i - dM·M) + dA_raw_i
    float ds = 0.f;
    if (MODE == AMIL_BWD_GATE) {
      float dot = 0.f;
      for (int c = lane; c < L; c += 32) dot = fmaf(__ldg(a.dM + c), __ldg(a.M + c), dot);
      dot = warp_sum(dot);
      if (row_ok) {
        const float p = __expf(__ldg(a.A_raw + row) - __ldg(a.ml)) / __ldg(a.ml + 1);
        ds = p * (t_i - dot);
        if (a.dA_raw) ds += __ldg(a.dA_raw + row);
      }
    }

    // ---------------- EPI2: gate + score (fwd) / gate backward (bwd) -------------------
    float s_acc = 0.f;
    const uint32_t rs_a = drop_row_state(a.seed, 1, (uint32_t)row);
    const uint32_t rs_g = drop_row_state(a.seed, 2, (uint32_t)row);
#pragma unroll 1
    for (int c = 0; c < C::NCH; ++c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(&bar_acc2_full[buf]), (c >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int pc = 0; pc < 4; ++pc) {
        const int d0 = c * 128 + pc * 32;  // first attention column of this piece
        float va[32], vg[32];
        tmem_ld32(tq + buf * C::CHN + pc * 32, va);
        if (GATED) tmem_ld32(tq + buf * C::CHN + 128 + pc * 32, vg);
        tmem_ld_wait();
        float dwc_v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 ba4 = __ldg(reinterpret_cast<const float4*>(a.bab + d0 + i));
          const float4 wc4 = __ldg(reinterpret_cast<const float4*>(a.wc + d0 + i));
          float4 bb4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (GATED) bb4 = __ldg(reinterpret_cast<const float4*>(a.bab + D + d0 + i));
          const float bav[4] = {ba4.x, ba4.y, ba4.z, ba4.w};
          const float bbv[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
          const float wcv[4] = {wc4.x, wc4.y, wc4.z, wc4.w};
          uint32_t bits_a = 0, bits_g = 0;
          if (drop_attn) {
            bits_a = drop_bits4(rs_a, (uint32_t)((d0 + i) >> 2));
            bits_g = drop_bits4(rs_g, (uint32_t)((d0 + i) >> 2));
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float av = tanh_fast(va[i + e] + bav[e]);
            const float gv = GATED ? sigmoid_fast(vg[i + e] + bbv[e]) : 1.f;
            const float ka = drop_attn ? drop_scale(bits_a, e) : 1.f;
            const float kg = (GATED && drop_attn) ? drop_scale(bits_g, e) : 1.f;
            const float ad = av * ka, gd = gv * kg;
            if (MODE == AMIL_FWD) {
              s_acc = fmaf(wcv[e], ad * gd, s_acc);
            } else {
              const float dq = ds * wcv[e];
              dwc_v[i + e] = ds * ad * gd;
              va[i + e] = dq * gd * ka * (1.f - av * av);             // d pre-tanh
              if (GATED) vg[i + e] = dq * ad * kg * gv * (1.f - gv);  // d pre-sigmoid
            }
          }
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
          // rows beyond N carry ds = 0, so every value above is already 0 for them
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
      mbar_arrive(smem_u32(&bar_acc2_empty[buf]));
    }

    if (MODE == AMIL_BWD_GATE) {
      const float dsum = warp_sum(ds);
      if (lane == 0) a.dbc_ws[(long long)tile * 4 + q] = dsum;
    } else {
      // ---------------- FWD: scores out + tile softmax partial ------------------------
      const float s = row_ok ? s_acc + __ldg(a.bc) : -INFINITY;
      if (row_ok) a.A_raw[row] = s;
      const float wm = warp_max(s);
      if (lane == 0) sRed[q] = wm;
      named_bar_sync(1, 128);
      const float m_t = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
      const float p = row_ok ? __expf(s - m_t) : 0.f;
      sP[r] = p;
      const float wsum = warp_sum(p);
      if (lane == 0) sRed[4 + q] = wsum;
      named_bar_sync(1, 128);
      const float l_t = sRed[4] + sRed[5] + sRed[6] + sRed[7];
      float* prow = a.partials + (long long)tile * (L + 2);
      if (r == 0) { prow[0] = m_t; prow[1] = l_t; }
      // thread r owns column pairs cp = j*128 + r (columns 2cp, 2cp+1)
#pragma unroll
      for (int j = 0; j < L / 256; ++j) {
        const uint32_t col = 2u * (j * 128u + r);
        const uint32_t kb = col >> 6, chunk = (col & 63u) >> 3, inb = (col & 7u) * 2u;
        const uint32_t blk = h_base + kb * 16384u;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
        for (uint32_t rr = 0; rr < 128; ++rr) {
          const float2 hf = unpack_bf16x2(ld_shared_b32(blk + sw128_offset(rr, chunk) + inb));
          const float pr = sP[rr];
          acc0 = fmaf(pr, hf.x, acc0);
          acc1 = fmaf(pr, hf.y, acc1);
        }
        prow[2 + col] = acc0;
        prow[2 + col + 1] = acc1;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// Combine n (m, l, acc[L]) partials. Block = (32 columns, 8 row groups); grid.x = L/32.
// Every block recomputes the global (m, l) (n is a few thousand at most) and owns 32 columns.
__global__ void __launch_bounds__(256)
amil_combine_kernel(const float* __restrict__ parts, long long n, int L, int normalize,
                    float* __restrict__ out, float* __restrict__ ml) {
  __shared__ float s_red[8];
  __shared__ float s_glob[2];
  __shared__ float s_acc[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tx;
  const long long stride = L + 2;
  float m = -INFINITY;
  for (long long t = tid; t < n; t += 256) m = fmaxf(m, parts[t * stride]);
  m = warp_max(m);
  if (tx == 0) s_red[ty] = m;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, s_red[i]);
    s_glob[0] = v;
  }
  __syncthreads();
  m = s_glob[0];
  float l = 0.f;
  for (long long t = tid; t < n; t += 256) {
    const float mt = parts[t * stride];
    if (mt > -INFINITY) l += parts[t * stride + 1] * __expf(mt - m);
  }
  l = warp_sum(l);
  __syncthreads();
  if (tx == 0) s_red[ty] = l;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += s_red[i];
    s_glob[1] = v;
  }
  __syncthreads();
  l = s_glob[1];
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < L) {
    for (long long t = ty; t < n; t += 8) {
      const float mt = parts[t * stride];
      if (mt > -INFINITY) acc = fmaf(parts[t * stride + 2 + c], __expf(mt - m), acc);
    }
  }
  s_acc[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < L) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += s_acc[i][tx];
    if (normalize) out[c] = v / l;
    else out[2 + c] = v;
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (normalize) { ml[0] = m; ml[1] = l; }
    else { out[0] = m; out[1] = l; }
  }
}

}  // namespace mmf
