// gemm2_tc.cuh — CTA-pair (cta_group::2) version of gemm_tc.cuh for the backward GEMMs.
//
// A cluster of two CTAs computes a 256 x BN output tile with M = 256 MMAs: each CTA stages its own
// 128 rows of A and HALF of B, so per 64-deep k-block a CTA pulls 32 KB (BN = 256) or 48 KB
// (BN = 512) through L2 for 2 x 128 x BN x 64 MACs — 1.5-2x the arithmetic intensity of the
// single-CTA 128 x 256 tile, which ncu showed to be L2-latency / bandwidth bound, and a 4-6 stage
// ring instead of 4. Eight epilogue warps (two per TMEM lane quadrant) halve the epilogue.
//
//   (A K-major , B MN-major, EPI_DU,  BN = L)    dU = (dG Wab + p dM^T) ⊙ relu'(H)
//   (A K-major , B MN-major, EPI_STORE)          dx = dU W1
//   (A MN-major, B MN-major, EPI_ATOMIC, split-K) dW1 += dU^T X ; dWab += dG^T H ; dWr += dY^T X
//
// Barrier protocol as in amil_tile2.cuh: full[s] lives in the leader and is completed by the TMA
// loads of both CTAs; empty[s] / acc are per-CTA and receive multicast commits.
#pragma once
#include "gemm_tc.cuh"

namespace mmf {

constexpr int GEMM2_THREADS = 384;

template <int BN>
struct Gemm2Cfg {
  static_assert(BN == 256 || BN == 512, "pair tile is 256 x 256 or 256 x 512");
  static constexpr int NH = BN / 256;
  static constexpr uint32_t A_BYTES = 16384;
  static constexpr uint32_t B_BYTES = NH * 16384u;
  static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 4;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE + 1024;
};

template <int A_MN, int B_MN, int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ TMapSet tmA, const __grid_constant__ TMapSet tmB,
                const GemmArgs g) {
  using C = Gemm2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES];
  __shared__ __align__(8) uint64_t bar_empty[C::STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pool = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int pair_m = blockIdx.x >> 1, n_tile = blockIdx.y;
  const int m0 = pair_m * 256 + 128 * (int)rank;   // first output row of this CTA
  const int n0 = n_tile * BN;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(&tmem_base_slot), BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------
    for (int i = 0; i < nkb; ++i) {
      const int kb = kb0 + i;
      const int s = i % C::STAGES;
      const uint32_t ph = (i / C::STAGES) & 1;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      const uint32_t full = smem_u32(&bar_full[s]);
      const uint32_t a_dst = pool + s * C::STAGE;
      const uint32_t b_dst = a_dst + C::A_BYTES;
      if (leader) mbar_arrive_expect_tx(full, 2 * C::STAGE);
      if (A_MN == 0) {
        const int seg = kb / g.a_seg_kb;
        tma_load_2d_pair(a_dst, &tmA.m[seg], full, (kb - seg * g.a_seg_kb) * 64, m0);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d_pair(a_dst + j * 8192, &tmA.m[0], full, m0 + j * 64, kb * 64);
      }
#pragma unroll
      for (int h = 0; h < C::NH; ++h) {
        const int nb = n0 + 256 * h + 128 * (int)rank;   // this CTA's 128 columns of MMA h
        if (B_MN == 0) {
          tma_load_2d_pair(b_dst + h * 16384, &tmB.m[0], full, kb * 64, nb);
        } else {
          const int seg = nb / g.b_seg_n;
          const int c0 = nb - seg * g.b_seg_n;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_2d_pair(b_dst + h * 16384 + j * 8192, &tmB.m[seg], full, c0 + j * 64, kb * 64);
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------- MMA issuer (leader) ------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, A_MN, B_MN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % C::STAGES;
      const uint32_t ph = (i / C::STAGES) & 1;
      mbar_wait(smem_u32(&bar_full[s]), ph);
      tc_fence_after();
      const uint32_t a_src = pool + s * C::STAGE;
      const uint32_t b_src = a_src + C::A_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = A_MN ? umma_desc_sw128(a_src + k * 2048, 8192, 1024)
                                 : umma_desc_sw128(a_src + k * 32, 16, 1024);
#pragma unroll
        for (int h = 0; h < C::NH; ++h) {
          const uint64_t bd = B_MN ? umma_desc_sw128(b_src + h * 16384 + k * 2048, 8192, 1024)
                                   : umma_desc_sw128(b_src + h * 16384 + k * 32, 16, 1024);
          umma_bf16_ss_pair(tmem + h * 256, ad, bd, idesc, (i | k) != 0);
        }
      }
      umma_commit_pair_mc(smem_u32(&bar_empty[s]), 3);
    }
    umma_commit_pair_mc(smem_u32(&bar_acc), 3);
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs, 8 warps) --------------------
    const uint32_t q = warp & 3;
    const uint32_t half = (warp - 4) >> 2;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < g.M;
    const int tile128 = pair_m * 2 + (int)rank;
    float p_row = 0.f;
    if (EPI == EPI_DU && row_ok) p_row = __expf(g.s_raw[row] - g.ml[0]) / g.ml[1];
    constexpr int PIECES = BN / 64;   // 32-column pieces per half
    // EPI_DU: the ReLU-mask source H is prefetched one piece ahead, the first piece while the
    // mainloop still runs (the loads were the epilogue's dominant stall in ncu: 20 % of all samples
    // sat on the first use of an H register)
    uint4 hpre[4];
    if (EPI == EPI_DU && row_ok) {
      const uint4* hsrc = reinterpret_cast<const uint4*>(g.H + (long long)row * g.ldh + n0 + half * PIECES * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) hpre[i] = __ldg(hsrc + i);
    }
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
#pragma unroll 1
    for (int ii = 0; ii < PIECES; ++ii) {
      const int cb = half * PIECES + ii;
      const int col0 = n0 + cb * 32;
      if (col0 >= g.N) break;
      float v[32];
      tmem_ld32(tmem + ((q * 32u) << 16) + cb * 32, v);
      uint4 hcur[4];
      if (EPI == EPI_DU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) hcur[i] = hpre[i];
        if (row_ok && ii + 1 < PIECES && col0 + 32 < g.N) {
          const uint4* hsrc = reinterpret_cast<const uint4*>(g.H + (long long)row * g.ldh + col0 + 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) hpre[i] = __ldg(hsrc + i);
        }
      }
      tmem_ld_wait();
      if (EPI == EPI_STORE) {
        if (g.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (row_ok) {
          if (g.c_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.c_bf16) +
                                                  (long long)row * g.ldc + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                  pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          } else {
            float4* dst = reinterpret_cast<float4*>(g.c_f32 + (long long)row * g.ldc + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      } else if (EPI == EPI_ATOMIC) {
        if (row_ok) {
          float* dst = g.c_f32 + (long long)row * g.ldc + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]),
                         "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      } else {  // EPI_DU
        if (row_ok) {
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t hw[4] = {hcur[i].x, hcur[i].y, hcur[i].z, hcur[i].w};
            const float4 dm0 = __ldg(reinterpret_cast<const float4*>(g.dM + col0 + 8 * i));
            const float4 dm1 = __ldg(reinterpret_cast<const float4*>(g.dM + col0 + 8 * i + 4));
            const float dmv[8] = {dm0.x, dm0.y, dm0.z, dm0.w, dm1.x, dm1.y, dm1.z, dm1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 hf = unpack_bf16x2(hw[j]);
              const int c = 8 * i + 2 * j;
              v[c] = hf.x > 0.f ? g.du_scale * fmaf(p_row, dmv[2 * j], v[c]) : 0.f;
              v[c + 1] = hf.y > 0.f ? g.du_scale * fmaf(p_row, dmv[2 * j + 1], v[c + 1]) : 0.f;
              packed[4 * i + j] = pack_bf16x2(v[c], v[c + 1]);
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.c_bf16) +
                                                (long long)row * g.ldc + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        const float cs = warp_colsum32(v);
        g.colsum_ws[((long long)tile128 * 4 + q) * g.N + col0 + lane] = cs;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem, BN);
}

}  // namespace mmf
