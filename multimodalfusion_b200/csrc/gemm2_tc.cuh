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
#define MMF_GSTAMP(g, i) do { if ((g).dbg) (g).dbg[(long long)blockIdx.x * 16 + (i)] = clock64(); } while (0)

template <int BN>
struct Gemm2Cfg {
  static_assert(BN == 256 || BN == 512, "pair tile is 256 x 256 or 256 x 512");
  static constexpr int NH = BN / 256;
  static constexpr uint32_t A_BYTES = 16384;
  static constexpr uint32_t B_BYTES = NH * 16384u;
  static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 4;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE + 2048 + 1024;   // ring + dM vector (EPI_DU) + slack
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

  // problem decode: plain grid (pair_m, n_tile, split) or the grouped single-wave list
  int pair_m = blockIdx.x >> 1, n_tile = blockIdx.y, split = blockIdx.z;
  int gM = g.M, gN = g.N, kb_per_split = g.kb_per_split, a_map = 0, b_map = -1, trans = 0, tma_red = 0, b_stream = 0;
  const CUtensorMap* cmap = &tmA.m[3];
  float* c_f32 = g.c_f32;
  long long ldc = g.ldc;
  if (g.n_groups > 0) {
    const int pair = blockIdx.x >> 1;
    const GemmArgs::Group& P = g.grp[(g.n_groups > 1 && pair >= g.grp[1].first_pair) ? 1 : 0];
    const int local = pair - P.first_pair;
    split = local % P.splits;
    const int t = local / P.splits;
    n_tile = t % P.tiles_n;
    pair_m = t / P.tiles_n;
    gM = P.M; gN = P.N; kb_per_split = P.kb_per_split; a_map = P.a_map; b_map = P.b_map; trans = P.trans;
    tma_red = P.tma_reduce; b_stream = P.b_stream;
    if (&P != &g.grp[0]) cmap = &tmB.m[3];
    c_f32 = P.c; ldc = P.ldc;
  }
  const int m0 = pair_m * 256 + 128 * (int)rank;   // first output row of this CTA
  const int n0 = n_tile * BN;
  const int kb0 = split * kb_per_split;
  const int kb1 = min(g.kb_total, kb0 + kb_per_split);
  const int nkb = kb1 - kb0;

  Timeline tl = timeline_start(EPI == EPI_ATOMIC ? 3 : 5);
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    MMF_GSTAMP(g, 0);
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
  griddep_wait();
  // (pointers to the previous kernel's outputs are re-derived after the wait, see pdl_fresh)
  const float* const g_dM = pdl_fresh(g.dM); const float* const g_sraw = pdl_fresh(g.s_raw); const float* const g_ml = pdl_fresh(g.ml);
  const uint32_t* const g_mask = pdl_fresh(g.mask); const float* const g_bias = pdl_fresh(g.bias);
  timeline_wait_done(tl);
  if (threadIdx.x == 0) MMF_GSTAMP(g, 1);

  if (warp == 0 && lane == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------
    const uint64_t pol_stream = l2_policy_evict_first();   // (see amil_tile2.cuh: the bag must not displace the stash in L2)
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
        for (int j = 0; j < 2; ++j) tma_load_2d_pair(a_dst + j * 8192, &tmA.m[a_map], full, m0 + j * 64, kb * 64);
      }
#pragma unroll
      for (int h = 0; h < C::NH; ++h) {
        const int nb = n0 + 256 * h + 128 * (int)rank;   // this CTA's 128 columns of MMA h
        if (B_MN == 0) {
          tma_load_2d_pair(b_dst + h * 16384, &tmB.m[0], full, kb * 64, nb);
        } else {
          const int seg = b_map >= 0 ? b_map : nb / g.b_seg_n;
          const int c0 = b_map >= 0 ? nb : nb - seg * g.b_seg_n;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (b_stream) tma_load_2d_pair_hint(b_dst + h * 16384 + j * 8192, &tmB.m[seg], full, c0 + j * 64, kb * 64, pol_stream);
            else tma_load_2d_pair(b_dst + h * 16384 + j * 8192, &tmB.m[seg], full, c0 + j * 64, kb * 64);
          }
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
      if (i == 0) MMF_GSTAMP(g, 2);
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
    MMF_GSTAMP(g, 3);
  } else if (warp >= 4 && EPI == EPI_DU) {
    // ------------------------------- dU epilogue (both CTAs, 8 warps) -----------------
    //   dU = (acc + p_i dM) ⊙ [h > 0] * scale  -> bf16, staged in the (now idle) operand ring in the
    //   TMA 128B-swizzled layout and written out with one TMA store per 64-column block; db1 column
    //   sums are taken from the staged tile. The ReLU mask comes as 1 bit per element (32 B per row
    //   and thread instead of 512 B of H): the first version read H with per-thread 16 B loads and
    //   stored dU the same way — its epilogue took ~35k of the kernel's 56k cycles.
    const uint32_t q = warp & 3;
    const uint32_t half = (warp - 4) >> 2;
    const uint32_t e = threadIdx.x - 128;
    const uint32_t r = q * 32 + lane;
    const int row = m0 + (int)r;
    const bool row_ok = row < g.M;
    constexpr int PIECES = BN / 64;   // 32-column pieces per half
    float* s_dm = reinterpret_cast<float*>(smem_raw + (pool - smem_u32(smem_raw)) + C::STAGES * C::STAGE);
    for (int i = e; i < BN; i += 256) s_dm[i] = (n0 + i < g.N) ? __ldg(g_dM + n0 + i) : 0.f;
    float p_row = 0.f;
    uint32_t mw[PIECES];
#pragma unroll
    for (int i = 0; i < PIECES; ++i) mw[i] = 0u;
    if (row_ok) {
      p_row = __expf(g_sraw[row] - g_ml[0]) / g_ml[1];
      const uint32_t* mp = g_mask + (long long)row * g.mask_ld + (n0 >> 5) + half * PIECES;
#pragma unroll
      for (int i = 0; i < PIECES; ++i) mw[i] = __ldg(mp + i);
    }
    named_bar_sync(1, 256);           // s_dm visible
    mbar_wait(smem_u32(&bar_acc), 0);  // every MMA retired: accumulator complete, operand ring idle
    tc_fence_after();
    if (e == 0) MMF_GSTAMP(g, 4);
    float v[2][32];
    tmem_ld32(tmem + ((q * 32u) << 16) + half * PIECES * 32, v[0]);
#pragma unroll
    for (int ii = 0; ii < PIECES; ++ii) {
      const int cb = half * PIECES + ii;
      tmem_ld_wait();
      if (ii + 1 < PIECES) tmem_ld32(tmem + ((q * 32u) << 16) + (cb + 1) * 32, v[(ii + 1) & 1]);
      float (&u)[32] = v[ii & 1];
      const uint32_t bits = mw[ii];
      const float4* dm4 = reinterpret_cast<const float4*>(s_dm + cb * 32);
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 d = dm4[i >> 2];
        const float o0 = (bits >> i) & 1u ? g.du_scale * fmaf(p_row, d.x, u[i]) : 0.f;
        const float o1 = (bits >> (i + 1)) & 1u ? g.du_scale * fmaf(p_row, d.y, u[i + 1]) : 0.f;
        const float o2 = (bits >> (i + 2)) & 1u ? g.du_scale * fmaf(p_row, d.z, u[i + 2]) : 0.f;
        const float o3 = (bits >> (i + 3)) & 1u ? g.du_scale * fmaf(p_row, d.w, u[i + 3]) : 0.f;
        packed[i >> 1] = pack_bf16x2(o0, o1);
        packed[(i >> 1) + 1] = pack_bf16x2(o2, o3);
      }
      const uint32_t blk = pool + (cb >> 1) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(blk + sw128_offset(r, (cb & 1) * 4 + j), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2],
                     packed[4 * j + 3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(1, 256);           // the staged tile is complete
    if (e == 0) {
      for (int kb = 0; kb < BN / 64; ++kb)
        if (n0 + kb * 64 < g.N) tma_store_2d(&tmA.m[3], pool + kb * 16384, n0 + kb * 64, m0);
      tma_store_commit();
    }
    // db1: column sums of the staged bf16 tile (rows past M were staged as zeros: their mask words are 0)
    for (uint32_t cp = e; cp < (uint32_t)BN / 2; cp += 256) {
      const uint32_t col = 2u * cp;
      const uint32_t blk = pool + (col >> 6) * 16384u, chunk = (col & 63u) >> 3, inb = (col & 7u) * 2u;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
      for (uint32_t rr = 0; rr < 128; ++rr) {
        const float2 f = unpack_bf16x2(ld_shared_b32(blk + sw128_offset(rr, chunk) + inb));
        a0 += f.x; a1 += f.y;
      }
      if (n0 + (int)col < g.N) { atomicAdd(g.db1 + n0 + col, a0); atomicAdd(g.db1 + n0 + col + 1, a1); }
    }
    if (e == 0) { tma_store_wait_exit(); MMF_GSTAMP(g, 5); }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (both CTAs, 8 warps) --------------------
    const uint32_t q = warp & 3;
    const uint32_t half = (warp - 4) >> 2;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < gM;
    constexpr int PIECES = BN / 64;   // 32-column pieces per half
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
    if (threadIdx.x == 128) MMF_GSTAMP(g, 4);
    if (EPI == EPI_ATOMIC && tma_red) {
      // split-K reduction through the TMA: each warp stages its 32 x 32 fp32 block in the idle operand ring
      // (128-byte rows, 128B swizzle, two 4 KB buffers per warp) and issues one cp.reduce.async.bulk.tensor
      // .add per block: full 128-byte lines reach the L2 instead of 32 scattered 16-byte red.global per
      // warp instruction (the red epilogue took 18k of the wgrad kernel's 51k cycles).
      const uint32_t wbuf = pool + (warp - 4) * 8192u;
      const int row_w = m0 + (int)q * 32;
#pragma unroll 1
      for (int ii = 0; ii < PIECES; ++ii) {
        const int cb = half * PIECES + ii;
        const int col0 = n0 + cb * 32;
        if (col0 >= gN || row_w >= gM) break;
        float v[32];
        tmem_ld32(tmem + ((q * 32u) << 16) + cb * 32, v);
        const uint32_t buf = wbuf + (ii & 1) * 4096u;
        if (ii >= 2) {
          if (lane == 0) bulk_store_wait_read<1>();
          __syncwarp();
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(buf + sw128_offset(lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                       __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(cmap, buf, col0, row_w);
          tma_store_commit();
        }
      }
      if (lane == 0) tma_store_wait_exit();
    } else
#pragma unroll 1
    for (int ii = 0; ii < PIECES; ++ii) {
      const int cb = half * PIECES + ii;
      const int col0 = n0 + cb * 32;
      if (col0 >= gN) break;
      float v[32];
      tmem_ld32(tmem + ((q * 32u) << 16) + cb * 32, v);
      tmem_ld_wait();
      if (EPI == EPI_STORE) {
        if (g_bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g_bias + col0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (row_ok) {
          if (g.c_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.c_bf16) +
                                                  (long long)row * ldc + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                  pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          } else {
            float4* dst = reinterpret_cast<float4*>(c_f32 + (long long)row * ldc + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      } else {  // EPI_ATOMIC
        if (row_ok && !trans) {
          float* dst = c_f32 + (long long)row * ldc + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]),
                         "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        } else if (row_ok) {
          // transposed store: the 32 lanes of a warp (consecutive rows) hit 32 consecutive floats
          float* dst = c_f32 + (long long)col0 * ldc + row;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < gN) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)i * ldc), "f"(v[i]) : "memory");
        }
      }
    }
    if (threadIdx.x == 128) MMF_GSTAMP(g, 5);
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) MMF_GSTAMP(g, 6);
  timeline_end(EPI == EPI_ATOMIC ? 3 : 5, tl);
  if (warp == 2) tmem_dealloc_pair(tmem, BN);
}

}  // namespace mmf
