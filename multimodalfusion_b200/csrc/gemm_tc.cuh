// gemm_tc.cuh — single-CTA warp-specialised tcgen05 GEMM (operand majors and epilogue are template parameters) and
// the argument structs shared with the CTA-pair kernel (gemm2_tc.cuh). Instantiated for
//
//   (A K-major , B K-major, EPI_STORE)  y = x W^T + b      radio reduce_dim forward (A as up to 4 K-segments)
//
// (the backward GEMMs — dU, dx, split-K weight gradients — run on the CTA-pair kernel).
//
// Tile 128 x 256 x 64, 4-stage TMA ring (48 KB / stage), accumulator in 256 TMEM columns.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue
// (TMEM lane quadrant = warp % 4).
#pragma once
#include "mmf_ptx.cuh"

namespace mmf {

struct TMapSet {
  CUtensorMap m[4];
};

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_DU = 2 };

struct GemmArgs {
  int M, N;          // C is [M, N]
  int kb_total;      // number of 64-wide k-blocks over the whole reduction
  int kb_per_split;  // k-blocks handled by one grid.z slice (split-K)
  int a_seg_kb;      // K-major A given as K-segments: k-blocks per segment map
  int b_seg_n;       // MN-major B given as N-segments: columns per segment map
  float* c_f32;      // EPI_STORE (fp32 out) / EPI_ATOMIC target
  void* c_bf16;      // EPI_STORE / EPI_DU bf16 out
  long long ldc;
  const float* bias; // EPI_STORE: [N] or null
  // EPI_DU extras
  const float* s_raw;        // [M] raw attention scores
  const float* ml;           // (m, l) global softmax statistics
  const float* dM;           // [N] gradient w.r.t. the pooled vector
  const __nv_bfloat16* H;    // [M, ldh] post-ReLU activations (mask source)
  long long ldh;
  float* colsum_ws;          // [gridDim.x * 4, N] per-warp column sums of the stored tile
  float du_scale;            // 1, or 1/(1-p) when train-mode dropout was applied to h
  // pair-kernel EPI_DU (gemm2_tc.cuh): 1-bit ReLU mask instead of H, direct db1 accumulation; the
  // output tensor map travels in tmA.m[3]
  const uint32_t* mask;      // [M, mask_ld] words, bit j of word w = (h[row, 32w + j] > 0)
  long long mask_ld;
  float* db1;                // [N] accumulated with atomicAdd
  // pair-kernel grouped split-K mode (EPI_ATOMIC, both operands MN-major): n_groups > 0 packs up to two
  // independent problems C_p += A_p^T B_p over the same reduction axis into ONE single-wave launch;
  // grid.x = 2 * sum_p (tiles_m * tiles_n * splits), grid.y = grid.z = 1
  unsigned long long* dbg;   // optional [gridDim.x, 16] clock64 phase stamps (mmf_debug_set_timing_buffer)
  int n_groups;
  struct Group {
    int M, N;                // C_p is [M, N] (or [N, M] when trans)
    int tiles_n, splits;     // tiles along N; split-K slices per tile
    int kb_per_split;
    int a_map, b_map;        // indices into tmA.m[] / tmB.m[]
    int first_pair;          // first CTA-pair index of this group
    int b_stream;            // the B operand streams through L2 once (the bag x): TMA loads carry evict_first
    int trans;               // store C^T: element (row, col) goes to c[col * ldc + row]
    int tma_reduce;          // reduce through the TMA (cp.reduce.async.bulk.tensor .add) — output map in
                             // tmA.m[3] (group 0) / tmB.m[3] (group 1), fp32 box [32 rows][32 cols]
    float* c;
    long long ldc;
  } grp[2];
};

constexpr int GEMM_BM = 128, GEMM_BN = 256, GEMM_BK = 64, GEMM_STAGES = 4;
constexpr uint32_t GEMM_A_BYTES = GEMM_BM * 128;             // 16 KB
constexpr uint32_t GEMM_B_BYTES = GEMM_BN * 128;             // 32 KB
constexpr uint32_t GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr uint32_t GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024;  // + alignment slack

template <int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ TMapSet tmA, const __grid_constant__ TMapSet tmB,
               const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[GEMM_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[GEMM_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pool = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int m_tile = blockIdx.x, n_tile = blockIdx.y;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;  // host guarantees nkb >= 1

  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------- TMA producer -------------------------------------
    for (int i = 0; i < nkb; ++i) {
      const int kb = kb0 + i;
      const int s = i % GEMM_STAGES;
      const uint32_t ph = (i / GEMM_STAGES) & 1;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      const uint32_t full = smem_u32(&bar_full[s]);
      const uint32_t a_dst = pool + s * GEMM_STAGE_BYTES;
      const uint32_t b_dst = a_dst + GEMM_A_BYTES;
      mbar_arrive_expect_tx(full, GEMM_STAGE_BYTES);
      if (A_MN == 0) {
        const int seg = kb / g.a_seg_kb;
        tma_load_2d(a_dst, &tmA.m[seg], full, (kb - seg * g.a_seg_kb) * 64, m_tile * GEMM_BM);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_2d(a_dst + j * 8192, &tmA.m[0], full, m_tile * GEMM_BM + j * 64, kb * 64);
      }
      if (B_MN == 0) {
        tma_load_2d(b_dst, &tmB.m[0], full, kb * 64, n_tile * GEMM_BN);
      } else {
        const int n0 = n_tile * GEMM_BN;
        const int seg = n0 / g.b_seg_n;
        const int c0 = n0 - seg * g.b_seg_n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_2d(b_dst + j * 8192, &tmB.m[seg], full, c0 + j * 64, kb * 64);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------- MMA issuer ---------------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GEMM_BN, A_MN, B_MN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % GEMM_STAGES;
      const uint32_t ph = (i / GEMM_STAGES) & 1;
      mbar_wait(smem_u32(&bar_full[s]), ph);
      tc_fence_after();
      const uint32_t a_src = pool + s * GEMM_STAGE_BYTES;
      const uint32_t b_src = a_src + GEMM_A_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = A_MN ? umma_desc_sw128(a_src + k * 2048, 8192, 1024)
                                 : umma_desc_sw128(a_src + k * 32, 16, 1024);
        const uint64_t bd = B_MN ? umma_desc_sw128(b_src + k * 2048, 8192, 1024)
                                 : umma_desc_sw128(b_src + k * 32, 16, 1024);
        umma_bf16_ss(tmem, ad, bd, idesc, (i | k) != 0);
      }
      umma_commit(smem_u32(&bar_empty[s]));
    }
    umma_commit(smem_u32(&bar_acc));
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------------
    const uint32_t q = warp & 3;
    const int row = m_tile * GEMM_BM + q * 32 + lane;
    const bool row_ok = row < g.M;
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();

    float p_row = 0.f;
    if (EPI == EPI_DU && row_ok) {
      const float m = g.ml[0], l = g.ml[1];
      p_row = __expf(g.s_raw[row] - m) / l;
    }
#pragma unroll 1
    for (int cb = 0; cb < GEMM_BN / 32; ++cb) {
      const int col0 = n_tile * GEMM_BN + cb * 32;
      if (col0 >= g.N) break;  // uniform across the CTA
      float v[32];
      tmem_ld32(tmem + ((q * 32u) << 16) + cb * 32, v);
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
          const uint4* hsrc = reinterpret_cast<const uint4*>(g.H + (long long)row * g.ldh + col0);
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 hv = __ldg(hsrc + i);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 hf = unpack_bf16x2(hw[j]);
              const int c = 8 * i + 2 * j;
              const float2 dm = __ldg(reinterpret_cast<const float2*>(g.dM + col0 + c));
              v[c] = hf.x > 0.f ? g.du_scale * fmaf(p_row, dm.x, v[c]) : 0.f;
              v[c + 1] = hf.y > 0.f ? g.du_scale * fmaf(p_row, dm.y, v[c + 1]) : 0.f;
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
        // column sums of the (fp32, pre-rounding) tile -> db1 partials
        const float cs = warp_colsum32(v);
        g.colsum_ws[((long long)m_tile * 4 + q) * g.N + col0 + lane] = cs;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

// Column sums of up to three row-major partial-sum workspaces in ONE launch:
//   seg k: out_k[c] += sum_r ws_k[r * ld_k + c],  c < ncols_k,  r < rows
// grid = (ceil(total_cols / 32), row_splits); block (32, 8); one atomicAdd per (block, column).
struct ReduceSeg {
  const float* ws;
  long long ld;
  int ncols;
  float* out;
};
struct ReduceSegs {
  ReduceSeg s[3];
  int n;
};
__global__ void reduce_rows_multi_kernel(const ReduceSegs segs, long long rows) {
  __shared__ float part[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  const float* ws = nullptr;
  long long ld = 0;
  float* out = nullptr;
  for (int k = 0; k < segs.n; ++k) {
    const int padded = (segs.s[k].ncols + 31) & ~31;
    if (c < padded) {
      if (c < segs.s[k].ncols) { ws = segs.s[k].ws + c; ld = segs.s[k].ld; out = segs.s[k].out + c; }
      break;
    }
    c -= padded;
  }
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  if (ws)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += ws[r * ld];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && out) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x];
    atomicAdd(out, t);
  }
}

inline void launch_reduce_rows(const ReduceSegs& segs, long long rows, cudaStream_t st) {
  int blocks_x = 0;
  for (int k = 0; k < segs.n; ++k) blocks_x += (segs.s[k].ncols + 31) / 32;
  int splits = (int)((rows + 63) / 64);
  if (splits < 1) splits = 1;
  if (splits > 16) splits = 16;
  reduce_rows_multi_kernel<<<dim3(blocks_x, splits), dim3(32, 8), 0, st>>>(segs, rows);
}

}  // namespace mmf
