// amil_tile.cuh — shared definitions of the fused gated attention-MIL tile kernel (amil_tile2.cuh): the kernel's
// argument block, the counter-based dropout bits (shared with the oracle) and the softmax-partial combine kernel.
//
// Math follows SURVEY.md Appendix A.1/A.2, i.e. the reference ops
//   models/model_attention_mil_path.py:20-21,29,52-56 and models/model_modules.py:84-85,105-110:
//   GEMM1  U[128,L] = X_tile[128,1024] W1^T;  EPI1  H = dropout(relu(U + b1)) -> bf16 in shared memory (GEMM2's A operand)
//   GEMM2  [A|G]_c  = H [Wa;Wb]_c^T per 128-wide chunk of D;  EPI2  s_i = sum_d wc_d tanh(A+ba) sigmoid(G+bb) (+ bc)
//   FWD :  tile softmax partial (m, l, sum e^{s-m} h) from the resident H tile; A_raw[N] out
//   BWD :  (recompute mode) ds_i = p_i (dM.h_i - dM.M) + dA_raw_i; dG -> HBM (bf16), H tile -> HBM, column sums
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
  uint8_t* discard_ptr;    // fwd train (optional): a DEAD region of the workspace (the previous step's dU: read by its wgrad,
  long long discard_n128;  // rewritten by this step's hidden-gradient kernel) whose dirty L2 lines are dropped, in 128-byte lines
  uint8_t* h_stash;        // fwd train (optional, MMF_DISCARD_DEAD_STASH): base of the H stash [N, L] bf16 — every CTA drops the
                           // dead lines of ITS OWN tile's rows of H and [a|g] (it rewrites exactly those rows later)
  uint32_t* mask_out; // fwd train (optional): ReLU mask words [N, L/32], bit j of word w = (h[row][32 w + j] > 0)
  float* z_out;       // fwd train (optional): z_i = Wk h_i as fp32 [N, zld] from the N = 16 side MMA (needs tmWk)
  int zld;            // 4 or 8
  float* tile_head;   // fwd train (optional): [tiles, 12] per-tile head rows (m_t, l_t, -, -, Wk·acc_t [8]): the backward's head
                      // merges these 48-byte rows instead of the (L + 2)-float partials (needs head_wk / head_k)
  const float* head_wk;  // classifier.weight fp32 [head_k, L]
  int head_k;
  int flags;
  const void* x_bulk;  // fwd (optional): the bag itself when its rows are contiguous (ldx == row width): this CTA's tile is then
                      // one contiguous block that a single cp.async.bulk.prefetch.L2 requests before griddepcontrol.wait
  int kb1;            // 64-wide k-blocks of GEMM1 (0 = 16: x is [N,1024]; 48: [N,3072] = [x_hi | x_lo | x_hi], MMF_PRECISE_FC)
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
// A seed argument with bit 63 set (MMF_SEED_DEVICE(ptr), include/mmf_b200.h) is the device address of the 64-bit seed:
// the kernels read it at run time (after griddepcontrol.wait), so that the replays of a CUDA graph that contains the
// launch see a fresh seed each — mmf_step_state_advance re-hashes the word at the start of every replay. A plain
// (non-nc, volatile, memory-clobbering) load: the word is written by a kernel earlier in the same stream.
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long seed_resolve(unsigned long long s) {
  if (s >> 63) {
    unsigned long long v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(s & 0x7FFFFFFFFFFFFFFFull) : "memory");
    return v;
  }
  return s;
}
#endif
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
// The 16 keep decisions of one hash word as a dense 16-bit mask (bit i = element i kept): 11 integer operations per
// 16 elements, after which `(mask >> i) & 1` compiles to R2P + predicated selects — per-element shift / and / compare
// on the 2-bit fields made Dropout(0.25) on h cost 3.2k of EPI1's 8.8k cycles.
__host__ __device__ __forceinline__ uint32_t drop_keep_mask16(uint32_t bits16) {
  uint32_t k = (bits16 | (bits16 >> 1)) & 0x55555555u;   // bit 2i = field i non-zero
  k = (k | (k >> 1)) & 0x33333333u;
  k = (k | (k >> 2)) & 0x0F0F0F0Fu;
  k = (k | (k >> 4)) & 0x00FF00FFu;
  return (k | (k >> 8)) & 0xFFFFu;
}
// 4-column view used by the single-CTA kernel: the 8 bits of columns 4*cg4 .. 4*cg4+3
__host__ __device__ __forceinline__ uint32_t drop_bits4(uint32_t row_state, uint32_t cg4) {
  return (drop_bits16(row_state, cg4 >> 2) >> (8u * (cg4 & 3u))) & 0xFFu;
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t bits4, uint32_t j) {
  return ((bits4 >> (2u * j)) & 3u) != 0u ? (1.0f / 0.75f) : 0.0f;
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
