// amil_head_tail.cuh — the training step's head (combine -> classifier -> hazards -> nll_surv -> dlogits), folded
// into the PROLOGUE of the hidden-gradient kernel (amil_hidden_fused.cuh, HEADPROJ form).
//
// Every CTA of that kernel merges the forward's per-tile softmax partials itself (<= 256 rows of L + 2 floats from
// L2, one column pair per thread), forms the pooled embedding M, the logits, the hazards / survival function
// (models/model_attention_mil_path.py:55-61), the nll_surv loss (utils/loss_utils.py:22-39) and its gradient w.r.t.
// the logits — redundantly, in parallel, while the first activation tiles are in flight; CTA 0 writes the step's
// outputs. No separate head launch (7.7 us + a 3-4 us dependent-kernel boundary in round 1) and no serial tail.
// (Round-2 measurement, gpurun_out/r2_phase2.log / r2_phase3.log: folding the head into the LAST tile CTA of the
// forward — atomic ticket, one- or two-level — cost 35-53k cycles on that CTA: its code is cold in the instruction
// cache and every phase is a dependent L2 round trip, serialised behind the slowest tile.)
//
// Head-projected backward (oracle.amil_backward_head_projected): with a linear classifier directly on M,
// dM = Wk^T dlogits, hence t_i = dM·h_i = dlogits·(Wk h_i) = dlogits·z_i with z_i emitted by the forward's tensor
// cores, and dM·M = dlogits·(logits - bk).
//
// nll_surv gradient in closed form (h = sigmoid(logit), S(k) = prod_{i<=k} (1 - h_i), d log S(k) / d logit_j = -h_j for
// j <= k, d log h_y / d logit_y = 1 - h_y; a clamp(min = eps) that is active passes no gradient):
//   loss = -(1-c) [log max(S(y-1), eps) + log max(h_y, eps)] - (1-alpha) c log max(S(y), eps)
//   dloss/dlogit_j = (1-c) ( h_j [j <= y-1][S(y-1) >= eps] - (1 - h_y) [j == y][h_y >= eps] )
//                    + (1-alpha) c h_j [j <= y][S(y) >= eps]
#pragma once
#include <math_constants.h>

#include "mmf_ptx.cuh"

namespace mmf {

constexpr int HEAD_MAX_K = 8;        // classes (hi + lo bf16 split of Wk fills the N = 16 side MMA)
constexpr int HEAD_GROUP = 16;       // tile partials per first-level merge group
constexpr int HEAD_MAX_GROUPS = 16;  // group rows every CTA merges itself
constexpr int HEAD_MAX_TILES = HEAD_GROUP * HEAD_MAX_GROUPS;  // = 256 tiles (N <= 32768); larger bags use the head kernel
// head scalars left in global memory: hs[0..K) = dlogits (already scaled by loss_scale), hs[HS_DOT] = dM·M
constexpr int HS_DOT = 8, HS_WORDS = 16;

struct HeadTail {
  const float* Wk;        // [K, L] classifier.weight (fp32)
  const float* bk;        // [K]
  const long long* Y;     // [1] discrete time bin
  const float* c;         // [1] censorship
  float alpha, eps, loss_scale;
  int K;
  float* M;               // [L]   out
  float* ml;              // [2]   out (m, l)
  float* hazards;         // [K]   out
  float* S;               // [K]   out
  long long* Y_hat;       // [1]   out (or null)
  float* loss;            // [1]   out (unscaled loss value)
  float* dM;              // [L]   out  = loss_scale * dLoss/dM
  float* hs;              // [HS_WORDS] out
  float* dWk;             // [K, L] accumulated (or null)
  float* dbk;             // [K]    accumulated (or null)
};

// Online-softmax merge of the partial rows r0, r0 + rstep, ... < count (row = (m_t, l_t, acc_t[L])) for the column
// pair `cp`, entirely per thread: 8 rows per batch = 16 independent 8-byte loads in flight (every thread reads the
// (m_t, l_t) pairs itself: same addresses across the CTA, no shared memory, no barrier). Loads go to L2 (ld.global.cg):
// the group rows are written by other CTAs of the same launch.
struct PoolAcc { float m, l, ax, ay; };
// NOT inlined and only 8 rows per batch: the head runs once per CTA, its cost is instruction fetch (cold code streams
// from L2 at ~50-100 cycles per 128-byte line: the first version — three inlined 16-row copies and an 8-way unrolled
// scalar section, ~1600 instructions — took 19k cycles for ~3 L2 round trips of data, gpurun_out/r2_phase8.log).
template <int L>
__device__ __noinline__ PoolAcc combine_rows_online(const float* rows, int count, int cp, int r0, int rstep) {
  constexpr int UB = 8;
  const long long stride = L + 2;
  PoolAcc r = {-CUDART_INF_F, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int b0 = r0; b0 < count; b0 += UB * rstep) {
    float2 mlv[UB], v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int t = b0 + u * rstep;
      const bool ok = t < count;
      mlv[u] = ok ? __ldcg(reinterpret_cast<const float2*>(rows + t * stride)) : make_float2(-CUDART_INF_F, 0.f);
      v[u] = ok ? __ldcg(reinterpret_cast<const float2*>(rows + t * stride + 2 + 2 * cp)) : make_float2(0.f, 0.f);
    }
    float m_new = r.m;
#pragma unroll
    for (int u = 0; u < UB; ++u) m_new = fmaxf(m_new, mlv[u].x);
    const float sc = (r.m > -CUDART_INF_F) ? __expf(r.m - m_new) : 0.f;
    r.l *= sc; r.ax *= sc; r.ay *= sc;
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const float w = (mlv[u].x > -CUDART_INF_F) ? __expf(mlv[u].x - m_new) : 0.f;
      r.l = fmaf(mlv[u].y, w, r.l);
      r.ax = fmaf(v[u].x, w, r.ax);
      r.ay = fmaf(v[u].y, w, r.ay);
    }
    r.m = m_new;
  }
  return r;
}

}  // namespace mmf
