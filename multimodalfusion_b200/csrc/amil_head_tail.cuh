// amil_head_tail.cuh — the training step's head (combine -> classifier -> hazards -> nll_surv -> dlogits), folded
// into the PROLOGUE of the hidden-gradient kernel (amil_hidden_fused.cuh, HEADPROJ form).
//
// The forward's tile CTAs emit, next to their softmax partial (m_t, l_t, acc_t[L]), a 12-float head row
// (m_t, l_t, -, -, Wk·acc_t [8]): with a linear classifier directly on the pooled embedding,
//   logits - bk = Wk·M = sum_t e^{m_t - m} (Wk·acc_t) / l,     l = sum_t e^{m_t - m} l_t,
// so every CTA of the backward kernel merges <= 512 rows of 48 bytes (one row per worker thread, one L2 round trip,
// three CTA barriers), forms the hazards / survival function (models/model_attention_mil_path.py:55-61), the
// nll_surv loss (utils/loss_utils.py:22-39) and its gradient w.r.t. the logits — redundantly, in parallel; CTA 0
// writes the step's outputs. The pooled embedding M itself (an output, and the operand of dWk += dlogits (x) M) is
// formed off the critical path by the otherwise idle warp 1 of the non-leader CTAs from the (L + 2)-float partials.
// History (gpurun_out/r2_phase*.log): a separate head kernel cost 7.7 us + a 3-4 us dependent-kernel boundary (round 1);
// folding the head into the LAST tile CTA of the forward cost 35-53k cycles on that CTA (cold code, serial L2 round
// trips); merging the full partial rows in every backward CTA, two levels with group flags, 14.5k cycles.
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
constexpr int HEAD_MAX_TILES = 512;  // one head row per worker thread of the backward kernel (N <= 65536); larger bags use the head kernel
constexpr int HEAD_ROW = 12;         // floats per head row
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

}  // namespace mmf
