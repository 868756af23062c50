// amil_head_tail.cuh — the training step's head, folded into the tail of the fused forward kernel.
//
// The last tile CTA of the forward to finish (atomic ticket) combines the per-tile softmax partials into the pooled
// embedding M, runs the discrete-hazard head (classifier -> sigmoid -> cumprod, models/model_attention_mil_path.py:
// 55-61), the nll_surv loss (utils/loss_utils.py:22-39) and their backward down to dM, dWk, dbk — what
// amil_head_step_cluster_kernel does as a separate launch. Folding it removes one launch and one dependent-kernel
// boundary (3-4 us each even with PDL) from the batch-1 step (utils/core_utils.py:200-247).
//
// It also leaves what the head-projected backward needs (oracle.amil_backward_head_projected): with a linear
// classifier directly on M, dM = Wk^T dlogits, hence t_i = dM·h_i = dlogits·(Wk h_i) = dlogits·z_i with z_i emitted
// by the forward's tensor cores, and dM·M = dlogits·(logits - bk).
#pragma once
#include <math_constants.h>

#include "mmf_ptx.cuh"

namespace mmf {

constexpr int HEAD_MAX_K = 8;        // classes (hi + lo bf16 split of Wk fills the N = 16 side MMA)
constexpr int HEAD_MAX_TILES = 4096; // softmax weights staged in shared memory
// head scalars left for the backward: hs[0..K) = dlogits (already scaled by loss_scale), hs[HS_DOT] = dM·M
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
  unsigned int* ticket;   // zero before the first launch; the kernel leaves it zero
};

// Executed by the NT "epilogue" threads (e = 0..NT-1) of the last CTA. `parts` = [n][L + 2] partial rows written by
// all CTAs of this launch (read through L2: ld.global.cg). Shared scratch: s_w[n], s_v[L * (RG > 1 ? RG : 1) + 64].
template <int L, int NT>
__device__ __forceinline__ void amil_head_tail(const HeadTail& h, const float* parts, int n, float* s_w, float* s_v,
                                               uint32_t e, uint32_t bar_id) {
  static_assert(NT == 256, "thread map assumes 256 threads");
  const uint32_t lane = e & 31u, wid = e >> 5;
  const long long stride = L + 2;
  float* s_M = s_v;                 // [L]
  float* s_red = s_v + L;           // [32]: max per warp | sum per warp | logits / dlogits
  float* s_acc = s_v + L + 64;      // [RG][L] when the rows are split over row groups
  // ---- global softmax statistics ---------------------------------------------------------------------------
  float m = -CUDART_INF_F;
  for (int t = e; t < n; t += NT) m = fmaxf(m, __ldcg(parts + t * stride));
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  named_bar_sync(bar_id, NT);
  m = s_red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) m = fmaxf(m, s_red[i]);
  float l = 0.f;
  for (int t = e; t < n; t += NT) {
    const float2 ml_t = __ldcg(reinterpret_cast<const float2*>(parts + t * stride));
    const float w = (ml_t.x > -CUDART_INF_F) ? __expf(ml_t.x - m) : 0.f;
    s_w[t] = w;
    l = fmaf(ml_t.y, w, l);
  }
  l = warp_sum(l);
  if (lane == 0) s_red[8 + wid] = l;
  named_bar_sync(bar_id, NT);        // s_w complete, warp sums visible
  l = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) l += s_red[8 + i];
  if (e == 0) { h.ml[0] = m; h.ml[1] = l; }
  // ---- combine: thread = one column pair x one row group, 32 independent 8-byte loads in flight ---------------
  constexpr int CP = L / 2;
  constexpr int RG = NT / CP;        // 1 (L = 512) or 2 (L = 256)
  static_assert(RG >= 1 && CP * RG == NT, "column pairs must tile the threads");
  const int cp = e % CP, rg = e / CP;
  const float* col = parts + 2 + 2 * cp;
  float ax = 0.f, ay = 0.f;
  constexpr int UB = 32;
  for (int t0 = rg; t0 < n; t0 += UB * RG) {
    float2 v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int t = t0 + u * RG;
      v[u] = (t < n) ? __ldcg(reinterpret_cast<const float2*>(col + t * stride)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int t = t0 + u * RG;
      const float w = (t < n) ? s_w[t] : 0.f;
      ax = fmaf(v[u].x, w, ax);
      ay = fmaf(v[u].y, w, ay);
    }
  }
  const float inv_l = 1.f / l;
  if (RG == 1) {
    s_M[2 * cp] = ax * inv_l; s_M[2 * cp + 1] = ay * inv_l;
  } else {
    s_acc[rg * L + 2 * cp] = ax; s_acc[rg * L + 2 * cp + 1] = ay;
    named_bar_sync(bar_id, NT);
    for (int c0 = e; c0 < L; c0 += NT) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < RG; ++r) v += s_acc[r * L + c0];
      s_M[c0] = v * inv_l;
    }
  }
  named_bar_sync(bar_id, NT);        // s_M complete
  for (int c0 = e; c0 < L; c0 += NT) h.M[c0] = s_M[c0];
  // ---- logits: warp j computes class j (K <= 8 = number of warps) ---------------------------------------------
  if ((int)wid < h.K) {
    float d = 0.f;
    const float* wrow = h.Wk + (long long)wid * L;
    for (int c0 = lane; c0 < L; c0 += 32) d = fmaf(s_M[c0], __ldg(wrow + c0), d);
    d = warp_sum(d);
    if (lane == 0) s_red[16 + wid] = d;   // logit - bk
  }
  named_bar_sync(bar_id, NT);
  if (e == 0) {
    const int K = h.K;
    float hz[HEAD_MAX_K], sv[HEAD_MAX_K], dh[HEAD_MAX_K], dS[HEAD_MAX_K];
    float surv = 1.f, best = -CUDART_INF_F;
    int besti = 0;
    for (int j = 0; j < K; ++j) {
      const float lg = s_red[16 + j] + __ldg(h.bk + j);
      if (lg > best) { best = lg; besti = j; }
      hz[j] = 1.f / (1.f + expf(-lg));
      surv *= (1.f - hz[j]);
      sv[j] = surv;
      h.hazards[j] = hz[j]; h.S[j] = surv;
      dh[j] = 0.f; dS[j] = 0.f;
    }
    if (h.Y_hat) *h.Y_hat = besti;
    const long long y = h.Y[0];
    const float cb = h.c[0], alpha = h.alpha, eps = h.eps;
    const float sp_y = (y == 0) ? 1.f : sv[y - 1], h_y = hz[y], sp_y1 = sv[y];
    const float unc = -(1.f - cb) * (logf(fmaxf(sp_y, eps)) + logf(fmaxf(h_y, eps)));
    const float cen = -cb * logf(fmaxf(sp_y1, eps));
    *h.loss = (1.f - alpha) * (cen + unc) + alpha * unc;
    if (y > 0 && sp_y >= eps) dS[y - 1] += -(1.f - cb) / sp_y;
    if (sp_y1 >= eps) dS[y] += -(1.f - alpha) * cb / sp_y1;
    if (h_y >= eps) dh[y] += -(1.f - cb) / h_y;
    float dot = 0.f;
    for (int j = 0; j < K; ++j) {
      float g = dh[j];
      float pre = 1.f;
      for (int i = 0; i < j; ++i) pre *= (1.f - hz[i]);
      float run = pre;
      for (int k = j; k < K; ++k) {
        if (k > j) run *= (1.f - hz[k]);
        g -= dS[k] * run;
      }
      const float dl = g * hz[j] * (1.f - hz[j]) * h.loss_scale;
      s_red[24 + j] = dl;
      h.hs[j] = dl;
      dot = fmaf(dl, s_red[16 + j], dot);   // dM·M = sum_j dlogit_j (logit_j - bk_j)
      if (h.dbk) h.dbk[j] += dl;
    }
    for (int j = K; j < HEAD_MAX_K; ++j) h.hs[j] = 0.f;
    h.hs[HS_DOT] = dot;
  }
  named_bar_sync(bar_id, NT);
  for (int c0 = e; c0 < L; c0 += NT) {
    float acc = 0.f;
    const float mv = s_M[c0];
    for (int j = 0; j < h.K; ++j) {
      const float dl = s_red[24 + j];
      acc = fmaf(dl, __ldg(h.Wk + (long long)j * L + c0), acc);
      if (h.dWk) h.dWk[(long long)j * L + c0] += dl * mv;
    }
    h.dM[c0] = acc;
  }
}

}  // namespace mmf
