// train_glue.cuh — the training-step glue either side of the hot path: fused multi-tensor
// Adam (+ weight decay, + the l1_reg_all penalty and its sign gradient) and the concordance index on the device.
//
//   Adam      torch.optim.Adam as the reference builds it (utils/utils.py:144-151: lr, weight_decay = --reg, default
//             betas / eps, no amsgrad), one launch for every parameter tensor of the model instead of ~10 ATen
//             kernels per tensor; the reference's l1 regulariser (utils/utils.py:249-257: sum |W| over all parameters,
//             added to the loss as lambda * sum, utils/core_utils.py:218-221,242) is folded in: its gradient
//             lambda * sign(W) joins the gradient and the penalty value is accumulated for logging.
//   c-index   sksurv.concordance_index_censored(event, time, risk, tied_tol) as called at utils/core_utils.py:258:
//             comparable pairs (event_i, t_i < t_j); concordant when risk_i > risk_j; |risk_i - risk_j| <= tied_tol
//             counts half. O(B^2) pair grid, integer counts (exact).
#pragma once
#include <stdint.h>

namespace mmf {

constexpr int ADAM_MAX_TENSORS = 48;
struct AdamTensors {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long start[ADAM_MAX_TENSORS + 1];   // prefix sum of element counts: a flat index space over all tensors
  int n;
};
struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay, grad_scale, l1_lambda;
  float bias1, bias2_sqrt;   // 1 - beta1^t, sqrt(1 - beta2^t)
  int zero_grad;             // also clear g (optimizer.zero_grad() fused)
  const unsigned long long* step_dev;   // non-null: the 1-based step count lives on the device (graph-captured training
                                        // step, mmf_step_state_advance); the bias corrections are then formed in the kernel
};

// Device-resident state of a graph-captured training step: word 0 = optimizer step count, words 1.. = dropout seeds.
// One launch at the head of every replay: the count goes up by one and every seed moves to the next value of its own
// splitmix64 sequence, so that each replay of the SAME graph runs with the bias corrections and dropout masks an eager
// loop would have drawn host-side.
__device__ __forceinline__ unsigned long long splitmix64_next(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (z ^ (z >> 31)) & 0x3FFFFFFFFFFFFFFFull;   // seeds stay below 2^62 like the host-drawn ones
}
__global__ void step_state_advance_kernel(unsigned long long* __restrict__ state, int n_seeds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) state[0] += 1ull;
  else if (i <= n_seeds) state[i] = splitmix64_next(state[i]);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamTensors T, const AdamHyper h, float* __restrict__ l1_out) {
  __shared__ float s_red[8];
  __shared__ float s_bias[2];
  float bias1 = h.bias1, bias2_sqrt = h.bias2_sqrt;
  if (h.step_dev != nullptr) {
    if (threadIdx.x == 0) {
      const double t = (double)*h.step_dev;
      s_bias[0] = (float)(1.0 - pow((double)h.beta1, t));
      s_bias[1] = (float)sqrt(1.0 - pow((double)h.beta2, t));
    }
    __syncthreads();
    bias1 = s_bias[0]; bias2_sqrt = s_bias[1];
  }
  const long long total = T.start[T.n];
  float l1 = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int k = 0;   // tensor of flat index i (linear scan: n <= 48, cached constants)
    while (i >= T.start[k + 1]) ++k;
    const long long j = i - T.start[k];
    const float p = T.p[k][j];
    float g = T.g[k][j] * h.grad_scale;
    if (h.l1_lambda != 0.f) g += h.l1_lambda * (p > 0.f ? 1.f : (p < 0.f ? -1.f : 0.f));
    l1 += fabsf(p);
    g = fmaf(h.weight_decay, p, g);
    const float m = fmaf(h.beta1, T.m[k][j], (1.f - h.beta1) * g);
    const float v = fmaf(h.beta2, T.v[k][j], (1.f - h.beta2) * g * g);
    T.m[k][j] = m;
    T.v[k][j] = v;
    const float denom = sqrtf(v) / bias2_sqrt + h.eps;
    T.p[k][j] = p - (h.lr / bias1) * (m / denom);
    if (h.zero_grad) const_cast<float*>(T.g[k])[j] = 0.f;
  }
  if (l1_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = l1;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += s_red[i];
      atomicAdd(l1_out, t);
    }
  }
}

// counts[0] concordant, [1] discordant, [2] tied in risk, over the comparable pairs of
// sksurv.metrics.concordance_index_censored (scikit-survival 0.2x, metrics.py:_get_comparable): sample i with an event
// is comparable to every j with t_j > t_i AND to every j censored at the SAME time t_j == t_i (an event at t counts
// as earlier than a censoring at t). Discretised survival times (months, days) make such ties common.
__global__ void __launch_bounds__(256) cindex_pairs_kernel(const float* __restrict__ risk, const float* __restrict__ times,
                                                           const float* __restrict__ event, int B, float tied_tol,
                                                           unsigned long long* __restrict__ counts) {
  __shared__ unsigned long long s_c[3];
  if (threadIdx.x < 3) s_c[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned int conc = 0, disc = 0, tied = 0;
  const int i = blockIdx.x;
  if (event[i] != 0.f) {
    const float ti = times[i], ri = risk[i];
    for (int j = threadIdx.x; j < B; j += 256) {
      const float tj = times[j];
      if (tj > ti || (tj == ti && event[j] == 0.f)) {
        const float d = ri - risk[j];
        if (fabsf(d) <= tied_tol) ++tied;
        else if (d > 0.f) ++conc;
        else ++disc;
      }
    }
  }
  atomicAdd(&s_c[0], (unsigned long long)conc);
  atomicAdd(&s_c[1], (unsigned long long)disc);
  atomicAdd(&s_c[2], (unsigned long long)tied);
  __syncthreads();
  if (threadIdx.x < 3 && s_c[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_c[threadIdx.x]);
}

// Attention scores -> percentiles, as utils/wsi_utils.py:171-174 (to_percentiles) and utils/heatmap_utils.py:32-34,99,138
// compute them with scipy.stats.percentileofscore(ref, x) (kind = 'rank'): (left + right + [left < right]) * 50 / n with
// left = #{ref < x}, right = #{ref <= x}.
// The reference runs one O(n) scipy call per patch (O(N n) on the host); here one thread per query counts over
// shared-memory tiles of the reference scores (exact, ties included).
__global__ void __launch_bounds__(256) percentile_of_score_kernel(const float* __restrict__ ref, int n_ref,
                                                                  const float* __restrict__ query, int n_query,
                                                                  float* __restrict__ out) {
  __shared__ float tile[2048];
  const int q = blockIdx.x * 256 + threadIdx.x;
  const float x = q < n_query ? query[q] : 0.f;
  unsigned int less = 0, leq = 0;
  for (int t0 = 0; t0 < n_ref; t0 += 2048) {
    const int nt = min(2048, n_ref - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += 256) tile[i] = ref[t0 + i];
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < nt; ++i) {
      const float r = tile[i];
      less += r < x;
      leq += r <= x;
    }
  }
  // scipy (kind='rank'): (left + right + [left < right]) * 50 / n
  if (q < n_query) out[q] = (float)(less + leq + (less < leq ? 1u : 0u)) * (50.f / (float)n_ref);
}

}  // namespace mmf
