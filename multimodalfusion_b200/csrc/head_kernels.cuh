// head_kernels.cuh — the remaining pieces of the pretrained-embedding fusion heads:
// BatchNorm1d (train / eval, forward + backward), the Highway gate mix, and the cross-entropy survival loss.
//   models/model_modules.py:5-27 (Highway), models/coxranking_models_pretrained.py:80-94,134-169 and
//   models/nll_models_pretrained.py:82-99 (early/late fcnn, early/late highway), utils/loss_utils.py:41-56 (ce_loss).
// All fp32, B = patients of a cohort batch (32 .. 512), F = 128 .. 768 features: latency-bound.
#pragma once
#include <stdint.h>

namespace mmf {

// Block (32 features, 8 row groups). Train: batch statistics (biased variance for normalisation, unbiased for the
// running estimate, as nn.BatchNorm1d), running stats updated in place with `momentum`; eval: running stats.
__global__ void __launch_bounds__(256)
batchnorm1d_fwd_kernel(const float* __restrict__ x, int B, int F, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float* __restrict__ run_mean, float* __restrict__ run_var,
                       int train, float momentum, float eps, float* __restrict__ y, float* __restrict__ save_mean,
                       float* __restrict__ save_invstd) {
  __shared__ float s_a[8][33], s_b[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int f = blockIdx.x * 32 + tx;
  const bool ok = f < F;
  float mean, invstd;
  if (train) {
    float s = 0.f;
    if (ok) for (int b = ty; b < B; b += 8) s += x[(long long)b * F + f];
    s_a[ty][tx] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_a[i][tx];
    mean = tot / (float)B;
    float q = 0.f;
    if (ok) for (int b = ty; b < B; b += 8) { const float d = x[(long long)b * F + f] - mean; q = fmaf(d, d, q); }
    s_b[ty][tx] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += s_b[i][tx];
    const float var_b = var / (float)B;
    invstd = rsqrtf(var_b + eps);
    if (ok && ty == 0) {
      if (save_mean) save_mean[f] = mean;
      if (save_invstd) save_invstd[f] = invstd;
      if (run_mean) run_mean[f] = (1.f - momentum) * run_mean[f] + momentum * mean;
      if (run_var) run_var[f] = (1.f - momentum) * run_var[f] + momentum * (var / (float)(B - 1));
    }
  } else {
    mean = ok ? run_mean[f] : 0.f;
    invstd = ok ? rsqrtf(run_var[f] + eps) : 0.f;
    if (ok && ty == 0) {
      if (save_mean) save_mean[f] = mean;
      if (save_invstd) save_invstd[f] = invstd;
    }
  }
  if (ok) {
    const float g = gamma ? gamma[f] : 1.f, bt = beta ? beta[f] : 0.f;
    for (int b = ty; b < B; b += 8) y[(long long)b * F + f] = fmaf((x[(long long)b * F + f] - mean) * invstd, g, bt);
  }
}

// train: dx = gamma*invstd/B * (B dy - sum dy - xhat * sum(dy xhat)); eval: dx = dy * gamma * invstd.
// dgamma += sum dy xhat, dbeta += sum dy.
__global__ void __launch_bounds__(256)
batchnorm1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int F,
                       const float* __restrict__ gamma, const float* __restrict__ save_mean,
                       const float* __restrict__ save_invstd, int train, float* __restrict__ dx,
                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_a[8][33], s_b[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int f = blockIdx.x * 32 + tx;
  const bool ok = f < F;
  const float mean = ok ? save_mean[f] : 0.f, invstd = ok ? save_invstd[f] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ok)
    for (int b = ty; b < B; b += 8) {
      const float d = dy[(long long)b * F + f];
      s1 += d;
      s2 = fmaf(d, (x[(long long)b * F + f] - mean) * invstd, s2);
    }
  s_a[ty][tx] = s1; s_b[ty][tx] = s2;
  __syncthreads();
  float sum_dy = 0.f, sum_dyx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sum_dy += s_a[i][tx]; sum_dyx += s_b[i][tx]; }
  if (!ok) return;
  if (ty == 0) {
    if (dgamma) dgamma[f] += sum_dyx;
    if (dbeta) dbeta[f] += sum_dy;
  }
  if (dx) {
    const float g = gamma ? gamma[f] : 1.f;
    const float k = g * invstd;
    for (int b = ty; b < B; b += 8) {
      const long long i = (long long)b * F + f;
      if (train) {
        const float xh = (x[i] - mean) * invstd;
        dx[i] = k * (dy[i] - (sum_dy + xh * sum_dyx) / (float)B);
      } else {
        dx[i] = k * dy[i];
      }
    }
  }
}

// Highway layer mix (models/model_modules.py:21-25): y = g * n + (1 - g) * l
__global__ void highway_mix_fwd_kernel(const float* __restrict__ g, const float* __restrict__ n, const float* __restrict__ l,
                                       long long count, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(g[i], n[i] - l[i], l[i]);
}
__global__ void highway_mix_bwd_kernel(const float* __restrict__ g, const float* __restrict__ n, const float* __restrict__ l,
                                       const float* __restrict__ dy, long long count, float* __restrict__ dg,
                                       float* __restrict__ dn, float* __restrict__ dl) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float d = dy[i], gi = g[i];
    dg[i] = d * (n[i] - l[i]);
    dn[i] = d * gi;
    dl[i] = d * (1.f - gi);
  }
}

// ce_loss (utils/loss_utils.py:41-56), forward + gradient, single block:
//   reg = -(1-c)(log(S_pad[Y] + eps) + log(max(h[Y], eps)));  ce = -c log(max(S[Y], eps)) - (1-c) log(1 - max(S[Y], eps))
//   loss = mean((1-alpha) ce + alpha reg)
__global__ void ce_surv_kernel(const float* __restrict__ haz, const float* __restrict__ S, const long long* __restrict__ Y,
                               const float* __restrict__ c, int B, int K, float alpha, float eps, float* loss,
                               float* d_haz, float* d_S) {
  __shared__ float red[32];
  float local = 0.f;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long y = Y[b];
    const float cb = c[b];
    const float* h = haz + (long long)b * K;
    const float* s = S + (long long)b * K;
    for (int k = 0; k < K; ++k) {
      if (d_haz) d_haz[(long long)b * K + k] = 0.f;
      if (d_S) d_S[(long long)b * K + k] = 0.f;
    }
    const float sp_y = (y == 0) ? 1.f : s[y - 1];
    const float h_y = h[y], s_y = s[y];
    const float sy = fmaxf(s_y, eps);
    const float reg = -(1.f - cb) * (logf(sp_y + eps) + logf(fmaxf(h_y, eps)));
    const float ce = -cb * logf(sy) - (1.f - cb) * logf(1.f - sy);
    local += (1.f - alpha) * ce + alpha * reg;
    if (d_S) {
      if (y > 0) d_S[(long long)b * K + y - 1] += alpha * (-(1.f - cb) / (sp_y + eps)) * invB;
      if (s_y >= eps) d_S[(long long)b * K + y] += (1.f - alpha) * (-cb / sy + (1.f - cb) / (1.f - sy)) * invB;
    }
    if (d_haz && h_y >= eps) d_haz[(long long)b * K + y] += alpha * (-(1.f - cb) / h_y) * invB;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *loss = t * invB;
  }
}

}  // namespace mmf
