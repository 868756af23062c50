// snn_mlp.cuh — the self-normalising MLP ("MaxNet" trunk: SNN_Block x n) as ONE forward launch and ONE backward chain
// launch (SURVEY.md §2.2 "K4"). Reference: models/model_modules.py:64-68 (SNN_Block = Linear -> SELU -> AlphaDropout),
// models/model_genomic.py:17-25,53-57 (fc_omic = 2 blocks, 256 / 1024 wide), models/model_mm_attention_mil.py (the omics
// branch of the multimodal model).
//
// The layers of ONE sample form a chain that needs nothing from other samples, so a CTA takes 4 samples through all
// layers: the activations stay in shared memory (two ping-pong buffers), every warp produces output units with
// lane-strided (coalesced) weight reads and a shuffle reduction. AlphaDropout is applied where the next layer reads its
// input: y_drop = a (y m + alpha' (1 - m)) + b with the keep mask m drawn by the caller (one ATen Bernoulli draw for all
// layers) — as separate ops it was ~5 ATen launches per block. Saved for the backward: the pre-dropout SELU outputs.
// Backward chain (same CTA shape): dpre_l = g_l * a m * selu'(y_l), g_{l-1} = dpre_l W_l — dpre_l goes to global memory
// and the weight gradients dW_l = dpre_l^T in_l run on the functor SGEMM with an operand loader that re-applies the
// dropout to the saved activations (in_l is never materialised).
#pragma once
#include <stdint.h>

namespace mmf {

constexpr int SNN_MAX_LAYERS = 4;
constexpr int SNN_MAX_WIDTH = 1024;     // hidden widths (the input width is free)
constexpr int SNN_RB = 4;               // samples per CTA
// torch.nn.functional.alpha_dropout: alpha' = -selu_scale * selu_alpha
#define MMF_ALPHA_PRIME (-1.7580993408473766f)

struct SnnLayers {
  int n; int B; int d[SNN_MAX_LAYERS + 1];          // d[0] = input width, d[l + 1] = width of layer l
  const float* W[SNN_MAX_LAYERS]; const float* b[SNN_MAX_LAYERS];
  const float* keep[SNN_MAX_LAYERS];                // [B, d[l + 1]] keep mask (0 / 1) of layer l's AlphaDropout, or null
  float da[SNN_MAX_LAYERS], db[SNN_MAX_LAYERS];     // the dropout's affine: a = ((1 - p)(1 + p alpha'^2))^-1/2, b = -a alpha' p
  float* y[SNN_MAX_LAYERS];                         // [B, d[l + 1]] pre-dropout SELU outputs (saved)
};

__device__ __forceinline__ float snn_drop(float y, float m, float a, float b) {
  return fmaf(a, fmaf(y, m, MMF_ALPHA_PRIME * (1.f - m)), b);
}

__global__ void __launch_bounds__(256) snn_mlp_fwd_kernel(const SnnLayers P, const float* __restrict__ x, float* __restrict__ out) {
  __shared__ float buf[2][SNN_RB][SNN_MAX_WIDTH];
  const int b0 = blockIdx.x * SNN_RB;
  const int rows = min(SNN_RB, P.B - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* in = x + (long long)b0 * P.d[0];
  long long in_ld = P.d[0];
  for (int l = 0; l < P.n; ++l) {
    const int d_in = P.d[l], d_out = P.d[l + 1];
    const float* W = P.W[l];
    float (*nxt)[SNN_MAX_WIDTH] = buf[l & 1];
    for (int o = warp; o < d_out; o += 8) {
      float acc[SNN_RB];
#pragma unroll
      for (int r = 0; r < SNN_RB; ++r) acc[r] = 0.f;
      const float* w = W + (long long)o * d_in;
      for (int k = lane; k < d_in; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int r = 0; r < SNN_RB; ++r)
          if (r < rows) acc[r] = fmaf(wv, in[r * in_ld + k], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < SNN_RB; ++r)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
      if (lane < rows) {
        float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        v += __ldg(P.b[l] + o);
        v = MMF_SELU_SCALE * (v > 0.f ? v : MMF_SELU_ALPHA * expm1f(v));
        const long long gi = (long long)(b0 + lane) * d_out + o;
        P.y[l][gi] = v;
        if (P.keep[l]) v = snn_drop(v, __ldg(P.keep[l] + gi), P.da[l], P.db[l]);
        nxt[lane][o] = v;
        if (l == P.n - 1) out[gi] = v;
      }
    }
    __syncthreads();
    in = &nxt[0][0]; in_ld = SNN_MAX_WIDTH;
  }
}

// backward chain: g = gradient w.r.t. the network output (post-dropout). Per layer (last to first):
//   dpre_l[r][o] = g[r][o] * (keep ? a m : 1) * selu'(y_l[r][o])     -> global (weight gradients, bias column sums)
//   g[r][k]      = sum_o dpre_l[r][o] W_l[o][k]                       (for l = 0 only when dx is wanted)
struct SnnBwd { float* dpre[SNN_MAX_LAYERS]; float* dx; };

__global__ void __launch_bounds__(256) snn_mlp_bwd_chain_kernel(const SnnLayers P, const float* __restrict__ dout, const SnnBwd G) {
  __shared__ float dps[SNN_RB][SNN_MAX_WIDTH];
  __shared__ float gs[SNN_RB][SNN_MAX_WIDTH];
  const int b0 = blockIdx.x * SNN_RB;
  const int rows = min(SNN_RB, P.B - b0);
  for (int l = P.n - 1; l >= 0; --l) {
    const int d_in = P.d[l], d_out = P.d[l + 1];
    for (int e = threadIdx.x; e < rows * d_out; e += 256) {
      const int r = e / d_out, o = e - r * d_out;
      const long long gi = (long long)(b0 + r) * d_out + o;
      const float g = (l == P.n - 1) ? dout[gi] : gs[r][o];
      const float yv = P.y[l][gi];
      const float fac = P.keep[l] ? P.da[l] * __ldg(P.keep[l] + gi) : 1.f;
      const float dp = g * fac * (yv > 0.f ? MMF_SELU_SCALE : (yv + MMF_SELU_SCALE * MMF_SELU_ALPHA));
      dps[r][o] = dp;
      G.dpre[l][gi] = dp;
    }
    __syncthreads();
    if (l > 0 || G.dx != nullptr) {
      const float* W = P.W[l];
      for (int k = threadIdx.x; k < d_in; k += 256) {
        float acc[SNN_RB];
#pragma unroll
        for (int r = 0; r < SNN_RB; ++r) acc[r] = 0.f;
        for (int o = 0; o < d_out; ++o) {
          const float wv = __ldg(W + (long long)o * d_in + k);
#pragma unroll
          for (int r = 0; r < SNN_RB; ++r) acc[r] = fmaf(dps[r][o], wv, acc[r]);
        }
        if (l > 0) {
#pragma unroll
          for (int r = 0; r < SNN_RB; ++r) gs[r][k] = acc[r];
        } else {
#pragma unroll
          for (int r = 0; r < SNN_RB; ++r)
            if (r < rows) G.dx[(long long)(b0 + r) * d_in + k] = acc[r];
        }
      }
    }
    __syncthreads();
  }
}

// operand loaders of the weight-gradient GEMMs  dW_l[o, k] = sum_b dpre_l[b, o] in_l[b, k]
struct LoadPlainT {              // A(m = o, k = b) = p[b * ld + o]   (m contiguous)
  const float* p; long long ld;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int o, int b) const { return p[b * ld + o]; }
};
struct LoadAlphaDropCol {        // B(n = k, k = b) = drop(y[b * ld + k])   (n contiguous): the layer input, re-formed on the fly
  const float* y; const float* keep; long long ld; float a, b_;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int k, int b) const {
    const float v = y[b * ld + k];
    return keep ? snn_drop(v, keep[b * ld + k], a, b_) : v;
  }
};
struct ElemPlain {               // column sums of dpre (bias gradients)
  const float* p; long long ld;
  __device__ __forceinline__ float operator()(int b, int o) const { return p[b * ld + o]; }
};

}  // namespace mmf
