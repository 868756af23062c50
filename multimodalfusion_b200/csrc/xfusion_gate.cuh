// xfusion_gate.cuh — the per-modality gated reduction of XlinearFusion as ONE forward launch and (up to) three backward
// launches for ALL modalities (SURVEY.md §2.2 "K5"). Reference: models/model_modules.py:156-166 —
//     h_i = relu(Wh_i v_i + bh_i)                      Linear(dim, S) + ReLU
//     z_i = sigmoid(Wz_i cat(v_1..v_m) + bz_i)         Linear(dim m, S)
//     o_i = dropout(relu(Wo_i (z_i * h_i) + bo_i))     Linear(S, S) + ReLU + Dropout
//     o_i <- cat(o_i, 1)                               the constant column of the Kronecker product
// with S = dim / scale_dim = 16 in every configuration of the reference (256 / 16 and 1024 / 64). As separate layers this
// was 3 m functor-SGEMM launches + 4 m ATen launches forward (gating product, dropout, ones, cat) and ~12 m launches
// backward (two skinny weight-gradient GEMMs, a column sum and an input-gradient GEMM per layer) — config 3 of BASELINE.json
// spent 2/3 of its 134 launches here. The concatenation is never formed: the z GEMM walks the m input tensors as k-segments
// and the h GEMM of modality i rides on segment i of the same pass (same x values, second weight row).
//
// Everything is fp32 on the CUDA cores (16 output units per layer: no tensor-core shape), deterministic (no atomics):
// the forward is one CTA per (8 samples, modality); the backward reduces the batch in up to 16 row slices whose partial
// sums a finalize kernel adds in a fixed order.
#pragma once
#include <stdint.h>

namespace mmf {

constexpr int XF_S = 16;           // gate width (dim / scale_dim)
constexpr int XF_MAX_MOD = 4;
constexpr int XF_RB = 8;           // samples per forward CTA
constexpr int XF_KC = 256;         // k-chunk staged in shared memory (dim % XF_KC == 0)

struct XfMod {                     // one modality's parameters (all fp32, row-major [out, in])
  const float* v;                  // [B, dim] embedding
  const float* Wh; const float* bh;   // [S, dim], [S]
  const float* Wz; const float* bz;   // [S, dim m], [S]
  const float* Wo; const float* bo;   // [S, S], [S]
};
struct XfMods { XfMod mod[XF_MAX_MOD]; int m; int B; int dim; };
struct XfGrads {
  float* dWh[XF_MAX_MOD]; float* dbh[XF_MAX_MOD]; float* dWz[XF_MAX_MOD]; float* dbz[XF_MAX_MOD];
  float* dWo[XF_MAX_MOD]; float* dbo[XF_MAX_MOD];
  float* dv[XF_MAX_MOD];           // [B, dim] or null (embeddings that need no gradient)
};

// forward: h, z [m, B, S]; o [m, B, S + 1] (last column = 1); mask [m, B, S] = 0 or 1 / (1 - p), or null
__global__ void __launch_bounds__(256) xfusion_gate_fwd_kernel(const XfMods P, const float* __restrict__ mask,
                                                               float* __restrict__ h_out, float* __restrict__ z_out,
                                                               float* __restrict__ o_out) {
  __shared__ __align__(16) float xs[XF_RB][XF_KC];
  __shared__ float pre_z[XF_RB][XF_S], pre_h[XF_RB][XF_S], gs[XF_RB][XF_S];
  const int i = blockIdx.y, b0 = blockIdx.x * XF_RB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const XfMod M = P.mod[i];
  const int dim = P.dim, KZ = P.dim * P.m, B = P.B;
  const int s0 = 2 * warp, s1 = 2 * warp + 1;
  float az0[XF_RB], az1[XF_RB], ah0[XF_RB], ah1[XF_RB];
#pragma unroll
  for (int r = 0; r < XF_RB; ++r) az0[r] = az1[r] = ah0[r] = ah1[r] = 0.f;
  for (int kc = 0; kc < KZ; kc += XF_KC) {
    const int j = kc / dim, kj = kc - j * dim;         // the chunk lies inside segment j (dim % XF_KC == 0)
    const float* vj = P.mod[j].v;
    __syncthreads();
    for (int e = threadIdx.x; e < XF_RB * (XF_KC / 4); e += 256) {
      const int r = e / (XF_KC / 4), c4 = e - r * (XF_KC / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < B) val = *reinterpret_cast<const float4*>(vj + (long long)(b0 + r) * dim + kj + 4 * c4);
      *reinterpret_cast<float4*>(&xs[r][4 * c4]) = val;
    }
    __syncthreads();
    const bool own = j == i;                            // segment i also feeds h_i
    const float* wz0 = M.Wz + (long long)s0 * KZ + kc;
    const float* wz1 = M.Wz + (long long)s1 * KZ + kc;
    const float* wh0 = M.Wh + (long long)s0 * dim + kj;
    const float* wh1 = M.Wh + (long long)s1 * dim + kj;
#pragma unroll 2
    for (int kk = lane; kk < XF_KC; kk += 32) {
      const float a0 = __ldg(wz0 + kk), a1 = __ldg(wz1 + kk);
      const float c0 = own ? __ldg(wh0 + kk) : 0.f, c1 = own ? __ldg(wh1 + kk) : 0.f;
#pragma unroll
      for (int r = 0; r < XF_RB; ++r) {
        const float x = xs[r][kk];
        az0[r] = fmaf(a0, x, az0[r]); az1[r] = fmaf(a1, x, az1[r]);
        ah0[r] = fmaf(c0, x, ah0[r]); ah1[r] = fmaf(c1, x, ah1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < XF_RB; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      az0[r] += __shfl_xor_sync(0xffffffffu, az0[r], o); az1[r] += __shfl_xor_sync(0xffffffffu, az1[r], o);
      ah0[r] += __shfl_xor_sync(0xffffffffu, ah0[r], o); ah1[r] += __shfl_xor_sync(0xffffffffu, ah1[r], o);
    }
    if (lane == 0) { pre_z[r][s0] = az0[r]; pre_z[r][s1] = az1[r]; pre_h[r][s0] = ah0[r]; pre_h[r][s1] = ah1[r]; }
  }
  __syncthreads();
  const int r = threadIdx.x >> 4, s = threadIdx.x & 15;      // 128 threads: one (sample, unit) each
  const bool act = threadIdx.x < XF_RB * XF_S && b0 + r < B;
  if (threadIdx.x < XF_RB * XF_S) {
    const float hv = fmaxf(pre_h[r][s] + __ldg(M.bh + s), 0.f);
    const float zv = 1.f / (1.f + __expf(-(pre_z[r][s] + __ldg(M.bz + s))));
    gs[r][s] = hv * zv;
    if (act) {
      const long long idx = ((long long)i * B + b0 + r) * XF_S + s;
      h_out[idx] = hv; z_out[idx] = zv;
    }
  }
  __syncthreads();
  if (act) {
    float acc = __ldg(M.bo + s);
#pragma unroll
    for (int t = 0; t < XF_S; ++t) acc = fmaf(__ldg(M.Wo + s * XF_S + t), gs[r][t], acc);
    float ov = fmaxf(acc, 0.f);
    if (mask) ov *= __ldg(mask + ((long long)i * B + b0 + r) * XF_S + s);
    float* orow = o_out + ((long long)i * B + b0 + r) * (XF_S + 1);
    orow[s] = ov;
    if (s == 0) orow[XF_S] = 1.f;
  }
}

// The backward splits the batch into Z row slices (xf_slices): every CTA reduces its slice and stores PARTIAL sums (plain
// stores), a finalize kernel adds the Z partials in a fixed order — deterministic, no atomics, and the two batch loops
// (latency-bound: one global-load round trip per 32 / 64 rows) run Z-wide instead of serially.
constexpr int XF_SMALL_OUT = XF_S * XF_S + 3 * XF_S;     // per (slice, modality): dWo [16 x 16] | dbo | dbh | dbz
__host__ __device__ inline int xf_slices(int B) { const int z = (B + 63) / 64; return z < 1 ? 1 : (z > 16 ? 16 : z); }
__host__ __device__ inline int xf_rows_per_slice(int B) { const int z = xf_slices(B); return ((B + z - 1) / z + 63) / 64 * 64; }

// backward A: grid (Z, m), one CTA per (row slice, modality), 64 samples per iteration:
//   dpre_o = d_o * mask * [o > 0];  dWo = dpre_o^T (z * h), dbo;  dg = dpre_o Wo;  dh = dg z [h > 0];  dz = dg h z (1 - z)
//   dhz[m, B, 2 S] = [dz | dh] for the weight / input gradient kernels;  dbh = sum dh, dbz = sum dz
__global__ void __launch_bounds__(1024) xfusion_gate_bwd_small_kernel(const XfMods P, const float* __restrict__ mask,
                                                                      const float* __restrict__ h, const float* __restrict__ z,
                                                                      const float* __restrict__ o, const float* __restrict__ d_o,
                                                                      float* __restrict__ dhz, float* __restrict__ part_small) {
  __shared__ float dps[64][XF_S + 1], gsm[64][XF_S + 1], dhs[64][XF_S + 1], dzs[64][XF_S + 1], wo[XF_S][XF_S];
  const int i = blockIdx.y, B = P.B;
  const int rows = xf_rows_per_slice(B), b_begin = blockIdx.x * rows, b_end = min(B, b_begin + rows);
  const int r = threadIdx.x >> 4, s = threadIdx.x & 15;
  if (threadIdx.x < XF_S * XF_S) wo[threadIdx.x >> 4][threadIdx.x & 15] = __ldg(P.mod[i].Wo + threadIdx.x);
  // threads 0..255 own dWo[s2][t2]; 256..271 dbo; 272..287 dbh; 288..303 dbz
  float red = 0.f;
  for (int b0 = b_begin; b0 < b_end; b0 += 64) {
    const int b = b0 + r;
    const bool ok = b < b_end;
    const long long idx = ((long long)i * B + b) * XF_S + s;
    float hv = 0.f, zv = 0.f, dp = 0.f;
    if (ok) {
      hv = h[idx]; zv = z[idx];
      const long long oi = ((long long)i * B + b) * (XF_S + 1) + s;
      dp = (o[oi] > 0.f) ? d_o[oi] * (mask ? mask[idx] : 1.f) : 0.f;
    }
    __syncthreads();                 // (previous iteration's readers are done; wo is staged)
    dps[r][s] = dp; gsm[r][s] = hv * zv;
    __syncthreads();
    float dg = 0.f;
#pragma unroll
    for (int t = 0; t < XF_S; ++t) dg = fmaf(dps[r][t], wo[t][s], dg);
    const float dh = (hv > 0.f) ? dg * zv : 0.f;
    const float dz = dg * hv * zv * (1.f - zv);
    dhs[r][s] = dh; dzs[r][s] = dz;
    if (ok) {
      float* drow = dhz + ((long long)i * B + b) * (2 * XF_S);
      drow[s] = dz; drow[XF_S + s] = dh;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      const int s2 = threadIdx.x >> 4, t2 = threadIdx.x & 15;
#pragma unroll 8
      for (int rr = 0; rr < 64; ++rr) red = fmaf(dps[rr][s2], gsm[rr][t2], red);
    } else if (threadIdx.x < XF_SMALL_OUT) {
      const int which = (threadIdx.x - 256) >> 4, s2 = threadIdx.x & 15;
      const float (*src)[XF_S + 1] = which == 0 ? dps : which == 1 ? dhs : dzs;
#pragma unroll 8
      for (int rr = 0; rr < 64; ++rr) red += src[rr][s2];
    }
  }
  if (threadIdx.x < XF_SMALL_OUT)
    part_small[((long long)blockIdx.x * P.m + i) * XF_SMALL_OUT + threadIdx.x] = red;
}

// backward B: weight gradients. One CTA per (64 columns of the concatenated input, modality i, row slice) -> partials
//   part_w[z][i][s][0 .. KZ) : sum_b dz_i[b][s] x[b][k];   on segment i also [KZ + (k - i dim)] : sum_b dh_i[b][s] v_i[b][k - i dim]
__global__ void __launch_bounds__(256) xfusion_gate_bwd_wgrad_kernel(const XfMods P, const float* __restrict__ dhz,
                                                                     float* __restrict__ part_w) {
  __shared__ float xs[32][64];
  __shared__ float ds[32][2 * XF_S];
  const int i = blockIdx.y, k0 = blockIdx.x * 64;
  const int dim = P.dim, KZ = P.dim * P.m, B = P.B;
  const int rows = xf_rows_per_slice(B), b_begin = blockIdx.z * rows, b_end = min(B, b_begin + rows);
  const int j = k0 / dim, kj = k0 - j * dim;
  const bool own = j == i;
  const float* vj = P.mod[j].v;
  const int kcol = threadIdx.x & 63, sg = threadIdx.x >> 6;     // units 4 sg .. 4 sg + 3
  float az[4] = {0.f, 0.f, 0.f, 0.f}, ah[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b0 = b_begin; b0 < b_end; b0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int rr = e >> 6, cc = e & 63;
      xs[rr][cc] = (b0 + rr < b_end) ? vj[(long long)(b0 + rr) * dim + kj + cc] : 0.f;
    }
    for (int e = threadIdx.x; e < 32 * 2 * XF_S; e += 256) {
      const int rr = e >> 5, cc = e & 31;
      ds[rr][cc] = (b0 + rr < b_end) ? dhz[((long long)i * B + b0 + rr) * (2 * XF_S) + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float x = xs[rr][kcol];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        az[q] = fmaf(ds[rr][4 * sg + q], x, az[q]);
        ah[q] = fmaf(ds[rr][XF_S + 4 * sg + q], x, ah[q]);
      }
    }
  }
  float* base = part_w + ((long long)blockIdx.z * P.m + i) * XF_S * (KZ + dim);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float* row = base + (long long)(4 * sg + q) * (KZ + dim);
    row[k0 + kcol] = az[q];
    if (own) row[KZ + kj + kcol] = ah[q];
  }
}

// backward finalize: the Z partials added in slice order -> dWz, dWh, dWo, dbo, dbh, dbz (written or accumulated)
__global__ void __launch_bounds__(256) xfusion_gate_bwd_finalize_kernel(const XfMods P, const float* __restrict__ part_small,
                                                                        const float* __restrict__ part_w, const XfGrads G,
                                                                        int accumulate) {
  const int dim = P.dim, KZ = P.dim * P.m, m = P.m, Z = xf_slices(P.B);
  const long long per_w = (long long)XF_S * (KZ + dim), per_mod = per_w + XF_SMALL_OUT;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < per_mod * m; e += (long long)gridDim.x * 256) {
    const int i = (int)(e / per_mod);
    const long long q = e - (long long)i * per_mod;
    float acc = 0.f;
    float* dst;
    if (q < per_w) {
      for (int zz = 0; zz < Z; ++zz) acc += part_w[((long long)zz * m + i) * per_w + q];
      const int s = (int)(q / (KZ + dim)), c = (int)(q - (long long)s * (KZ + dim));
      dst = c < KZ ? G.dWz[i] + (long long)s * KZ + c : G.dWh[i] + (long long)s * dim + (c - KZ);
    } else {
      const int t = (int)(q - per_w);
      for (int zz = 0; zz < Z; ++zz) acc += part_small[((long long)zz * m + i) * XF_SMALL_OUT + t];
      dst = t < 256 ? G.dWo[i] + t : t < 272 ? G.dbo[i] + (t - 256) : t < 288 ? G.dbh[i] + (t - 272) : G.dbz[i] + (t - 288);
    }
    *dst = accumulate ? *dst + acc : acc;
  }
}

// backward C (only when an embedding needs its gradient — the multimodal model, not the cohort heads):
//   dv_j[b][k] = sum_i sum_s dz_i[b][s] Wz_i[s][j dim + k]  +  sum_s dh_j[b][s] Wh_j[s][k]
__global__ void __launch_bounds__(256) xfusion_gate_bwd_dv_kernel(const XfMods P, const float* __restrict__ dhz, const XfGrads G) {
  __shared__ float ds[XF_MAX_MOD][XF_RB][2 * XF_S];
  const int j = blockIdx.y, b0 = blockIdx.x * XF_RB;
  const int dim = P.dim, KZ = P.dim * P.m, B = P.B, m = P.m;
  float* out = G.dv[j];
  if (out == nullptr) return;
  for (int e = threadIdx.x; e < m * XF_RB * 2 * XF_S; e += 256) {
    const int i = e / (XF_RB * 2 * XF_S), rem = e - i * (XF_RB * 2 * XF_S), rr = rem >> 5, cc = rem & 31;
    ds[i][rr][cc] = (b0 + rr < B) ? dhz[((long long)i * B + b0 + rr) * (2 * XF_S) + cc] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < dim; k += 256) {
    float acc[XF_RB];
#pragma unroll
    for (int rr = 0; rr < XF_RB; ++rr) acc[rr] = 0.f;
    for (int i = 0; i < m; ++i) {
      const float* wz = P.mod[i].Wz + (long long)j * dim + k;
#pragma unroll 4
      for (int s = 0; s < XF_S; ++s) {
        const float w = __ldg(wz + (long long)s * KZ);
#pragma unroll
        for (int rr = 0; rr < XF_RB; ++rr) acc[rr] = fmaf(ds[i][rr][s], w, acc[rr]);
      }
    }
    const float* wh = P.mod[j].Wh + k;
#pragma unroll 4
    for (int s = 0; s < XF_S; ++s) {
      const float w = __ldg(wh + (long long)s * dim);
#pragma unroll
      for (int rr = 0; rr < XF_RB; ++rr) acc[rr] = fmaf(ds[j][rr][XF_S + s], w, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < XF_RB; ++rr)
      if (b0 + rr < B) out[(long long)(b0 + rr) * dim + k] = acc[rr];
  }
}

}  // namespace mmf
