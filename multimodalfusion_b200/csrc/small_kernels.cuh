// small_kernels.cuh — fp32 SIMT kernels for the latency-bound parts of the path: dense layers
// (SNN blocks, classifier heads, Xlinear reduce/encoder layers), the Kronecker-fusion encoder,
// the discrete-hazard head, and the survival losses (NLL, Cox, pairwise ranking).
//
// One functor-driven 64x64x16 SGEMM serves every small matrix product: operand elements are
// produced by loader functors, so activation derivatives, the hazard/cumprod chain rule and the
// Kronecker outer product are formed on the fly instead of being materialised.
#pragma once
#include "amil_tile.cuh"   // counter-hash dropout bits (KronElem)
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mmf {

// [h > 0] as 1 bit per element: word w of the output covers 32 consecutive bf16 values of H
// (recompute backward only; the stash backward's gate kernel writes the same words itself).
__global__ void relu_mask_kernel(const __nv_bfloat16* __restrict__ H, long long n_words, uint32_t* __restrict__ mask) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint4* src = reinterpret_cast<const uint4*>(H) + w * 4;
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = __ldg(src + j);
    const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // bf16 > 0: sign clear and magnitude non-zero
      const uint32_t lo = ws[k] & 0xFFFFu, hi = ws[k] >> 16;
      bits |= (uint32_t)(lo != 0u && lo < 0x8000u) << (8 * j + 2 * k);
      bits |= (uint32_t)(hi != 0u && hi < 0x8000u) << (8 * j + 2 * k + 1);
    }
  }
  mask[w] = bits;
}


// -------------------------------------------------------------------------------------------
// activations (y = act(pre)); derivative expressed through y so no pre-activation is stored
// -------------------------------------------------------------------------------------------
#define MMF_SELU_ALPHA 1.6732632423543772848170429916717f
#define MMF_SELU_SCALE 1.0507009873554804934193349852946f

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case 1: return fmaxf(x, 0.f);
    case 2: return MMF_SELU_SCALE * (x > 0.f ? x : MMF_SELU_ALPHA * expm1f(x));
    case 3: return 1.f / (1.f + expf(-x));
    case 4: return tanhf(x);
    default: return x;
  }
}
__device__ __forceinline__ float act_grad_from_y(int act, float y) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;
    case 2: return y > 0.f ? MMF_SELU_SCALE : (y + MMF_SELU_SCALE * MMF_SELU_ALPHA);
    case 3: return y * (1.f - y);
    case 4: return 1.f - y * y;
    default: return 1.f;
  }
}

// -------------------------------------------------------------------------------------------
// loader / epilogue functors
// -------------------------------------------------------------------------------------------
struct LoadRowMajor {            // element (i, k) = p[i*ld + k]       (k contiguous)
  const float* p; long long ld;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int i, int k) const { return p[i * ld + k]; }
};
struct LoadColMajor {            // element (i, k) = p[k*ld + i]       (i contiguous)
  const float* p; long long ld;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int i, int k) const { return p[k * ld + i]; }
};
// dpre(b, o) = dy[b,o] * act'(y[b,o])
struct LoadDpre {                // element (b, o), o contiguous -> as A(m=b, k=o): k contiguous
  const float* dy; long long lddy; const float* y; long long ldy; int act;
  static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int b, int o) const {
    return dy[b * lddy + o] * act_grad_from_y(act, y[b * ldy + o]);
  }
};
struct LoadDpreT {               // element (o, b): A(m=o, k=b): m contiguous
  const float* dy; long long lddy; const float* y; long long ldy; int act;
  static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int o, int b) const {
    return dy[b * lddy + o] * act_grad_from_y(act, y[b * ldy + o]);
  }
};
// Kronecker product element kk of o_1 ⊗ o_2 (⊗ o_3 (⊗ o_4)) for sample b (first factor slowest).
// drop != 0: train-mode Dropout(0.25) on the product (XlinearFusion.post_fusion_dropout, models/model_modules.py:170):
// element (b, kk) is kept iff the counter hash of (seed, stream 3, row b, column kk) says so and scaled by 1 / 0.75 —
// the same hash as the attention-MIL dropout (amil_tile.cuh), regenerated in the backward and by
// oracle.dropout_scale_mask(seed, 3, B, E^m); the product is never materialised.
constexpr uint32_t KRON_DROP_STREAM = 3;
struct KronElem {
  const float* o0; const float* o1; const float* o2; const float* o3; int m; int E;
  unsigned long long seed; int drop;   // seed with bit 63 set = device address of the seed word (MMF_SEED_DEVICE, amil_tile.cuh)
  float inv_keep;                      // 1 / (1 - p) of the 16-bit scheme
  // drop: 0 = no dropout; 1 = p = 0.25 with the 2-bit fields shared with the attention-MIL kernels; 2..65535 = drop
  // probability drop / 65536 (any rate, e.g. the cohort heads' 0.7): one 32-bit hash per column PAIR, 16 bits per element,
  // dropped iff the field is below `drop` (oracle.dropout_scale_mask_p)
  __device__ __forceinline__ float keep_scale(int b, int kk) const {
    if (!drop) return 1.f;
    // (these kernels are launched without the PDL attribute: a read-only load of the word an EARLIER kernel wrote is safe,
    // and the compiler may hoist it out of the k loop)
    const unsigned long long sd = (seed >> 63) ? __ldg(reinterpret_cast<const unsigned long long*>(seed & 0x7FFFFFFFFFFFFFFFull)) : seed;
    const uint32_t rs = drop_row_state(sd, KRON_DROP_STREAM, (uint32_t)b);
    if (drop > 1) {
      const uint32_t w = mix32(rs ^ (((uint32_t)kk >> 1) * 0xC2B2AE35U));
      return ((w >> (16u * ((uint32_t)kk & 1u))) & 0xFFFFu) >= (uint32_t)drop ? inv_keep : 0.f;
    }
    return drop_keep(drop_bits16(rs, (uint32_t)kk >> 4), (uint32_t)kk & 15u) ? (1.0f / 0.75f) : 0.f;
  }
  __device__ __forceinline__ float at(int b, int kk) const { return raw(b, kk) * keep_scale(b, kk); }
  __device__ __forceinline__ float raw(int b, int kk) const {
    if (m == 2) { const int i = kk / E, j = kk - i * E; return o0[b * E + i] * o1[b * E + j]; }
    if (m == 3) {
      const int i = kk / (E * E); const int r = kk - i * E * E; const int j = r / E, k = r - j * E;
      return o0[b * E + i] * o1[b * E + j] * o2[b * E + k];
    }
    const int E2 = E * E;
    const int ij = kk / E2, kl = kk - ij * E2;
    const int i = ij / E, j = ij - i * E, k = kl / E, l = kl - k * E;
    return o0[b * E + i] * o1[b * E + j] * o2[b * E + k] * o3[b * E + l];
  }
  __host__ __device__ long long width() const {
    long long w = 1;
    for (int t = 0; t < m; ++t) w *= E;
    return w;
  }
};
struct LoadKronA {               // A(m=b, k=kk)
  KronElem e; static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int b, int kk) const { return e.at(b, kk); }
};
struct LoadKronB {               // B(n=kk, k=b)  (n contiguous)
  KronElem e; static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int kk, int b) const { return e.at(b, kk); }
};
// d(logit) of the hazard head given d_hazards and d_S (cumprod chain rule), K <= 16.
struct HazardDlogit {
  const float* haz; const float* dhaz; const float* dS; int K;
  __device__ __forceinline__ float at(int b, int j) const {
    const float* h = haz + (long long)b * K;
    float g = dhaz ? dhaz[(long long)b * K + j] : 0.f;
    if (dS) {
      // dS_k/dh_j = -prod_{i<=k, i!=j} (1-h_i)  for k >= j
      float pre = 1.f;
      for (int i = 0; i < j; ++i) pre *= (1.f - h[i]);
      float run = pre;  // prod_{i<=k, i != j}
      for (int k = j; k < K; ++k) {
        if (k > j) run *= (1.f - h[k]);
        g -= dS[(long long)b * K + k] * run;
      }
    }
    return g * h[j] * (1.f - h[j]);
  }
};
struct LoadHazA { HazardDlogit d; static constexpr bool kContig = true;
  __device__ __forceinline__ float operator()(int b, int j) const { return d.at(b, j); } };
struct LoadHazAT { HazardDlogit d; static constexpr bool kContig = false;
  __device__ __forceinline__ float operator()(int j, int b) const { return d.at(b, j); } };

struct EpiBiasAct {              // y[m*ld+n] = act(acc + bias[n])
  float* y; long long ld; const float* bias; int act;
  static constexpr bool kLinear = false;       // (non-linear epilogue: no split-K)
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    y[m * ld + n] = act_fwd(act, acc + (bias ? bias[n] : 0.f));
  }
  __device__ __forceinline__ void add(int, int, float) const {}
};
struct EpiStoreAcc {             // c[m*ld+n] = (accumulate ? c : 0) + acc
  float* c; long long ld; int accumulate;
  static constexpr bool kLinear = true;        // split-K slices add their partial sums atomically (c zeroed first if !accumulate)
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    float* p = c + m * ld + n;
    *p = accumulate ? *p + acc : acc;
  }
  __device__ __forceinline__ void add(int m, int n, float acc) const { atomicAdd(c + m * ld + n, acc); }
};

// C(m,n) = sum_k A(m,k) * B(n,k).  256 threads, BK = 16, MT x MT micro-tile per thread: MT = 4 -> 64 x 64 tile (skinny /
// small outputs), MT = 8 -> 128 x 128 tile (ncu, gpurun_out/r2y_sgemm.ncu-rep: the 4 x 4 form issues one LDS.128 per 8 FFMA
// and, with ~9 resident warps per scheduler-quarter at best, stalls on the shared-memory latency — issue slots 63 % active,
// FMA pipe 42 % on the 4913-wide encoder GEMMs; 8 x 8 halves the LDS : FFMA ratio and gives every warp 64 independent
// FFMA per k). The 8 rows / columns of a thread are two groups of 4, 64 apart: the operand reads stay conflict-free
// LDS.128. The operand elements of the NEXT k-step are fetched into registers while the current one is multiplied (the
// functor loads are scalar global loads: with the fetch inside the step a skinny GEMM of 48 k-steps took ~1 us per step —
// latency, not work). gridDim.z > 1: split-K, slice z takes k-steps z, z + gridDim.z, ... and adds its partial tile
// atomically (linear epilogues only; EpiSlice stores per-slice partials instead).
template <class ALoad, class BLoad, class Epi, int MT = 4>
__global__ void __launch_bounds__(256, MT == 8 ? 2 : 3)
sgemm_functor_kernel(int M, int N, int K, ALoad la, BLoad lb, Epi epi) {
  constexpr int TS = 16 * MT;                      // tile side
  __shared__ __align__(16) float As[16][TS + 4];
  __shared__ __align__(16) float Bs[16][TS + 4];
  const int tid = threadIdx.x;
  const int tm = tid / 16, tn = tid % 16;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[MT][MT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) acc[i][j] = 0.f;
  float ra[MT], rb[MT];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < MT; ++e) {
      const int idx = tid + e * 256;
      {
        const int mm = ALoad::kContig ? idx / 16 : idx % TS;
        const int kk = ALoad::kContig ? idx % 16 : idx / TS;
        const int gm = m0 + mm, gk = k0 + kk;
        ra[e] = (gm < M && gk < K) ? la(gm, gk) : 0.f;
      }
      {
        const int nn = BLoad::kContig ? idx / 16 : idx % TS;
        const int kk = BLoad::kContig ? idx % 16 : idx / TS;
        const int gn = n0 + nn, gk = k0 + kk;
        rb[e] = (gn < N && gk < K) ? lb(gn, gk) : 0.f;
      }
    }
  };
  const int kstep = 16 * (int)gridDim.z;
  int k0 = 16 * (int)blockIdx.z;
  if (k0 < K) fetch(k0);
  for (; k0 < K; k0 += kstep) {
#pragma unroll
    for (int e = 0; e < MT; ++e) {
      const int idx = tid + e * 256;
      As[ALoad::kContig ? idx % 16 : idx / TS][ALoad::kContig ? idx / 16 : idx % TS] = ra[e];
      Bs[BLoad::kContig ? idx % 16 : idx / TS][BLoad::kContig ? idx / 16 : idx % TS] = rb[e];
    }
    __syncthreads();
    if (k0 + kstep < K) fetch(k0 + kstep);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[MT], bv[MT];
#pragma unroll
      for (int g = 0; g < MT / 4; ++g) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][g * 64 + tm * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][g * 64 + tn * 4]);
        av[4 * g] = a4.x; av[4 * g + 1] = a4.y; av[4 * g + 2] = a4.z; av[4 * g + 3] = a4.w;
        bv[4 * g] = b4.x; bv[4 * g + 1] = b4.y; bv[4 * g + 2] = b4.z; bv[4 * g + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int gm = m0 + (i / 4) * 64 + tm * 4 + (i % 4), gn = n0 + (j / 4) * 64 + tn * 4 + (j % 4);
      if (gm < M && gn < N) {
        if (Epi::kLinear && split) epi.add(gm, gn, acc[i][j]);
        else epi(gm, gn, acc[i][j]);
      }
    }
}

// Deterministic split-K forward (mmf_dense_fwd_ws): slice z stores its partial tile to ws[z][m][n] (plain stores, no
// atomics), the fix-up kernel adds the slices in a fixed order, then bias + activation.
struct EpiSlice {
  float* ws; long long MN; int N;
  static constexpr bool kLinear = false;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const { ws[blockIdx.z * MN + (long long)m * N + n] = acc; }
  __device__ __forceinline__ void add(int, int, float) const {}
};
__global__ void __launch_bounds__(256) splitk_fixup_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                           const float* __restrict__ bias, int act, float* __restrict__ y,
                                                           long long ld) {
  const long long total = (long long)M * N;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += ws[z * total + i];
    y[m * ld + n] = act_fwd(act, acc + (bias ? bias[n] : 0.f));
  }
}
// split count of the deterministic split-K forward: few output tiles and a long k loop (the radiology reduce_dim,
// [155, 4096] -> 1024: 48 CTAs of 256 serial k-steps took 205 us); 8 warps per CTA cannot fill an SM's FMA pipes
// (the k loop is latency-bound), so the slices aim at ~4 CTAs per SM. 1 = no split.
// 128 x 128 tiles (MT = 8) when both output dimensions fill them (at most a quarter of the padded tile rows / columns idle);
// 64 x 64 otherwise
inline bool sgemm_big_tiles(int M, int N) {
  auto fits = [](int d) { return d >= 128 && (long long)((d + 127) / 128) * 128 * 4 <= 5ll * d; };
  return fits(M) && fits(N) && (long long)M * N >= 128 * 128 * 8;
}
inline int dense_fwd_splits(int B, int in_dim, int out_dim) {
  const bool big = sgemm_big_tiles(B, out_dim);
  const int ts = big ? 128 : 64, want = (big ? 2 : 4) * 148;
  const long long tiles = (long long)((B + ts - 1) / ts) * ((out_dim + ts - 1) / ts);
  const int ksteps = (in_dim + 15) / 16;
  if (tiles >= want / 2 || in_dim < 512) return 1;
  int splits = (int)((want + tiles - 1) / tiles);
  if (splits > ksteps / 4) splits = ksteps / 4;
  if (splits > 32) splits = 32;
  return splits < 1 ? 1 : splits;
}

// Split-K for grids that would leave most of the 148 SMs idle (the fusion heads' weight gradients: 16 x 768 outputs
// over a 512-sample cohort = 12 tiles of 32 k-steps): only for linear epilogues; an output that is not accumulated
// into is cleared first.
template <class ALoad, class BLoad, class Epi>
inline void launch_sgemm(int M, int N, int K, ALoad la, BLoad lb, Epi epi, cudaStream_t st) {
  // 128 x 128 tiles only pay in the plain-operand forward (launch_sgemm_slices: 75 -> 64 us on the two encoder layers of
  // config 3); with the derivative-forming operand functors of the backward GEMMs they measured SLOWER (LoadDpre 100 -> 125 us,
  // LoadDpreT 95 -> 111 us, gpurun_out/r2z / r3a_cfg3_profile.log): those loops are bound by the functor loads, not by LDS / FFMA
  const bool big = false && sgemm_big_tiles(M, N);
  const int ts = big ? 128 : 64;
  dim3 grid((N + ts - 1) / ts, (M + ts - 1) / ts);
  int splits = 1;
  if constexpr (Epi::kLinear) {
    const int tiles = (int)(grid.x * grid.y), ksteps = (K + 15) / 16;
    // (ncu, gpurun_out/r2y_sgemm.ncu-rep: with one 8-warp CTA per SM the k loop is latency-bound — issue slots 19 % active,
    // FMA pipe 11 %: grids under two waves of single CTAs are split towards 4 (64 x 64) / 2 (128 x 128) CTAs per SM)
    const int want = (big ? 2 : 4) * 148;
    if (tiles < want / 2 && ksteps >= 8) {
      splits = (want + tiles - 1) / tiles;
      if (splits > ksteps / 4) splits = ksteps / 4;      // at least 4 k-steps per slice
      if (splits > 32) splits = 32;
      if (splits < 1) splits = 1;
    }
    if (splits > 1 && !epi.accumulate) {
      if (epi.ld == N) cudaMemsetAsync(epi.c, 0, sizeof(float) * (size_t)M * (size_t)N, st);
      else cudaMemset2DAsync(epi.c, sizeof(float) * (size_t)epi.ld, 0, sizeof(float) * (size_t)N, (size_t)M, st);
    }
  }
  grid.z = splits;
  if (big) sgemm_functor_kernel<ALoad, BLoad, Epi, 8><<<grid, 256, 0, st>>>(M, N, K, la, lb, epi);
  else sgemm_functor_kernel<ALoad, BLoad, Epi, 4><<<grid, 256, 0, st>>>(M, N, K, la, lb, epi);
}

// deterministic split-K forward (EpiSlice partials + fix-up): launches the slices, returns the split count used
template <class ALoad, class BLoad>
inline void launch_sgemm_slices(int M, int N, int K, ALoad la, BLoad lb, float* ws, int splits, cudaStream_t st) {
  const bool big = sgemm_big_tiles(M, N);
  const int ts = big ? 128 : 64;
  dim3 grid((N + ts - 1) / ts, (M + ts - 1) / ts, splits);
  EpiSlice epi{ws, (long long)M * N, N};
  if (big) sgemm_functor_kernel<ALoad, BLoad, EpiSlice, 8><<<grid, 256, 0, st>>>(M, N, K, la, lb, epi);
  else sgemm_functor_kernel<ALoad, BLoad, EpiSlice, 4><<<grid, 256, 0, st>>>(M, N, K, la, lb, epi);
}

// out[o] (+)= sum_b elem(b, o): one warp per 32 columns x a slice of the rows, 8 warps per block meet in shared memory
// (one thread per column walking all B rows took 32 us per call at B = 512)
template <class Elem>
__global__ void __launch_bounds__(256) colsum_functor_kernel(int B, int O, Elem e, float* out, int accumulate) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (o < O)
    for (int b = w; b < B; b += 8) s += e(b, o);
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && o < O) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    out[o] = accumulate ? out[o] + t : t;
  }
}

// -------------------------------------------------------------------------------------------
// Kronecker encoder backward, input side: given dkron[B, E^m] contract against the other factors
//   d_o0[b,i] = sum_{j,k} dkron[b,(i,j,k)] o1[b,j] o2[b,k]   etc. (2, 3 or 4 factors).   One block per sample.
// -------------------------------------------------------------------------------------------
__global__ void kron_contract_kernel(const float* __restrict__ dkron, KronElem e, int B,
                                     float* d0, float* d1, float* d2, float* d3) {
  extern __shared__ float sh[];  // 4*E accumulators + 4*E factor values (absent factors read as 1)
  const int E = e.E, m = e.m, b = blockIdx.x;
  float* acc = sh;
  float* f = sh + 4 * E;
  for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    f[i] = e.o0[b * E + i];
    f[E + i] = e.o1[b * E + i];
    f[2 * E + i] = (m >= 3) ? e.o2[b * E + i] : 1.f;
    f[3 * E + i] = (m >= 4) ? e.o3[b * E + i] : 1.f;
  }
  __syncthreads();
  const int KK = (int)e.width();
  const float* row = dkron + (long long)b * KK;
  for (int kk = threadIdx.x; kk < KK; kk += blockDim.x) {
    const float g = row[kk] * e.keep_scale(b, kk);   // d(product) = d(dropped product) x mask
    int i, j, k = 0, l = 0;
    if (m == 4) {
      const int E2 = E * E, ij = kk / E2, kl = kk - ij * E2;
      i = ij / E; j = ij - i * E; k = kl / E; l = kl - k * E;
    } else if (m == 3) { i = kk / (E * E); const int r = kk - i * E * E; j = r / E; k = r - j * E; }
    else { i = kk / E; j = kk - i * E; }
    const float v0 = f[i], v1 = f[E + j], v2 = f[2 * E + k], v3 = f[3 * E + l];
    atomicAdd(&acc[i], g * v1 * v2 * v3);
    atomicAdd(&acc[E + j], g * v0 * v2 * v3);
    if (m >= 3) atomicAdd(&acc[2 * E + k], g * v0 * v1 * v3);
    if (m >= 4) atomicAdd(&acc[3 * E + l], g * v0 * v1 * v2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    d0[b * E + i] = acc[i];
    d1[b * E + i] = acc[E + i];
    if (m >= 3) d2[b * E + i] = acc[2 * E + i];
    if (m >= 4) d3[b * E + i] = acc[3 * E + i];
  }
}

// -------------------------------------------------------------------------------------------
// hazard head forward: one warp per sample.
// -------------------------------------------------------------------------------------------
__global__ void hazard_head_fwd_kernel(const float* __restrict__ M, int B, int Lin,
                                       const float* __restrict__ Wk, const float* __restrict__ bk,
                                       int K, float* __restrict__ hazards, float* __restrict__ S,
                                       long long* __restrict__ Y_hat) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* mrow = M + (long long)warp * Lin;
  float d[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) d[j] = 0.f;
  for (int l = lane; l < Lin; l += 32) {
    const float mv = mrow[l];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < K) d[j] = fmaf(mv, Wk[(long long)j * Lin + l], d[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
  if (lane == 0) {
    float best = -CUDART_INF_F, surv = 1.f;
    int besti = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < K) {
        const float logit = d[j] + bk[j];
        if (logit > best) { best = logit; besti = j; }
        const float h = 1.f / (1.f + expf(-logit));
        surv *= (1.f - h);
        hazards[(long long)warp * K + j] = h;
        S[(long long)warp * K + j] = surv;
      }
    }
    if (Y_hat) Y_hat[warp] = besti;
  }
}

// fp32 [n, 1024] -> bf16 [n, 3072] = [hi | lo | hi]: the split-precision bag format (MMF_PRECISE_FC)
__global__ void split_f32_bf16x3_kernel(const float* __restrict__ x, long long n, long long ldx,
                                        __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 1024) return;
  const long long r = i >> 10, c = i & 1023;
  const float v = x[r * ldx + c];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* o = out + r * 3072 + c;
  o[0] = hi; o[1024] = lo; o[2048] = hi;
}

// bf16 hi / lo split of the classifier for the forward's z = Wk h side MMA (N = 16: rows 0..7 go to the even CTA of
// the pair, rows 8..15 to the odd one): row j < K = bf16(Wk[j]), row 8 + j = bf16(Wk[j] - bf16(Wk[j])), others 0.
__global__ void pack_head_weights_kernel(const float* __restrict__ Wk, int K, int L, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * L) return;
  const int r = i / L, c = i - r * L, j = r & 7;
  float v = 0.f;
  if (j < K) {
    const float w = Wk[(long long)j * L + c];
    const float hi = __bfloat162float(__float2bfloat16_rn(w));
    v = (r < 8) ? hi : w - hi;
  }
  out[i] = __float2bfloat16_rn(v);
}

// -------------------------------------------------------------------------------------------
// Single-bag training step tail, ONE launch: combine the tile partials -> M, (m,l); hazard head;
// nll_surv loss; gradient back to M and to the classifier (dWk, dbk accumulated).
// Replaces 7 launches (combine, head fwd, nll, 3x head bwd) on the batch-1 hot loop
// (utils/core_utils.py:200-247).  One block of 512 threads, n <= 4096 partials, L <= 1024, K <= 16.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
amil_head_step_kernel(const float* __restrict__ parts, int n, int L, const float* __restrict__ Wk,
                      const float* __restrict__ bk, int K, const long long* __restrict__ Yp,
                      const float* __restrict__ cp, float alpha, float eps, float* __restrict__ M,
                      float* __restrict__ ml, float* __restrict__ hazards, float* __restrict__ S,
                      long long* __restrict__ Y_hat, float* __restrict__ loss, float* __restrict__ dM,
                      float* __restrict__ dWk, float* __restrict__ dbk) {
  // The kernel is a chain of latency-bound phases on ONE SM, so every phase issues all of its global
  // loads before the first use: (m_t, l_t) pairs are read once into registers, the combine walks the
  // partial rows with 16 independent float2 loads in flight per thread (2 columns x row groups), and the
  // classifier weights are staged in shared memory while the combine runs. (The first version walked
  // the rows 4 at a time per column: 32 dependent L2 round trips, 15.6 us at n = 128.)
  __shared__ float s_w[4096];
  __shared__ float s_M[1024];
  __shared__ float s_acc[1024];
  __shared__ float s_Wk[4096];
  __shared__ float s_red[16];
  __shared__ float s_logit[16], s_dlogit[16];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long stride = L + 2;
  const bool wk_smem = K * L <= 4096;
  griddep_launch_dependents();
  griddep_wait();
  parts = pdl_fresh(parts);
  // (m_t, l_t) of up to 8 partials per thread, kept in registers for both reductions
  float2 mlv[8];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = tid + 512 * j;
    mlv[j] = (t < n) ? *reinterpret_cast<const float2*>(parts + t * stride) : make_float2(-CUDART_INF_F, 0.f);
    m = fmaxf(m, mlv[j].x);
  }
  if (wk_smem)
    for (int i = tid; i < K * L; i += 512) s_Wk[i] = Wk[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) m = fmaxf(m, s_red[i]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = tid + 512 * j;
    if (t < n) {
      const float w = (mlv[j].x > -CUDART_INF_F) ? __expf(mlv[j].x - m) : 0.f;
      s_w[t] = w;
      l = fmaf(mlv[j].y, w, l);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __syncthreads();   // s_red (max) consumed by everyone; s_w complete
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) l += s_red[i];
  if (tid == 0) { ml[0] = m; ml[1] = l; }
  // combine: CT column threads (2 columns each) x RG row groups
  const int half = L >> 1;
  const int CT = (half <= 512 && 512 % half == 0) ? half : 512;
  const int RG = 512 / CT;
  const int rg = tid / CT;
  for (int cpair = tid % CT; cpair < half; cpair += CT) {
    const float* col = parts + 2 + 2 * cpair;
    float ax = 0.f, ay = 0.f;
    for (int t0 = rg; t0 < n; t0 += 16 * RG) {
      float2 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int t = t0 + u * RG;
        v[u] = (t < n) ? *reinterpret_cast<const float2*>(col + t * stride) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int t = t0 + u * RG;
        const float w = (t < n) ? s_w[t] : 0.f;
        ax = fmaf(v[u].x, w, ax);
        ay = fmaf(v[u].y, w, ay);
      }
    }
    if (RG == 1) {
      const float inv = 1.f / l;
      s_M[2 * cpair] = ax * inv; s_M[2 * cpair + 1] = ay * inv;
      M[2 * cpair] = ax * inv; M[2 * cpair + 1] = ay * inv;
    } else {   // RG * L <= 1024 by construction
      s_acc[rg * L + 2 * cpair] = ax; s_acc[rg * L + 2 * cpair + 1] = ay;
    }
  }
  if (RG > 1) {
    __syncthreads();
    for (int c0 = tid; c0 < L; c0 += 512) {
      float v = 0.f;
      for (int r = 0; r < RG; ++r) v += s_acc[r * L + c0];
      v /= l;
      s_M[c0] = v;
      M[c0] = v;
    }
  }
  __syncthreads();
  // logits: warp j computes class j
  if (wid < K) {
    float d = 0.f;
    const float* wrow = wk_smem ? s_Wk + wid * L : Wk + (long long)wid * L;
    for (int c0 = lane; c0 < L; c0 += 32) d = fmaf(s_M[c0], wrow[c0], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) s_logit[wid] = d + bk[wid];
  }
  __syncthreads();
  if (tid == 0) {
    float h[16], sv[16], dh[16], dS[16];
    float surv = 1.f, best = -CUDART_INF_F;
    int besti = 0;
    for (int j = 0; j < K; ++j) {
      const float lg = s_logit[j];
      if (lg > best) { best = lg; besti = j; }
      h[j] = 1.f / (1.f + expf(-lg));
      surv *= (1.f - h[j]);
      sv[j] = surv;
      hazards[j] = h[j]; S[j] = surv;
      dh[j] = 0.f; dS[j] = 0.f;
    }
    if (Y_hat) *Y_hat = besti;
    const long long y = Yp[0];
    const float cb = cp[0];
    const float sp_y = (y == 0) ? 1.f : sv[y - 1], h_y = h[y], sp_y1 = sv[y];
    const float unc = -(1.f - cb) * (logf(fmaxf(sp_y, eps)) + logf(fmaxf(h_y, eps)));
    const float cen = -cb * logf(fmaxf(sp_y1, eps));
    *loss = (1.f - alpha) * (cen + unc) + alpha * unc;
    if (y > 0 && sp_y >= eps) dS[y - 1] += -(1.f - cb) / sp_y;
    if (sp_y1 >= eps) dS[y] += -(1.f - alpha) * cb / sp_y1;
    if (h_y >= eps) dh[y] += -(1.f - cb) / h_y;
    for (int j = 0; j < K; ++j) {
      float g = dh[j];
      float pre = 1.f;
      for (int i = 0; i < j; ++i) pre *= (1.f - h[i]);
      float run = pre;
      for (int k = j; k < K; ++k) {
        if (k > j) run *= (1.f - h[k]);
        g -= dS[k] * run;
      }
      s_dlogit[j] = g * h[j] * (1.f - h[j]);
    }
  }
  __syncthreads();
  for (int c0 = tid; c0 < L; c0 += 512) {
    float acc = 0.f;
    const float mv = s_M[c0];
    for (int j = 0; j < K; ++j) {
      const float dl = s_dlogit[j];
      acc = fmaf(dl, wk_smem ? s_Wk[j * L + c0] : Wk[(long long)j * L + c0], acc);
      if (dWk) dWk[(long long)j * L + c0] += dl * mv;
    }
    dM[c0] = acc;
  }
  if (dbk && tid < K) dbk[tid] += s_dlogit[tid];
}

// -------------------------------------------------------------------------------------------
// Cohort inference tail: one CTA per bag of a packed multi-bag forward. Combines the bag's tile partials
// (tiles seg_off[b] .. seg_off[b+1]) into M[b], then the discrete-hazard head: hazards, S, Y_hat, risk = -sum S.
// (models/model_attention_mil_path.py:55-61 + utils/core_utils.py:207, for every slide of a cohort in one launch)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
amil_seg_head_kernel(const float* __restrict__ parts, const int* __restrict__ seg_off, int L,
                     const float* __restrict__ Wk, const float* __restrict__ bk, int K, float* __restrict__ M,
                     float* __restrict__ ml, float* __restrict__ hazards, float* __restrict__ S,
                     float* __restrict__ risk, long long* __restrict__ Y_hat) {
  __shared__ float s_w[1024];
  __shared__ float s_M[1024];
  __shared__ float s_red[8];
  __shared__ float s_logit[16];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = seg_off[b], n = seg_off[b + 1] - t0;
  const long long stride = L + 2;
  const float* p0 = parts + (long long)t0 * stride;
  float m = -CUDART_INF_F;
  for (int t = tid; t < n; t += 256) m = fmaxf(m, p0[t * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i]);
  __syncthreads();
  float l = 0.f;
  // tiles are processed in blocks of 1024 softmax weights (a 64k-instance slide has 512 tiles)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tb = 0; tb < n; tb += 1024) {
    const int nb = min(1024, n - tb);
    for (int t = tid; t < nb; t += 256) {
      const float mt = p0[(tb + t) * stride];
      const float w = (mt > -CUDART_INF_F) ? __expf(mt - m) : 0.f;
      s_w[t] = w;
      l = fmaf(p0[(tb + t) * stride + 1], w, l);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = tid + 256 * j;
      if (c0 < L) {
        float a0 = 0.f, a1 = 0.f;
        int t = 0;
        for (; t + 2 <= nb; t += 2) {
          a0 = fmaf(p0[(tb + t) * stride + 2 + c0], s_w[t], a0);
          a1 = fmaf(p0[(tb + t + 1) * stride + 2 + c0], s_w[t + 1], a1);
        }
        if (t < nb) a0 = fmaf(p0[(tb + t) * stride + 2 + c0], s_w[t], a0);
        acc[j] += a0 + a1;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) l += s_red[i];
  const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c0 = tid + 256 * j;
    if (c0 < L) {
      const float v = acc[j] * inv;
      s_M[c0] = v;
      M[(long long)b * L + c0] = v;
    }
  }
  if (tid == 0 && ml) { ml[2 * b] = m; ml[2 * b + 1] = l; }
  __syncthreads();
  for (int j = wid; j < K; j += 8) {
    float d = 0.f;
    for (int c0 = lane; c0 < L; c0 += 32) d = fmaf(s_M[c0], Wk[(long long)j * L + c0], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) s_logit[j] = d + bk[j];
  }
  __syncthreads();
  if (tid == 0) {
    float surv = 1.f, best = -CUDART_INF_F, rsum = 0.f;
    int besti = 0;
    for (int j = 0; j < K; ++j) {
      const float lg = s_logit[j];
      if (lg > best) { best = lg; besti = j; }
      const float h = 1.f / (1.f + expf(-lg));
      surv *= (1.f - h);
      hazards[(long long)b * K + j] = h;
      S[(long long)b * K + j] = surv;
      rsum += surv;
    }
    if (risk) risk[b] = -rsum;
    if (Y_hat) Y_hat[b] = besti;
  }
}

// -------------------------------------------------------------------------------------------
// Cluster version of amil_head_step_kernel: 8 CTAs (one thread-block cluster) split the L pooled
// columns, so the combine is ONE batch of independent loads per thread (n/RG rows each) instead of
// four dependent batches on a single SM, and the K partial logits of every CTA are exchanged through
// distributed shared memory (st.shared::cluster) + one cluster barrier. Everything after the logits
// (hazards, survival, nll_surv, dlogit) is a few hundred scalar cycles and is recomputed by every
// CTA; CTA 0 writes the scalars. Same math, same outputs as the single-CTA kernel (kept for L that
// does not split into 8 x 2 x power-of-two and as the reference implementation).
// -------------------------------------------------------------------------------------------
constexpr int HEAD_CLUSTER = 8;

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

__global__ void __cluster_dims__(HEAD_CLUSTER, 1, 1) __launch_bounds__(512)
amil_head_step_cluster_kernel(const float* __restrict__ parts, int n, int L, const float* __restrict__ Wk,
                              const float* __restrict__ bk, int K, const long long* __restrict__ Yp,
                              const float* __restrict__ cp, float alpha, float eps, float* __restrict__ M,
                              float* __restrict__ ml, float* __restrict__ hazards, float* __restrict__ S,
                              long long* __restrict__ Y_hat, float* __restrict__ loss, float* __restrict__ dM,
                              float* __restrict__ dWk, float* __restrict__ dbk, const int* __restrict__ seg,
                              float loss_scale) {
  __shared__ float s_w[4096];
  __shared__ float s_acc[1024];            // [RG][LC] partial column sums of this CTA's slice
  __shared__ float s_M[128];               // this CTA's slice of M
  __shared__ float s_Wk[16 * 128];         // this CTA's slice of the classifier [K][LC]
  __shared__ float s_red[16];
  __shared__ float s_plog[HEAD_CLUSTER][16];   // partial logits of every CTA (filled through DSMEM)
  __shared__ float s_dlogit[16];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const long long stride = L + 2;
  const int LC = L / HEAD_CLUSTER;         // columns per CTA (32 .. 128)
  const int c_lo = rank * LC;
  const int CP = LC >> 1;                  // column pairs per CTA
  const int RG = 512 / CP;                 // row groups
  Timeline tl = timeline_start(1);
  griddep_launch_dependents();
  griddep_wait();
  parts = pdl_fresh(parts);
  timeline_wait_done(tl);   // PDL: the partials come from the tile kernel launched just before
  // seg != null: a WINDOW of bags (varlen-packed training, blockIdx.y = bag): the bag's tiles are partial rows
  // seg[bag] .. seg[bag + 1], every per-bag input / output is indexed by the bag, the classifier gradients of the bags are
  // added atomically and the loss gradient carries loss_scale (1 / gc of the accumulation window)
  const bool window = seg != nullptr;
  if (window) {
    const int bag = blockIdx.y;
    const int t0 = seg[bag];
    n = seg[bag + 1] - t0;
    parts += (long long)t0 * (L + 2);
    Yp += bag; cp += bag; M += (long long)bag * L; ml += 2 * bag; hazards += (long long)bag * K; S += (long long)bag * K;
    if (Y_hat) Y_hat += bag;
    loss += bag; dM += (long long)bag * L;
  }
  // early scalar loads (independent of everything else)
  long long y = 0; float cb = 0.f;
  if (tid == 0) { y = Yp[0]; cb = cp[0]; }
  float2 mlv[8];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = tid + 512 * j;
    mlv[j] = (t < n) ? *reinterpret_cast<const float2*>(parts + t * stride) : make_float2(-CUDART_INF_F, 0.f);
    m = fmaxf(m, mlv[j].x);
  }
  for (int i = tid; i < K * LC; i += 512) s_Wk[i] = Wk[(long long)(i / LC) * L + c_lo + (i % LC)];
  // the combine's loads do not depend on the softmax weights: issue them now (16 rows per batch)
  const int cpair = tid % CP, rg = tid / CP;
  const float* col = parts + 2 + c_lo + 2 * cpair;
  float2 v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int t = rg + u * RG;
    v[u] = (t < n) ? *reinterpret_cast<const float2*>(col + t * stride) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) m = fmaxf(m, s_red[i]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = tid + 512 * j;
    if (t < n) {
      const float w = (mlv[j].x > -CUDART_INF_F) ? __expf(mlv[j].x - m) : 0.f;
      s_w[t] = w;
      l = fmaf(mlv[j].y, w, l);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __syncthreads();
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) l += s_red[i];
  if (rank == 0 && tid == 0) { ml[0] = m; ml[1] = l; }
  float ax = 0.f, ay = 0.f;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int t = rg + u * RG;
    const float w = (t < n) ? s_w[t] : 0.f;
    ax = fmaf(v[u].x, w, ax);
    ay = fmaf(v[u].y, w, ay);
  }
  for (int t0 = rg + 16 * RG; t0 < n; t0 += 16 * RG) {   // n > 16 RG: further batches
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = t0 + u * RG;
      v[u] = (t < n) ? *reinterpret_cast<const float2*>(col + t * stride) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int t = t0 + u * RG;
      const float w = (t < n) ? s_w[t] : 0.f;
      ax = fmaf(v[u].x, w, ax);
      ay = fmaf(v[u].y, w, ay);
    }
  }
  s_acc[rg * LC + 2 * cpair] = ax;
  s_acc[rg * LC + 2 * cpair + 1] = ay;
  __syncthreads();
  if (tid < LC) {
    float a = 0.f;
    for (int r = 0; r < RG; ++r) a += s_acc[r * LC + tid];
    a /= l;
    s_M[tid] = a;
    M[c_lo + tid] = a;
  }
  __syncthreads();
  // partial logits of this CTA's slice -> every CTA's s_plog[rank][j]
  if (wid < K) {
    float d = 0.f;
    for (int c0 = lane; c0 < LC; c0 += 32) d = fmaf(s_M[c0], s_Wk[wid * LC + c0], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane < HEAD_CLUSTER) st_cluster_f32(mapa_cluster(smem_u32(&s_plog[rank][wid]), lane), d);
  }
  cluster_sync_all();
  if (tid == 0) {
    float h[16], sv[16], dh[16], dS[16];
    float surv = 1.f, best = -CUDART_INF_F;
    int besti = 0;
    for (int j = 0; j < K; ++j) {
      float lg = bk[j];
      for (int r = 0; r < HEAD_CLUSTER; ++r) lg += s_plog[r][j];
      if (lg > best) { best = lg; besti = j; }
      h[j] = 1.f / (1.f + expf(-lg));
      surv *= (1.f - h[j]);
      sv[j] = surv;
      dh[j] = 0.f; dS[j] = 0.f;
    }
    const float sp_y = (y == 0) ? 1.f : sv[y - 1], h_y = h[y], sp_y1 = sv[y];
    if (rank == 0) {
      for (int j = 0; j < K; ++j) { hazards[j] = h[j]; S[j] = sv[j]; }
      if (Y_hat) *Y_hat = besti;
      const float unc = -(1.f - cb) * (logf(fmaxf(sp_y, eps)) + logf(fmaxf(h_y, eps)));
      const float cen = -cb * logf(fmaxf(sp_y1, eps));
      *loss = (1.f - alpha) * (cen + unc) + alpha * unc;
    }
    if (y > 0 && sp_y >= eps) dS[y - 1] += -(1.f - cb) / sp_y;
    if (sp_y1 >= eps) dS[y] += -(1.f - alpha) * cb / sp_y1;
    if (h_y >= eps) dh[y] += -(1.f - cb) / h_y;
    for (int j = 0; j < K; ++j) {
      float g = dh[j];
      float pre = 1.f;
      for (int i = 0; i < j; ++i) pre *= (1.f - h[i]);
      float run = pre;
      for (int k = j; k < K; ++k) {
        if (k > j) run *= (1.f - h[k]);
        g -= dS[k] * run;
      }
      s_dlogit[j] = g * h[j] * (1.f - h[j]) * loss_scale;
    }
  }
  __syncthreads();
  if (tid < LC) {
    float acc = 0.f;
    const float mv = s_M[tid];
    for (int j = 0; j < K; ++j) {
      const float dl = s_dlogit[j];
      acc = fmaf(dl, s_Wk[j * LC + tid], acc);
      if (dWk) {
        float* pw = dWk + (long long)j * L + c_lo + tid;
        if (window) atomicAdd(pw, dl * mv); else *pw += dl * mv;
      }
    }
    dM[c_lo + tid] = acc;
  }
  if (rank == 0 && dbk && tid < K) {
    if (window) atomicAdd(dbk + tid, s_dlogit[tid]); else dbk[tid] += s_dlogit[tid];
  }
  timeline_end(1, tl);
}

// -------------------------------------------------------------------------------------------
// NLL survival loss (utils/loss_utils.py:22-39), forward + gradient. Single block.
// -------------------------------------------------------------------------------------------
__global__ void nll_surv_kernel(const float* __restrict__ haz, const float* __restrict__ S,
                                const long long* __restrict__ Y, const float* __restrict__ c, int B,
                                int K, float alpha, float eps, float* loss, float* d_haz, float* d_S) {
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
    const float sp_y = (y == 0) ? 1.f : s[y - 1];       // S_padded[Y]
    const float h_y = h[y];
    const float sp_y1 = s[y];                             // S_padded[Y+1]
    const float unc = -(1.f - cb) * (logf(fmaxf(sp_y, eps)) + logf(fmaxf(h_y, eps)));
    const float cen = -cb * logf(fmaxf(sp_y1, eps));
    local += (1.f - alpha) * (cen + unc) + alpha * unc;
    if (d_S) {
      if (y > 0 && sp_y >= eps) d_S[(long long)b * K + y - 1] += -(1.f - cb) / sp_y * invB;
      if (sp_y1 >= eps) d_S[(long long)b * K + y] += -(1.f - alpha) * cb / sp_y1 * invB;
    }
    if (d_haz && h_y >= eps) d_haz[(long long)b * K + y] += -(1.f - cb) / h_y * invB;
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

// -------------------------------------------------------------------------------------------
// Cox partial likelihood (utils/loss_utils.py:124-139), B <= 2048, single block of 1024.
// sort by time descending (bitonic on (time, index)), inclusive prefix sums of e^{theta-max}
// extended to tie-group ends, suffix sums of (1-c)/E for the gradient.
// -------------------------------------------------------------------------------------------
#define MMF_COX_MAXB 2048
__global__ void __launch_bounds__(1024)
cox_kernel(const float* __restrict__ theta, const float* __restrict__ times,
           const float* __restrict__ cens, int B, float* loss, float* dtheta) {
  __shared__ float s_t[MMF_COX_MAXB];
  __shared__ int s_i[MMF_COX_MAXB];
  __shared__ float s_a[MMF_COX_MAXB];   // scan buffer A
  __shared__ float s_b[MMF_COX_MAXB];   // scan buffer B
  __shared__ float s_red[33];
  const int tid = threadIdx.x;
  int P = 1;
  while (P < B) P <<= 1;
  // load; padding sorts to the end (time = -inf in a descending sort)
  float tmax = -CUDART_INF_F;
  for (int i = tid; i < P; i += blockDim.x) {
    s_t[i] = i < B ? times[i] : -CUDART_INF_F;
    s_i[i] = i < B ? i : -1;
    if (i < B) tmax = fmaxf(tmax, theta[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = tmax;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < 32; ++i) v = fmaxf(v, s_red[i]);
    s_red[32] = v;
  }
  __syncthreads();
  tmax = s_red[32];
  // bitonic sort, descending by time
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool desc = (i & k) == 0;
          const float a = s_t[i], b = s_t[ixj];
          if (desc ? (a < b) : (a > b)) {
            s_t[i] = b; s_t[ixj] = a;
            const int ti = s_i[i]; s_i[i] = s_i[ixj]; s_i[ixj] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
  // w_k = exp(theta_k - max) in sorted order; inclusive scan (Hillis-Steele, ping-pong)
  for (int i = tid; i < P; i += blockDim.x) s_a[i] = (s_i[i] >= 0) ? expf(theta[s_i[i]] - tmax) : 0.f;
  __syncthreads();
  float* src = s_a; float* dst = s_b;
  for (int off = 1; off < P; off <<= 1) {
    for (int i = tid; i < P; i += blockDim.x) dst[i] = src[i] + (i >= off ? src[i - off] : 0.f);
    __syncthreads();
    float* t = src; src = dst; dst = t;
  }
  // src = inclusive prefix. E_k = prefix[end of tie group]; q_k = (1-c)/E_k ; loss terms
  float local = 0.f;
  for (int i = tid; i < P; i += blockDim.x) {
    float q = 0.f;
    if (i < B) {
      int e = i;
      while (e + 1 < B && s_t[e + 1] == s_t[i]) ++e;
      const float E = src[e];
      const int orig = s_i[i];
      const float ev = 1.f - cens[orig];
      local += ev * (theta[orig] - tmax - logf(E));
      q = ev / E;
    }
    dst[i] = q;
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += s_red[i];
    *loss = -t / (float)B;
  }
  if (dtheta == nullptr) return;
  // suffix sums of q (reverse inclusive scan) : ping-pong between dst (holds q) and src
  float* a = dst; float* b = src;
  for (int off = 1; off < P; off <<= 1) {
    for (int i = tid; i < P; i += blockDim.x) b[i] = a[i] + (i + off < P ? a[i + off] : 0.f);
    __syncthreads();
    float* t = a; a = b; b = t;
  }
  // a = suffix sums. grad_k = -(1/B) [ (1-c_k) - w_k * Q[start of tie group] ]
  for (int i = tid; i < B; i += blockDim.x) {
    int s0 = i;
    while (s0 > 0 && s_t[s0 - 1] == s_t[i]) --s0;
    const int orig = s_i[i];
    const float w = expf(theta[orig] - tmax);
    dtheta[orig] = -((1.f - cens[orig]) - w * a[s0]) / (float)B;
  }
}

// -------------------------------------------------------------------------------------------
// pairwise ranking loss (utils/loss_utils.py:58-101)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float rank_phi(int phi, float r) {
  return phi == 0 ? 1.f / (1.f + expf(-r)) : fmaxf(r, 0.f);
}
__device__ __forceinline__ float rank_dphi(int phi, float r) {
  if (phi == 0) { const float s = 1.f / (1.f + expf(-r)); return s * (1.f - s); }
  return r > 0.f ? 1.f : 0.f;
}
// acc[0] = sum phi, acc[1] = pair count (as double to stay exact); g[i] = unnormalised d(sum phi)/dr_i
__global__ void ranking_pairs_kernel(const float* __restrict__ risks, const float* __restrict__ times,
                                     const float* __restrict__ cens, int B, int phi, double* acc,
                                     float* g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float sum = 0.f, gi = 0.f;
  unsigned int cnt = 0;
  if (i < B) {
    const float ri = risks[i], ti = times[i];
    const bool ei = (1.f - cens[i]) != 0.f;
    for (int j = 0; j < B; ++j) {
      if (j == i) continue;
      const float rj = risks[j], tj = times[j];
      if (ei && ti < tj) {             // (i risky, j safe)
        sum += rank_phi(phi, ri - rj);
        gi += rank_dphi(phi, ri - rj);
        ++cnt;
      }
      if (((1.f - cens[j]) != 0.f) && tj < ti) {  // (j risky, i safe)
        gi -= rank_dphi(phi, rj - ri);
      }
    }
    g[i] = gi;
  }
  // block reduce
  __shared__ float s_sum[32];
  __shared__ unsigned int s_cnt[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, n = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s_sum[w]; n += s_cnt[w]; }
    atomicAdd(&acc[0], t);
    atomicAdd(&acc[1], n);
  }
}
__global__ void ranking_finalize_kernel(const double* acc, int B, int reduction, float* loss,
                                        float* drisks, const float* g, long long* n_pairs) {
  const double n = acc[1];
  const float scale = (n > 0.0) ? (reduction == 0 ? (float)(1.0 / n) : 1.f) : 0.f;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B && drisks) drisks[i] = -g[i] * scale;
  if (i == 0) {
    *loss = (n > 0.0) ? (float)(-(acc[0]) * (reduction == 0 ? 1.0 / n : 1.0)) : 0.f;
    if (n_pairs) *n_pairs = (long long)n;
  }
}

// -------------------------------------------------------------------------------------------
// format helpers
// -------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst,
                                     long long n4, const float* src_tail, __nv_bfloat16* dst_tail,
                                     int tail) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    dst[i] = o;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = __float2bfloat16_rn(src_tail[threadIdx.x]);
}

// packed[c*CHN + r][l] = (r < 128 ? Wa[c*128 + r] : Wb[c*128 + r - 128])[l]   (gated)
__global__ void pack_wab_kernel(const uint4* __restrict__ Wab, uint4* __restrict__ packed, int L,
                                int D, int gated) {
  const int chn = gated ? 256 : 128;
  const int rows = gated ? 2 * D : D;
  const int vec_per_row = L / 8;
  const long long total = (long long)rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int prow = (int)(i / vec_per_row), v = (int)(i % vec_per_row);
    const int c = prow / chn, r = prow % chn;
    const int srow = gated ? (r < 128 ? c * 128 + r : D + c * 128 + (r - 128)) : prow;
    packed[i] = Wab[(long long)srow * vec_per_row + v];
  }
}

// out[c] (+)= sum_r Y[r][c], bf16 input
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ Y, long long rows, int cols,
                                   long long ld, float* out, int accumulate) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (long long r = threadIdx.y; r < rows; r += blockDim.y) acc += __bfloat162float(Y[r * ld + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

}  // namespace mmf
