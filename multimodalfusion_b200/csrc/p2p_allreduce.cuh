// p2p_allreduce.cuh — SUM all-reduce of a small fp32 buffer over NVLink peer memory (sm_100a).
//
// The cohort-data-parallel training step ends with one all-reduce of the flat gradient buffer
// (3.7 MB big preset). At that size a collective is latency-bound, and an NCCL kernel launched next to
// the 1-CTA-per-SM tensor-core kernels of the step has to fight them for SMs. This is the exchange
// written as ONE kernel on the step's own stream (it can be captured in the step's CUDA graph):
//
//   every rank's buffer lives in symmetric memory (same size on every rank, all peers mapped);
//   rank r owns slice r of the buffer. CTA b of rank r
//     1. tells CTA b of every peer that rank r's data is ready and waits for theirs (st.release.sys /
//        ld.acquire.sys on flag words in peer memory; flags carry a per-CTA launch epoch, so nothing
//        is ever reset and replaying a captured graph works),
//     2. reduce-scatter + all-gather in one pass: loads its part of slice r from all `world` buffers
//        (peer loads, 16 B per lane), sums in a fixed rank order (bitwise identical result on every
//        rank), and stores the sum into slice r of EVERY rank's buffer (peer stores),
//     3. fences, signals "done" to every peer and waits for theirs: when the kernel retires, all
//        slices of the local buffer hold the global sum.
//   Remote traffic per rank: (world-1)/world of the buffer in each direction (two-shot), vs world-1
//   buffers for a one-shot exchange. When the buffer also has an NVLS multicast mapping, step 2 is
//   multimem.ld_reduce + multimem.st: the NVSwitch sums the copies and broadcasts the result.
#pragma once
#include <stdint.h>

namespace mmf {

constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_MAX_CTAS = 64;
constexpr int P2P_THREADS = 512;
// flag layout per rank (uint32 words): ready[P2P_MAX_WORLD][P2P_MAX_CTAS], done[...][...], epoch[P2P_MAX_CTAS]
constexpr int P2P_FLAG_WORDS = 2 * P2P_MAX_WORLD * P2P_MAX_CTAS + P2P_MAX_CTAS;

struct P2pArgs {
  float* mc;                       // NVLS multicast mapping of the buffer (all ranks), or null
  float* buf[P2P_MAX_WORLD];       // buf[p] = rank p's buffer as mapped in THIS process
  uint32_t* flags[P2P_MAX_WORLD];  // flags[p] = rank p's flag block as mapped in this process
  long long n;                     // elements (multiple of 4)
  int world, rank;
  unsigned long long* stamps;      // debug (optional): [0] = record count, records of 4 x %globaltimer ns from [8]:
                                   // kernel start, ready handshake done, data phase done, done handshake done (CTA 0)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_sum_kernel(const P2pArgs a) {
  griddep_launch_dependents();
  griddep_wait();   // the local gradients come from the kernel launched just before
  const int b = blockIdx.x, tid = threadIdx.x;
  unsigned long long* rec = nullptr;
  if (a.stamps && b == 0 && tid == 0) {
    const unsigned long long k = atomicAdd(a.stamps, 1ull);
    if (k < 4000) { rec = a.stamps + 8 + 4 * k; rec[0] = globaltimer_ns(); }
  }
  uint32_t* my = a.flags[a.rank];
  __shared__ uint32_t s_epoch;
  if (tid == 0) s_epoch = my[2 * P2P_MAX_WORLD * P2P_MAX_CTAS + b] + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // 1. ready handshake (thread p talks to rank p)
  if (tid < a.world) {
    st_release_sys(a.flags[tid] + a.rank * P2P_MAX_CTAS + b, epoch);
    const uint32_t* w = my + tid * P2P_MAX_CTAS + b;
    uint32_t spins = 0;
    while ((int)(ld_acquire_sys(w) - epoch) < 0) {
      if (++spins > (1u << 28)) { printf("mmf p2p: ready wait timed out (rank %d cta %d peer %d)\n", a.rank, b, tid); __trap(); }
    }
  }
  __syncthreads();
  if (rec) rec[1] = globaltimer_ns();
  // 2. reduce my slice from every rank, push the sum to every rank
  const long long n4 = a.n >> 2;
  const long long per_rank = (n4 + a.world - 1) / a.world;
  const long long lo = a.rank * per_rank;
  const long long hi = lo + per_rank < n4 ? lo + per_rank : n4;
  if (a.mc != nullptr) {
    // NVSwitch does the arithmetic (NVLS): one multimem.ld_reduce returns the sum over every rank's copy,
    // one multimem.st writes it back to every rank — each rank moves 1/world of the buffer, once.
    // (8 independent ld_reduce in flight per thread: a switch round trip is ~2.7 us)
    constexpr int MU = 8;
    const long long step = (long long)gridDim.x * P2P_THREADS;
    for (long long i0 = lo + (long long)b * P2P_THREADS + tid; i0 < hi; i0 += step * MU) {
      float4 v[MU];
#pragma unroll
      for (int u = 0; u < MU; ++u)
        if (i0 + u * step < hi)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(a.mc + 4 * (i0 + u * step)) : "memory");
#pragma unroll
      for (int u = 0; u < MU; ++u)
        if (i0 + u * step < hi)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                       ::"l"(a.mc + 4 * (i0 + u * step)), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
    }
  } else {
  // NVLink loads are latency-bound (~2.5 us round trip): every thread keeps UNR x world independent 16-byte
  // loads in flight (the first version had one per peer and ran the 3.7 MB exchange at 127 GB/s)
  constexpr int UNR = 4;
  const long long step = (long long)gridDim.x * P2P_THREADS;
  for (long long i0 = lo + (long long)b * P2P_THREADS + tid; i0 < hi; i0 += step * UNR) {
    float4 acc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.world <= 4) {
      float4 v[4][UNR];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (p < a.world && i0 + u * step < hi) v[p][u] = reinterpret_cast<const float4*>(a.buf[p])[i0 + u * step];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (p < a.world && i0 + u * step < hi) {
            acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w;
          }
    } else {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {   // two batches of 4 peers: 16 loads in flight, fixed summation order
        float4 v[4][UNR];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            if (4 * ph + p < a.world && i0 + u * step < hi)
              v[p][u] = reinterpret_cast<const float4*>(a.buf[4 * ph + p])[i0 + u * step];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            if (4 * ph + p < a.world && i0 + u * step < hi) {
              acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w;
            }
      }
    }
#pragma unroll
    for (int p = 0; p < P2P_MAX_WORLD; ++p)
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (p < a.world && i0 + u * step < hi) reinterpret_cast<float4*>(a.buf[p])[i0 + u * step] = acc[u];
  }
  }
  // 3. done handshake: my pushes are visible system-wide before any peer sees the flag
  __threadfence_system();
  __syncthreads();
  if (rec) rec[2] = globaltimer_ns();
  if (tid < a.world) {
    st_release_sys(a.flags[tid] + (P2P_MAX_WORLD + a.rank) * P2P_MAX_CTAS + b, epoch);
    const uint32_t* w = my + (P2P_MAX_WORLD + tid) * P2P_MAX_CTAS + b;
    uint32_t spins = 0;
    while ((int)(ld_acquire_sys(w) - epoch) < 0) {
      if (++spins > (1u << 28)) { printf("mmf p2p: done wait timed out (rank %d cta %d peer %d)\n", a.rank, b, tid); __trap(); }
    }
  }
  __syncthreads();
  if (rec) rec[3] = globaltimer_ns();
  if (tid == 0) my[2 * P2P_MAX_WORLD * P2P_MAX_CTAS + b] = epoch;
}

}  // namespace mmf
