// mmf_ptx.cuh — thin inline-PTX layer for sm_100a (B200): mbarrier, TMA, tcgen05/TMEM.
//
// Everything the tensor-core kernels of this library need from the Blackwell ISA lives
// here, so each kernel file reads as an algorithm and not as an assembler listing.
// Conventions:
//   * shared-memory addresses are passed as 32-bit .shared::cta window offsets
//     (`smem_u32(ptr)`), barriers included;
//   * every spin loop has a deadlock guard (MMF_SPIN_LIMIT): a kernel that would hang
//     traps instead, so a descriptor or phase bug costs an error code, not a GPU box;
//   * operand tiles are always [rows][64 bf16] = 128-byte rows with the 128-byte TMA /
//     UMMA swizzle (16-byte chunk index XOR (row & 7)), 1024-byte aligned.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mmf {

#ifndef MMF_SPIN_LIMIT
#define MMF_SPIN_LIMIT (1u << 24)   // try_wait polls before trapping (each poll can sleep in HW)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrive_count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrive_count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Blocks until the barrier phase with the given parity has completed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > MMF_SPIN_LIMIT) {
      printf("mmf: mbarrier deadlock guard tripped (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// generic-proxy writes to smem -> visible to the async proxy (TMA store / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barrier among a subset of the CTA's warps
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------
// Debug timeline (only in builds with -DMMF_DEBUG_TIMELINE=1, tools/step_timeline.py; the release library carries no
// device-global state and mmf_debug_set_timeline_buffer is a no-op there): every CTA of every hot-path kernel appends one
// record (kernel id, blockIdx.x | cycles << 32, %globaltimer at CTA start, after griddepcontrol.wait, at CTA end) to a log.
// Buffer layout (uint64): [0] record count, [16 + 5 i ...] = the i-th record.
// ----------------------------------------------------------------------------------------
#ifndef MMF_DEBUG_TIMELINE
#define MMF_DEBUG_TIMELINE 0
#endif
struct Timeline { unsigned long long t0, t1; long long c1; };
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#if MMF_DEBUG_TIMELINE
__device__ unsigned long long* d_timeline = nullptr;
constexpr unsigned long long TIMELINE_MAX_RECORDS = 32768;
__device__ __forceinline__ Timeline timeline_start(int) {
  Timeline tl = {0ull, 0ull, 0ll};
  if (d_timeline && threadIdx.x == 0) tl.t0 = globaltimer_ns();
  return tl;
}
__device__ __forceinline__ void timeline_wait_done(Timeline& tl) {
  if (d_timeline && threadIdx.x == 0) { tl.t1 = globaltimer_ns(); tl.c1 = clock64(); }
}
__device__ __forceinline__ void timeline_end(int id, const Timeline& tl) {
  if (d_timeline && threadIdx.x == 0) {
    const unsigned long long t2 = globaltimer_ns();
    const long long c2 = clock64();
    const unsigned long long i = atomicAdd(d_timeline, 1ull);
    if (i < TIMELINE_MAX_RECORDS) {
      unsigned long long* r = d_timeline + 16 + 5 * i;
      // [1]: blockIdx.x in the low 32 bits, SM cycles between the wait-return and the end in the high 32 bits
      r[0] = (unsigned long long)id; r[1] = (unsigned long long)blockIdx.x | ((unsigned long long)(c2 - tl.c1) << 32);
      r[2] = tl.t0; r[3] = tl.t1; r[4] = t2;
    }
  }
}
#else
__device__ __forceinline__ Timeline timeline_start(int) { return Timeline{0ull, 0ull, 0ll}; }
__device__ __forceinline__ void timeline_wait_done(Timeline&) {}
__device__ __forceinline__ void timeline_end(int, const Timeline&) {}
#endif

// ----------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). A kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream
// is still draining: its CTAs run their prologue (barrier init, TMEM allocation, descriptor prefetch)
// on SMs the predecessor has already left and block in griddep_wait() until the predecessor has
// completed and its writes are visible. Both are no-ops for a normally launched kernel.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// IMPORTANT: nvcc treats ld.global.nc (__ldg, loads through const __restrict__ pointers) as loads of immutable memory and
// schedules them freely — also ABOVE the volatile, memory-clobbering griddepcontrol.wait (seen in the SASS of the head
// kernel: the (m_t, l_t) loads of the forward's partials sat before the wait and intermittently returned the previous
// contents of a recycled buffer). Every pointer to data written by the preceding kernel goes through pdl_fresh() after
// the wait: the address then depends on a volatile asm that cannot move above it.
template <class T>
__device__ __forceinline__ T* pdl_fresh(T* p) {
  asm volatile("" : "+l"(p) : : "memory");
  return p;
}
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiles
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared, completes `bytes` on the mbarrier. c0 = innermost (contiguous) coordinate.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// same with an L2 eviction-priority hint (createpolicy value)
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst_smem, const CUtensorMap* m,
                                                 uint32_t bar, int32_t c0, int32_t c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// shared -> global (bulk async group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int32_t c0,
                                             int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global with element-wise ADD at the destination (split-K reduction done by the TMA / L2)
// MMF_L2_HINTS >= 2 (A/B candidate): the step's stash / dG / dU stores carry evict_last
#ifndef MMF_L2_HINTS
#define MMF_L2_HINTS 1
#endif
__device__ __forceinline__ void tma_store_2d_keep(const CUtensorMap* m, uint32_t src_smem, int32_t c0, int32_t c1) {
#if MMF_L2_HINTS >= 2
  const uint64_t policy = l2_policy_evict_last();
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
#else
  tma_store_2d(m, src_smem, c0, c1);
#endif
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src_smem, int32_t c0,
                                                  int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the smem source of all committed stores has been read (smem reusable)
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// wait until all committed stores are complete (global writes performed)
// Before a CTA exits (or reuses the staging memory) its bulk stores only have to have READ shared memory: the writes of
// a grid's bulk-async stores are performed before the grid completes (and before a dependent grid's griddepcontrol.wait
// returns). Waiting for full completion instead kept every CTA alive for an extra L2 write round trip at each kernel's tail.
__device__ __forceinline__ void tma_store_wait_exit() { tma_store_wait_read(); }
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// 1-D bulk copies (no tensor map): contiguous spans, 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(src_smem), "r"(bytes)
               : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------
// One full warp must execute alloc / dealloc. ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Shared-memory matrix descriptor for a tile of 128-byte rows, 128B-swizzled (UMMA
// LayoutType::SWIZZLE_128B = 2, descriptor version 1 for sm_100).
//   K-major operand  : rows are M/N indices, the 64 bf16 of a row are K. SBO = 1024 B
//                      (8-row group pitch); LBO unused for swizzled K-major (set to 1).
//   MN-major operand : rows are K indices, the 64 bf16 of a row are M/N. SBO = 1024 B
//                      (8-k group pitch), LBO = byte pitch between consecutive 64-wide
//                      M/N blocks.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                // [46,48) descriptor version = 1
  d |= (uint64_t)2 << 61;                                // [61,64) SWIZZLE_128B
  return d;
}

// Instruction descriptor, kind::f16, BF16 x BF16 -> FP32, dense.
//   a_mn / b_mn : 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4)                    // c_format  = F32
         | (1u << 7)                  // a_format  = BF16
         | (1u << 10)                 // b_format  = BF16
         | ((uint32_t)a_mn << 15)     // a_major
         | ((uint32_t)b_mn << 16)     // b_major
         | ((uint32_t)(N >> 3) << 17) // n_dim
         | ((uint32_t)(M >> 4) << 24);// m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues. accumulate == 0 overwrites D.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}

// TMEM -> registers: the calling warp reads its 32-lane quadrant (lanes 32*(warp%4)...),
// thread i gets lane i of the quadrant, 32 / 16 consecutive fp32 columns from `taddr`.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variant (the z = Wk h_i side product of the training forward)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a cluster (ranks 2k, 2k+1 = one TPC) run ONE MMA
// of M = 256: each CTA holds its 128 rows of A and of the accumulator, and half of B's N rows.
// The even CTA (leader) issues the MMAs; TMA loads of both CTAs complete on the leader's barrier.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// address of the same shared variable in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// wait with cluster-scope acquire (barrier arrived on by the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > MMF_SPIN_LIMIT) {
      printf("mmf: cluster mbarrier deadlock guard tripped (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair peer bit of a shared::cluster address
// TMA load executed by either CTA of the pair; the bytes complete on the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// same with an L2 eviction-priority hint (createpolicy value): evict_first for operands that stream through once
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                                      int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d_hint(const CUtensorMap* m, int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.L2::cache_hint [%0, {%1, %2}], %3;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
// one contiguous block (bytes % 16 == 0) into L2 with an eviction-priority hint: a single instruction for a whole tile
// (a tensor prefetch per 16 KB box costs the issuing thread ~280 cycles each: 16 of them delayed the prologue by 4.5k cycles)
__device__ __forceinline__ void bulk_prefetch_l2_hint(const void* src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the same-offset mbarrier of every CTA in `cta_mask` once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// 128B-swizzled tile addressing (generic-proxy side; must match TMA SWIZZLE_128B and
// umma_desc_sw128). `tile_base` is 1024-byte aligned; a tile is [rows][128 bytes].
// Returns the byte offset of the 16-byte chunk `chunk` (0..7) of row `row`.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// ReLU mask bits from packed bf16: 0xFFFF per 16-bit half of w whose bf16 value is > 0 (false for -0 and NaN)
__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t w) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w), __float2bfloat162_rn(0.f));
}
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// bit i of the result = (element i > 0) for the 8 bf16 values packed in w0..w3 (element 2k = low half of wk):
// 4 HSET2.BF16 + 2 PRMT + 5 integer ops instead of 16 compares + selects on up-converted floats
__device__ __forceinline__ uint32_t relu_mask_byte(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  const uint32_t X = prmt_b32(bf16x2_gt0_mask(w0), bf16x2_gt0_mask(w1), 0x6420u);   // one 0x00 / 0xFF byte per element 0..3
  const uint32_t Y = prmt_b32(bf16x2_gt0_mask(w2), bf16x2_gt0_mask(w3), 0x6420u);   // elements 4..7
  return (((X & 0x08040201u) + ((Y & 0x08040201u) << 4)) * 0x01010101u) >> 24;      // byte sum: no carries
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_b32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// fast activations (MUFU.TANH; sigmoid(x) = 0.5*tanh(0.5x)+0.5 : one MUFU each)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) {
  return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums of a 32x32 register block held as v[32] per lane (lane = row, index = column):
// after the call lane l holds sum over the 32 rows of column l. 31 shuffles.
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      // lanes with bit `half` set keep columns [half, 2*half) of the current window
      float send = upper ? v[j] : v[j + half];
      float keep = upper ? v[j + half] : v[j];
      float got = __shfl_xor_sync(0xffffffffu, send, half);
      v[j] = keep + got;
    }
  }
  return v[0];
}

}  // namespace mmf
