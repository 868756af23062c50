// capi.cu — the extern "C" surface declared in include/mmf_b200.h. Host-side orchestration only:
// argument checks, TMA descriptor construction, workspace carving, kernel launches on the
// caller's stream. No allocation, no synchronisation, no global state.
#include <cuda_bf16.h>

#include "../../include/mmf_b200.h"
#include "amil_tile.cuh"
#include "amil_tile2.cuh"
#include "amil_hidden_fused.cuh"
#include <stdlib.h>
#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "mmf_host.cuh"
#include "small_kernels.cuh"
#include "p2p_allreduce.cuh"
#include "train_glue.cuh"
#include "head_kernels.cuh"
#include "xfusion_gate.cuh"
#include "snn_mlp.cuh"
#include <math.h>

using namespace mmf;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Opt a kernel in to > 48 KB of dynamic shared memory, once per (kernel instantiation, device): the attribute is
// per device, a process may drive several (benign race between host threads: the call is idempotent).
#define MMF_CONFIGURE_SMEM(kern, bytes)                                                                        \
  do {                                                                                                         \
    static bool done_[64] = {};                                                                                \
    int dev_ = 0;                                                                                              \
    MMF_TRY(cuda_rc(cudaGetDevice(&dev_)));                                                                    \
    if (dev_ < 0 || dev_ >= 64 || !done_[dev_]) {                                                              \
      MMF_TRY(cuda_rc(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))); \
      if (dev_ >= 0 && dev_ < 64) done_[dev_] = true;                                                          \
    }                                                                                                          \
  } while (0)

// MMF_DEBUG_STAMPS = 1 (debug builds only: tools/ab_variant.py build stamps -DMMF_DEBUG_STAMPS=1): process-global pointers
// to the clock64 phase-stamp buffer [gridDim.x][16] of the tensor-core kernels (tools/phase_*.py) and to the %globaltimer
// stamps of the peer all-reduce (tools/dp_diag.py). The RELEASE library holds no global state: the setters are no-ops and
// every kernel receives a null stamp pointer.
#ifndef MMF_DEBUG_STAMPS
#define MMF_DEBUG_STAMPS 0
#endif
#if MMF_DEBUG_STAMPS
unsigned long long* g_timing_buffer = nullptr;
// MMF_STAMP_KERNEL=<id> (debug, tools/phase_instep.py): only the kernel with this timeline id (0 forward tile, 2 gate + hidden,
// 3 grouped wgrad, 4 recompute gate, 5 other pair GEMMs) receives the phase-stamp buffer — the kernels of a whole step can
// then run back to back (CUDA graph) while one of them is stamped.
unsigned long long* stamp_buf(int id) {
  static const int only = [] { const char* e = getenv("MMF_STAMP_KERNEL"); return e ? atoi(e) : -1; }();
  return (only < 0 || only == id) ? g_timing_buffer : nullptr;
}
unsigned long long* g_p2p_stamps = nullptr;
#else
inline unsigned long long* stamp_buf(int) { return nullptr; }
constexpr unsigned long long* g_p2p_stamps = nullptr;
#endif

struct BwdWs {
  size_t off_H, off_dG, off_dU, off_cs, off_dbc, off_db1, off_mask, off_z, off_thead, total;
};
BwdWs bwd_layout(int64_t N, int L, int D, int gated) {
  const int64_t tiles = ((N + 255) / 256) * 2;  // padded to whole CTA pairs
  const int KD = gated ? 2 * D : D;
  const int ncols = gated ? 3 * D : 2 * D;
  BwdWs w;
  size_t o = 0;
  w.off_H = o;   o = align_up(o + (size_t)N * L * 2, 1024);
  w.off_dG = o;  o = align_up(o + (size_t)N * KD * 2, 1024);
  w.off_dU = o;  o = align_up(o + (size_t)N * L * 2, 1024);
  w.off_cs = o;  o = align_up(o + (size_t)tiles * 4 * ncols * 4, 1024);
  w.off_dbc = o; o = align_up(o + (size_t)tiles * 4 * 4, 1024);
  w.off_db1 = o; o = align_up(o + (size_t)tiles * 4 * L * 4, 1024);
  w.off_mask = o; o = align_up(o + (size_t)N * (L / 32) * 4, 1024);   // 1 bit per element of H: [h > 0]
  w.off_z = o;    o = align_up(o + (size_t)N * 8 * 4, 1024);          // z_i = Wk h_i (head-projected backward), fp32 [N, 4 | 8]
  w.off_thead = o; o = align_up(o + (size_t)HEAD_MAX_TILES * HEAD_ROW * 4, 1024);    // head rows (m_t, l_t, Wk·acc_t) per tile
  w.total = o;
  return w;
}

template <int A_MN, int B_MN, int EPI>
int launch_gemm(const TMapSet& tmA, const TMapSet& tmB, const GemmArgs& g, int splits, cudaStream_t st) {
  auto kern = gemm_tc_kernel<A_MN, B_MN, EPI>;
  MMF_CONFIGURE_SMEM(kern, GEMM_SMEM_BYTES);
  dim3 grid((g.M + GEMM_BM - 1) / GEMM_BM, (g.N + GEMM_BN - 1) / GEMM_BN, splits);
  kern<<<grid, 256, GEMM_SMEM_BYTES, st>>>(tmA, tmB, g);
  return launch_status();
}

template <int A_MN, int B_MN, int EPI, int BN>
int launch_gemm2(const TMapSet& tmA, const TMapSet& tmB, const GemmArgs& g, int splits, cudaStream_t st) {
  using C = Gemm2Cfg<BN>;
  auto kern = gemm2_tc_kernel<A_MN, B_MN, EPI, BN>;
  MMF_CONFIGURE_SMEM(kern, C::SMEM_BYTES);
  dim3 grid(2 * ((g.M + 255) / 256), (g.N + BN - 1) / BN, splits);
  return launch_pdl(kern, grid, dim3(GEMM2_THREADS), C::SMEM_BYTES, st, tmA, tmB, g);
}

// Grouped split-K launch of the pair kernel (both operands MN-major, EPI_ATOMIC, 256 x 512 tiles): up to two
// problems over the same reduction axis share ONE wave of <= 74 CTA pairs (148 SMs). 256 x 512 tiles keep the
// per-CTA operand stream at 48 KB per 1024 MMA cycles (47 B/clk, at the ~43 B/clk/SM L2 cap) where the
// 256 x 256 tiles of two back-to-back launches needed 64 B/clk and ran the tensor pipe at 35-40 %.
int launch_gemm2_grouped(const TMapSet& tmA, const TMapSet& tmB, GemmArgs ga, cudaStream_t st) {
  constexpr int BN = 512;
  using C = Gemm2Cfg<BN>;
  auto kern = gemm2_tc_kernel<1, 1, EPI_ATOMIC, BN>;
  MMF_CONFIGURE_SMEM(kern, C::SMEM_BYTES);
  int tiles_total = 0;
  for (int p = 0; p < ga.n_groups; ++p) {
    GemmArgs::Group& P = ga.grp[p];
    P.tiles_n = (P.N + BN - 1) / BN;
    tiles_total += ((P.M + 255) / 256) * P.tiles_n;
  }
  // MMF_WGRAD_PAIR_SLOTS (diagnostic): CTA-pair slots to fill. 74 = one full wave (lowest latency of a lone launch);
  // fewer slots = fewer, longer split-K slices (less reduction traffic and prologue per FLOP) for concurrent lanes
  static const int slots = [] { const char* e = getenv("MMF_WGRAD_PAIR_SLOTS"); int v = e ? atoi(e) : 74; return v > 0 ? v : 74; }();
  int splits = slots / tiles_total;
  if (splits < 1) splits = 1;
  if (splits > ga.kb_total) splits = ga.kb_total;
  const int per = (ga.kb_total + splits - 1) / splits;
  splits = (ga.kb_total + per - 1) / per;
  int pairs = 0;
  for (int p = 0; p < ga.n_groups; ++p) {
    GemmArgs::Group& P = ga.grp[p];
    P.splits = splits; P.kb_per_split = per; P.first_pair = pairs;
    pairs += ((P.M + 255) / 256) * P.tiles_n * splits;
  }
  ga.dbg = stamp_buf(3);
  return launch_pdl(kern, dim3(2 * pairs), dim3(GEMM2_THREADS), C::SMEM_BYTES, st, tmA, tmB, ga);
}

// split-K factor for the pair kernel: fill the 74 CTA-pair slots once
int pick_splits_pair(int out_tiles, int kb_total, int* kb_per_split) {
  int splits = 74 / out_tiles;   // floor: one wave of CTA pairs (148 SMs = 74 pairs)
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  int per = (kb_total + splits - 1) / splits;
  splits = (kb_total + per - 1) / per;
  *kb_per_split = per;
  return splits;
}

// CTA-pair kernel (amil_tile2.cuh): grid = 2 * ceil(N / 256), cluster (2,1,1)
template <int L, int D, bool GATED, int MODE, bool DROPH, bool DROPA>
int launch_amil2v(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, const AmilArgs& a,
                  void* Hbuf, const void* wk_split, cudaStream_t st) {
  using C = Amil2Cfg<L, D, GATED>;
  auto kern = amil_tile2_kernel<L, D, GATED, MODE, DROPH, DROPA>;
  MMF_CONFIGURE_SMEM(kern, C::SMEM_BYTES);
  CUtensorMap tmX, tmW1, tmWab, tmH, tmWk, tmAGs;
  const bool precise = (a.flags & MMF_PRECISE_FC) != 0;
  const uint64_t kin = precise ? 3072 : 1024;
  if (precise && (!w->W1_split || ldx < 3072)) return MMF_E_INVALID;
  MMF_TRY(make_tmap_bf16(&tmX, x, (uint64_t)N, kin, (uint64_t)ldx, 128));
  MMF_TRY(make_tmap_bf16(&tmW1, precise ? w->W1_split : w->W1, L, kin, kin, 128));
  MMF_TRY(make_tmap_bf16(&tmWab, w->Wab_packed, (uint64_t)C::NCH * C::CHN, L, L, C::CHN / 2));
  if (Hbuf) MMF_TRY(make_tmap_bf16(&tmH, Hbuf, (uint64_t)N, L, L, 128));
  else tmH = tmX;
  if (wk_split) MMF_TRY(make_tmap_bf16(&tmWk, wk_split, 16, L, L, 8));   // [Wk_hi ; Wk_lo], 8 rows per CTA of the pair
  else tmWk = tmX;
  if (MODE == AMIL_FWD && a.AG) MMF_TRY(make_tmap_2b_sw64(&tmAGs, a.AG, (uint64_t)N, (uint64_t)a.ldag, (uint64_t)a.ldag, 32));
  else tmAGs = tmX;
  const int pairs = (int)((N + 255) / 256);
  AmilArgs a2 = a;
  a2.kb1 = (int)(kin / 64);
  a2.x_bulk = ((uint64_t)ldx == kin) ? x : nullptr;
  return launch_pdl(kern, dim3(2 * pairs), dim3(AMIL2_THREADS), C::SMEM_BYTES, st, tmX, tmW1, tmWab, tmH, tmWk, tmAGs, a2);
}

template <int L, int D, bool GATED, int MODE>
int launch_amil2(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, const AmilArgs& a,
                 void* Hbuf, const void* wk_split, cudaStream_t st) {
  const bool dh = (a.flags & MMF_DROPOUT_H) != 0, da = (a.flags & MMF_DROPOUT_ATTN) != 0;
  if (dh) return da ? launch_amil2v<L, D, GATED, MODE, true, true>(x, N, ldx, w, a, Hbuf, wk_split, st)
                    : launch_amil2v<L, D, GATED, MODE, true, false>(x, N, ldx, w, a, Hbuf, wk_split, st);
  return da ? launch_amil2v<L, D, GATED, MODE, false, true>(x, N, ldx, w, a, Hbuf, wk_split, st)
            : launch_amil2v<L, D, GATED, MODE, false, false>(x, N, ldx, w, a, Hbuf, wk_split, st);
}

template <int MODE>
int dispatch_amil(int L, int D, int gated, const void* x, int64_t N, int64_t ldx,
                  const MmfAmilWeights* w, const AmilArgs& a, void* Hbuf, cudaStream_t st, const void* wk_split = nullptr) {
#define MMF_CASE(LL, DD)                                                                  \
  if (L == LL && D == DD)                                                                    \
    return gated ? launch_amil2<LL, DD, true, MODE>(x, N, ldx, w, a, Hbuf, wk_split, st)    \
                 : launch_amil2<LL, DD, false, MODE>(x, N, ldx, w, a, Hbuf, wk_split, st);
  MMF_CASE(256, 256)
  MMF_CASE(512, 384)
  MMF_CASE(256, 384)
#undef MMF_CASE
  return MMF_E_UNSUPPORTED;
}

int check_amil_common(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D) {
  if (!x || !w || N <= 0 || ldx < 1024) return MMF_E_INVALID;
  if (!w->W1 || !w->b1 || !w->Wab || !w->Wab_packed || !w->bab || !w->wc || !w->bc) return MMF_E_INVALID;
  if (!((L == 256 && D == 256) || (L == 512 && D == 384) || (L == 256 && D == 384))) return MMF_E_UNSUPPORTED;
  if (N > 0x7fffff00LL) return MMF_E_INVALID;
  return MMF_OK;
}

template <int L, int D, bool GATED, bool DROP, bool HEADPROJ>
int launch_hidden_fused3(const HiddenFusedArgs& a, const CUtensorMap& tmAG, const CUtensorMap& tmWab,
                         const CUtensorMap& tmDU, cudaStream_t st) {
  using C = HiddenFusedCfg<L, D, GATED>;
  auto kern = amil_hidden_fused_kernel<L, D, GATED, DROP, HEADPROJ>;
  MMF_CONFIGURE_SMEM(kern, C::SMEM_BYTES);
  const int pairs = (int)((a.N + 255) / 256);
  return launch_pdl(kern, dim3(2 * pairs), dim3(HIDDEN_THREADS), C::SMEM_BYTES, st, tmAG, tmAG, tmWab, tmDU, a);
}
template <int L, int D, bool GATED, bool DROP>
int launch_hidden_fused2(const HiddenFusedArgs& a, const CUtensorMap& tmAG, const CUtensorMap& tmWab,
                         const CUtensorMap& tmDU, cudaStream_t st) {
  return a.z ? launch_hidden_fused3<L, D, GATED, DROP, true>(a, tmAG, tmWab, tmDU, st)
             : launch_hidden_fused3<L, D, GATED, DROP, false>(a, tmAG, tmWab, tmDU, st);
}
template <int L, int D, bool GATED>
int launch_hidden_fused(const HiddenFusedArgs& a, int flags, const CUtensorMap& tmAG, const CUtensorMap& tmWab,
                        const CUtensorMap& tmDU, cudaStream_t st) {
  return (flags & MMF_DROPOUT_ATTN) ? launch_hidden_fused2<L, D, GATED, true>(a, tmAG, tmWab, tmDU, st)
                                    : launch_hidden_fused2<L, D, GATED, false>(a, tmAG, tmWab, tmDU, st);
}


}  // namespace

extern "C" {

int mmf_version(void) { return MMF_ABI_VERSION; }

void mmf_debug_set_timing_buffer(void* device_u64_buffer) {
#if MMF_DEBUG_STAMPS
  g_timing_buffer = reinterpret_cast<unsigned long long*>(device_u64_buffer);
#else
  (void)device_u64_buffer;   // release build: no global state (rebuild with -DMMF_DEBUG_STAMPS=1)
#endif
}

void mmf_debug_set_p2p_stamp_buffer(void* device_u64_buffer) {
#if MMF_DEBUG_STAMPS
  g_p2p_stamps = reinterpret_cast<unsigned long long*>(device_u64_buffer);
#else
  (void)device_u64_buffer;
#endif
}

/* 1 when this library was compiled with -DMMF_DEBUG_STAMPS=1 (the phase-stamp tools refuse to run otherwise). */
int mmf_debug_stamps_enabled(void) { return MMF_DEBUG_STAMPS; }

void mmf_debug_set_timeline_buffer(void* device_u64_buffer) {
#if MMF_DEBUG_TIMELINE
  unsigned long long* p = reinterpret_cast<unsigned long long*>(device_u64_buffer);
  cudaMemcpyToSymbol(d_timeline, &p, sizeof(p));
#else
  (void)device_u64_buffer;   // release build: no device-global debug state (rebuild with -DMMF_DEBUG_TIMELINE=1)
#endif
}

const char* mmf_error_string(int rc) {
  switch (rc) {
    case MMF_OK: return "ok";
    case MMF_E_INVALID: return "invalid argument";
    case MMF_E_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case MMF_E_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    case MMF_E_TMAP: return "tensor map encode failed";
    case MMF_E_UNSUPPORTED: return "unsupported configuration";
    case MMF_E_WORKSPACE: return "workspace too small";
    default:
      if (rc <= -1000) return cudaGetErrorString((cudaError_t)(-rc - 1000));
      return "unknown error";
  }
}

int mmf_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return MMF_E_INVALID;
  if (n == 0) return MMF_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(dst) & 7u)) return MMF_E_ALIGN;
  const long long n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n4, src + n4 * 4,
      reinterpret_cast<__nv_bfloat16*>(dst) + n4 * 4, tail);
  return launch_status();
}

int mmf_split_f32_bf16x3(const float* x, int64_t n_rows, int64_t ldx, void* out, void* stream) {
  if (!x || !out || n_rows <= 0 || ldx < 1024) return MMF_E_INVALID;
  split_f32_bf16x3_kernel<<<(int)((n_rows * 1024 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, (long long)n_rows, (long long)ldx, reinterpret_cast<__nv_bfloat16*>(out));
  return launch_status();
}

int mmf_pack_wab(const void* Wab, void* packed, int L, int D, int gated, void* stream) {
  if (!Wab || !packed || L % 8 || D % 128) return MMF_E_INVALID;
  pack_wab_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(Wab),
                                                         reinterpret_cast<uint4*>(packed), L, D, gated);
  return launch_status();
}

int64_t mmf_amil_num_tiles(int64_t N) { return (N + MMF_TILE_ROWS - 1) / MMF_TILE_ROWS; }

int mmf_amil_fwd(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                 int flags, uint64_t seed, float* A_raw, float* partials, void* H_stash, void* stream) {
  MMF_TRY(check_amil_common(x, N, ldx, w, L, D));
  if (!A_raw || !partials) return MMF_E_INVALID;
  AmilArgs a = {};
  a.N = N; a.b1 = w->b1; a.bab = w->bab; a.wc = w->wc; a.bc = w->bc;
  a.A_raw = A_raw; a.partials = partials; a.store_h = H_stash != nullptr;
  a.flags = flags; a.seed = seed; a.dbg = stamp_buf(0);
  return dispatch_amil<AMIL_FWD>(L, D, flags & MMF_GATED, x, N, ldx, w, a, H_stash, (cudaStream_t)stream);
}

// Cohort inference: ONE fused-forward launch over a packed multi-bag buffer + one head launch for all bags.
int mmf_amil_infer_varlen(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                          const int32_t* tile_valid, const int32_t* seg_tile_offsets, int n_bags, const float* Wk,
                          const float* bk, int K, float* A_raw, float* partials, float* M, float* ml, float* hazards,
                          float* S, float* risk, int64_t* Y_hat, void* stream) {
  MMF_TRY(check_amil_common(x, R, ldx, w, L, D));
  if (!tile_valid || !seg_tile_offsets || n_bags <= 0 || !Wk || !bk || !A_raw || !partials || !M || !hazards || !S)
    return MMF_E_INVALID;
  if (K <= 0 || K > 16 || L > 1024 || (R % 128) != 0) return MMF_E_UNSUPPORTED;
  if (flags & (MMF_DROPOUT_H | MMF_DROPOUT_ATTN)) return MMF_E_INVALID;   // inference only
  AmilArgs a = {};
  a.N = R; a.b1 = w->b1; a.bab = w->bab; a.wc = w->wc; a.bc = w->bc;
  a.A_raw = A_raw; a.partials = partials; a.tile_valid = tile_valid;
  a.flags = flags; a.dbg = stamp_buf(0);
  MMF_TRY(dispatch_amil<AMIL_FWD>(L, D, flags & MMF_GATED, x, R, ldx, w, a, nullptr, (cudaStream_t)stream));
  amil_seg_head_kernel<<<n_bags, 256, 0, (cudaStream_t)stream>>>(partials, seg_tile_offsets, L, Wk, bk, K, M, ml, hazards,
                                                                 S, risk, reinterpret_cast<long long*>(Y_hat));
  return launch_status();
}

int mmf_amil_combine(const float* partials, int64_t n, int L, int normalize, float* out, float* ml,
                     void* stream) {
  if (!partials || !out || n <= 0 || L <= 0 || (normalize && !ml)) return MMF_E_INVALID;
  amil_combine_kernel<<<(L + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(partials, n, L, normalize, out, ml);
  return launch_status();
}

size_t mmf_amil_bwd_workspace_bytes(int64_t N, int L, int D, int flags) {
  if (N <= 0) return 0;
  return bwd_layout(N, L, D, flags & MMF_GATED).total;
}

}  // extern "C"

namespace {
int check_head(const MmfHeadStep* h, int64_t N) {
  if (!h->Wk || !h->bk || !h->Wk_split || !h->Y || !h->c || !h->M || !h->ml || !h->hazards || !h->S || !h->loss ||
      !h->dM || !h->hs)
    return MMF_E_INVALID;
  if (h->K <= 0 || h->K > HEAD_MAX_K) return MMF_E_UNSUPPORTED;
  if ((N + 127) / 128 > HEAD_MAX_TILES) return MMF_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(h->Wk_split) & 15u) return MMF_E_ALIGN;
  return MMF_OK;
}

// Training forward: mmf_amil_fwd that also leaves H (bf16 [N,L]), the branch activations (fp16 [N,KD], in the dG
// slot) and the ReLU mask words in the backward workspace, for mmf_amil_bwd(... | MMF_STASHED); with a head block
// additionally z = Wk h (fp32 [N, 4|8]) and the folded head step run by the last tile CTA.
int fwd_train_impl(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                   uint64_t seed, float* A_raw, float* partials, void* workspace, size_t workspace_bytes,
                   float* zero_buf, int64_t zero_count, const MmfHeadStep* head, void* stream,
                   const int32_t* tile_valid = nullptr) {
  MMF_TRY(check_amil_common(x, N, ldx, w, L, D));
  if (!A_raw || !partials || !workspace) return MMF_E_INVALID;
  if (tile_valid && (head || (N & 127))) return MMF_E_INVALID;   // packed windows: whole 128-row tiles, general head
  const int gated = flags & MMF_GATED;
  const BwdWs lay = bwd_layout(N, L, D, gated);
  if (workspace_bytes < lay.total) return MMF_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023u) return MMF_E_ALIGN;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  AmilArgs a = {};
  a.N = N; a.b1 = w->b1; a.bab = w->bab; a.wc = w->wc; a.bc = w->bc;
  a.A_raw = A_raw; a.partials = partials; a.store_h = 1; a.tile_valid = tile_valid;
  a.AG = reinterpret_cast<uint16_t*>(ws + lay.off_dG); a.ldag = gated ? 2 * D : D;
  a.mask_out = reinterpret_cast<uint32_t*>(ws + lay.off_mask);
  // (the dU slot of the workspace: dead between the previous step's wgrad and this step's hidden-gradient kernel)
  a.discard_ptr = ws + lay.off_dU; a.discard_n128 = (long long)N * L * 2 / 128;
  a.h_stash = ws + lay.off_H;
  if (zero_buf) {
    if ((reinterpret_cast<uintptr_t>(zero_buf) & 15u) || zero_count < 0 || (zero_count & 3)) return MMF_E_ALIGN;
    a.zero_ptr = reinterpret_cast<float4*>(zero_buf); a.zero_n4 = zero_count >> 2;
  }
  a.flags = flags; a.seed = seed; a.dbg = stamp_buf(0);
  const void* wk_split = nullptr;
  if (head) {
    MMF_TRY(check_head(head, N));
    wk_split = head->Wk_split;
    a.z_out = reinterpret_cast<float*>(ws + lay.off_z);
    a.zld = head->K <= 4 ? 4 : 8;
    a.tile_head = reinterpret_cast<float*>(ws + lay.off_thead);
    a.head_wk = head->Wk; a.head_k = head->K;
  }
  return dispatch_amil<AMIL_FWD>(L, D, gated, x, N, ldx, w, a, ws + lay.off_H, (cudaStream_t)stream, wk_split);
}
}  // namespace

extern "C" {

int mmf_amil_fwd_train(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                       int flags, uint64_t seed, float* A_raw, float* partials, void* workspace,
                       size_t workspace_bytes, float* zero_buf, int64_t zero_count, void* stream) {
  return fwd_train_impl(x, N, ldx, w, L, D, flags, seed, A_raw, partials, workspace, workspace_bytes, zero_buf,
                        zero_count, nullptr, stream);
}

int mmf_amil_fwd_train_head(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                            int flags, uint64_t seed, float* A_raw, float* partials, void* workspace,
                            size_t workspace_bytes, float* zero_buf, int64_t zero_count, const MmfHeadStep* head,
                            void* stream) {
  if (!head) return MMF_E_INVALID;
  return fwd_train_impl(x, N, ldx, w, L, D, flags, seed, A_raw, partials, workspace, workspace_bytes, zero_buf,
                        zero_count, head, stream);
}

int mmf_pack_head_weights(const float* Wk, int K, int L, void* Wk_split, void* stream) {
  if (!Wk || !Wk_split || K <= 0 || K > HEAD_MAX_K || L <= 0 || (L & 7)) return MMF_E_INVALID;
  pack_head_weights_kernel<<<(16 * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      Wk, K, L, reinterpret_cast<__nv_bfloat16*>(Wk_split));
  return launch_status();
}

namespace {
struct BwdCtx {
  BwdWs lay; int KD, ncols, gated; int64_t tiles;
  __nv_bfloat16 *Hb, *dG, *dU; float *cs, *dbc_ws, *db1_ws; uint32_t* mask;
};
int bwd_ctx(BwdCtx* c, const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
            int flags, void* workspace, size_t workspace_bytes) {
  MMF_TRY(check_amil_common(x, N, ldx, w, L, D));
  if (!workspace) return MMF_E_INVALID;
  c->gated = flags & MMF_GATED;
  c->lay = bwd_layout(N, L, D, c->gated);
  if (workspace_bytes < c->lay.total) return MMF_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023u) return MMF_E_ALIGN;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  c->Hb = reinterpret_cast<__nv_bfloat16*>(ws + c->lay.off_H);
  c->dG = reinterpret_cast<__nv_bfloat16*>(ws + c->lay.off_dG);
  c->dU = reinterpret_cast<__nv_bfloat16*>(ws + c->lay.off_dU);
  c->cs = reinterpret_cast<float*>(ws + c->lay.off_cs);
  c->dbc_ws = reinterpret_cast<float*>(ws + c->lay.off_dbc);
  c->db1_ws = reinterpret_cast<float*>(ws + c->lay.off_db1);
  c->mask = reinterpret_cast<uint32_t*>(ws + c->lay.off_mask);
  c->KD = c->gated ? 2 * D : D;
  c->ncols = c->gated ? 3 * D : 2 * D;
  c->tiles = (N + 127) / 128;
  return MMF_OK;
}
}  // namespace

// Stage 1: row-stationary pass — recompute H and the attention branches tile by tile, emit
// dG (bf16 [N,2D]) and H (bf16 [N,L]) into the workspace, reduce dwc / dbab / dbc.
int mmf_amil_bwd_gate(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                      int flags, uint64_t seed, const float* A_raw, const float* ml, const float* M,
                      const float* dM, const float* dA_raw, const MmfAmilGrads* g, void* workspace,
                      size_t workspace_bytes, void* stream) {
  BwdCtx c;
  MMF_TRY(bwd_ctx(&c, x, N, ldx, w, L, D, flags, workspace, workspace_bytes));
  if (!A_raw || !ml || !M || !dM || !g || !g->dbab || !g->dwc || !g->dbc) return MMF_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  AmilArgs a = {};
  a.N = N; a.b1 = w->b1; a.bab = w->bab; a.wc = w->wc; a.bc = w->bc;
  a.A_raw = const_cast<float*>(A_raw); a.flags = flags; a.seed = seed;
  a.ml = ml; a.M = M; a.dM = dM; a.dA_raw = dA_raw;
  a.dG = c.dG; a.lddg = c.KD; a.colsum_ws = c.cs; a.dbc_ws = c.dbc_ws; a.dbg = stamp_buf(4);
  MMF_TRY(dispatch_amil<AMIL_BWD_GATE>(L, D, c.gated, x, N, ldx, w, a, c.Hb, st));
  // the pair dU GEMM reads [h > 0] as a bitmask (the training forward of the stash path emits it itself)
  relu_mask_kernel<<<(int)((N * (L / 32) + 255) / 256), 256, 0, st>>>(c.Hb, N * (L / 32), c.mask);
  MMF_TRY(launch_status());
  ReduceSegs segs = {};
  segs.n = 3;
  segs.s[0] = ReduceSeg{c.cs, c.ncols, D, g->dwc};
  segs.s[1] = ReduceSeg{c.cs + D, c.ncols, c.KD, g->dbab};
  segs.s[2] = ReduceSeg{c.dbc_ws, 1, 1, g->dbc};
  launch_reduce_rows(segs, c.tiles * 4, st);
  return launch_status();
}

// Stages 1 + 2 of the MMF_STASHED backward fused: the gate backward is the A-operand producer of the dU GEMM
// (amil_hidden_fused.cuh). Leaves dG and dU in the workspace for the wgrad stage; accumulates dwc, dbab, dbc, db1.
}  // extern "C"
namespace {
int bwd_gate_hidden_stashed_impl(int64_t N, const MmfAmilWeights* w, int L, int D, int flags, uint64_t seed,
                                 const float* A_raw, const float* ml, const float* M, const float* dM,
                                 const float* dA_raw, const MmfAmilGrads* g, void* workspace,
                                 size_t workspace_bytes, const MmfHeadStep* head, const float* partials, void* stream,
                                 const int32_t* tile_bag = nullptr, const int32_t* tile_valid = nullptr) {
  if (!w || !w->wc || !w->Wab || N <= 0 || !workspace) return MMF_E_INVALID;
  if ((tile_bag != nullptr) != (tile_valid != nullptr) || (tile_bag && (head || (N & 127)))) return MMF_E_INVALID;
  if (!((L == 256 && D == 256) || (L == 512 && D == 384) || (L == 256 && D == 384))) return MMF_E_UNSUPPORTED;
  if (!A_raw || !g || !g->dbab || !g->dwc || !g->dbc || !g->db1) return MMF_E_INVALID;
  if (!head && (!ml || !M || !dM)) return MMF_E_INVALID;
  if (N > 0x7fffff00LL) return MMF_E_INVALID;
  const int gated = flags & MMF_GATED;
  const BwdWs lay = bwd_layout(N, L, D, gated);
  if (workspace_bytes < lay.total) return MMF_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023u) return MMF_E_ALIGN;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int KD = gated ? 2 * D : D;
  CUtensorMap tmAG, tmWab, tmDU;
  MMF_TRY(make_tmap_bf16(&tmAG, ws + lay.off_dG, (uint64_t)N, KD, KD, 128));   // fp16 in / bf16 out: 2-byte elements
  MMF_TRY(make_tmap_bf16(&tmWab, w->Wab, KD, L, L, 64));
  MMF_TRY(make_tmap_bf16(&tmDU, ws + lay.off_dU, (uint64_t)N, L, L, 128));
  HiddenFusedArgs a = {};
  a.N = N; a.H = reinterpret_cast<const __nv_bfloat16*>(ws + lay.off_H);
  a.A_raw = A_raw; a.ml = ml; a.M = M; a.dM = dM; a.dA_raw = dA_raw; a.wc = w->wc;
  a.dwc = g->dwc; a.dbab = g->dbab; a.dbc = g->dbc; a.db1 = g->db1;
  a.mask = reinterpret_cast<const uint32_t*>(ws + lay.off_mask);
  if (head) {   // the step's head runs in the kernel's prologue; z was left in the workspace by mmf_amil_fwd_train_head
    MMF_TRY(check_head(head, N));
    if (!partials) return MMF_E_INVALID;
    a.z = reinterpret_cast<const float*>(ws + lay.off_z); a.zld = head->K <= 4 ? 4 : 8;
    a.partials = partials; a.n_tiles = (int)((N + 127) / 128);
    a.tile_head = reinterpret_cast<const float*>(ws + lay.off_thead);
    HeadTail& t = a.head;
    t.Wk = head->Wk; t.bk = head->bk; t.Y = reinterpret_cast<const long long*>(head->Y); t.c = head->c;
    t.alpha = head->alpha; t.eps = head->eps; t.loss_scale = head->loss_scale; t.K = head->K;
    t.M = head->M; t.ml = head->ml; t.hazards = head->hazards; t.S = head->S;
    t.Y_hat = reinterpret_cast<long long*>(head->Y_hat); t.loss = head->loss; t.dM = head->dM; t.hs = head->hs;
    t.dWk = head->dWk; t.dbk = head->dbk;
  }
  a.du_scale = (flags & MMF_DROPOUT_H) ? (1.0f / 0.75f) : 1.0f;
  a.tile_bag = tile_bag; a.tile_valid = tile_valid;
  a.seed = seed; a.dbg = stamp_buf(2);
  cudaStream_t st = (cudaStream_t)stream;
  if (L == 256 && D == 256) return gated ? launch_hidden_fused<256, 256, true>(a, flags, tmAG, tmWab, tmDU, st) : launch_hidden_fused<256, 256, false>(a, flags, tmAG, tmWab, tmDU, st);
  if (L == 512 && D == 384) return gated ? launch_hidden_fused<512, 384, true>(a, flags, tmAG, tmWab, tmDU, st) : launch_hidden_fused<512, 384, false>(a, flags, tmAG, tmWab, tmDU, st);
  return gated ? launch_hidden_fused<256, 384, true>(a, flags, tmAG, tmWab, tmDU, st) : launch_hidden_fused<256, 384, false>(a, flags, tmAG, tmWab, tmDU, st);
}
}  // namespace
extern "C" {

int mmf_amil_bwd_gate_hidden_stashed(int64_t N, const MmfAmilWeights* w, int L, int D, int flags, uint64_t seed,
                                     const float* A_raw, const float* ml, const float* M, const float* dM,
                                     const float* dA_raw, const MmfAmilGrads* g, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return bwd_gate_hidden_stashed_impl(N, w, L, D, flags, seed, A_raw, ml, M, dM, dA_raw, g, workspace, workspace_bytes,
                                      nullptr, nullptr, stream);
}

// gate + hidden stage of mmf_amil_bwd_head alone (stage timing / tests)
int mmf_amil_bwd_gate_hidden_head(int64_t N, const MmfAmilWeights* w, int L, int D, int flags, uint64_t seed,
                                  const float* A_raw, const float* partials, const MmfHeadStep* head,
                                  const float* dA_raw, const MmfAmilGrads* g, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!head) return MMF_E_INVALID;
  return bwd_gate_hidden_stashed_impl(N, w, L, D, flags, seed, A_raw, nullptr, nullptr, nullptr, dA_raw, g, workspace,
                                      workspace_bytes, head, partials, stream);
}

// Backward of the fused training step (mmf_amil_fwd_train_head on the same workspace and head block): the gate + hidden
// stage reads t_i = dlogits·z_i instead of the 512-long dot products dM·h_i, then the grouped wgrad GEMM.
int mmf_amil_bwd_head(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                      uint64_t seed, const float* A_raw, const float* partials, const MmfHeadStep* head,
                      const float* dA_raw, const MmfAmilGrads* g, void* dx, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!head) return MMF_E_INVALID;
  if (!g || !g->dW1 || !g->db1 || !g->dWab || !g->dbab || !g->dwc || !g->dbc) return MMF_E_INVALID;
  MMF_TRY(check_amil_common(x, N, ldx, w, L, D));
  MMF_TRY(bwd_gate_hidden_stashed_impl(N, w, L, D, flags, seed, A_raw, nullptr, nullptr, nullptr, dA_raw, g, workspace,
                                       workspace_bytes, head, partials, stream));
  return mmf_amil_bwd_wgrad(x, N, ldx, w, L, D, flags, g, dx, workspace, workspace_bytes, stream);
}

// Stage 2: dU = (dG Wab + p dM^T) ⊙ relu'(H) [* 1/(1-p)] -> workspace (bf16 [N,L]); db1 += colsum.
int mmf_amil_bwd_hidden(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                        int flags, const float* A_raw, const float* ml, const float* dM,
                        const MmfAmilGrads* g, void* workspace, size_t workspace_bytes, void* stream) {
  BwdCtx c;
  MMF_TRY(bwd_ctx(&c, x, N, ldx, w, L, D, flags, workspace, workspace_bytes));
  if (!A_raw || !ml || !dM || !g || !g->db1) return MMF_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TMapSet tA = {}, tB = {};
  MMF_TRY(make_tmap_bf16(&tA.m[0], c.dG, (uint64_t)N, c.KD, c.KD, 128));
  MMF_TRY(make_tmap_bf16(&tB.m[0], w->Wab, c.KD, L, L, 64));
  GemmArgs ga = {};
  ga.M = (int)N; ga.N = L; ga.kb_total = c.KD / 64; ga.kb_per_split = ga.kb_total;
  ga.a_seg_kb = ga.kb_total; ga.b_seg_n = L;
  ga.c_bf16 = c.dU; ga.ldc = L;
  ga.s_raw = A_raw; ga.ml = ml; ga.dM = dM; ga.H = c.Hb; ga.ldh = L; ga.colsum_ws = c.db1_ws;
  ga.du_scale = (flags & MMF_DROPOUT_H) ? (1.0f / 0.75f) : 1.0f;
  ga.mask = c.mask; ga.mask_ld = L / 32; ga.db1 = g->db1; ga.dbg = stamp_buf(5);
  MMF_TRY(make_tmap_bf16(&tA.m[3], c.dU, (uint64_t)N, L, L, 128));   // output map (TMA store of the staged tile)
  if (L == 512) return launch_gemm2<0, 1, EPI_DU, 512>(tA, tB, ga, 1, st);
  return launch_gemm2<0, 1, EPI_DU, 256>(tA, tB, ga, 1, st);
}

// Stage 3: dW1 += dU^T X, dWab += dG^T H (split-K over the instance axis), optional dx = dU W1.
int mmf_amil_bwd_wgrad(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                       int flags, const MmfAmilGrads* g, void* dx, void* workspace,
                       size_t workspace_bytes, void* stream) {
  BwdCtx c;
  MMF_TRY(bwd_ctx(&c, x, N, ldx, w, L, D, flags, workspace, workspace_bytes));
  if (!g || !g->dW1 || !g->dWab) return MMF_E_INVALID;
  if ((flags & MMF_NEED_DX) && !dx) return MMF_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int kb_rows = (int)((N + 63) / 64);
  {
    // dW1 += dU^T X and dWab += dG^T H in one single-wave grouped launch
    TMapSet tA = {}, tB = {};
    MMF_TRY(make_tmap_bf16(&tA.m[0], c.dU, (uint64_t)N, L, L, 64));
    MMF_TRY(make_tmap_bf16(&tA.m[1], c.dG, (uint64_t)N, c.KD, c.KD, 64));
    MMF_TRY(make_tmap_bf16(&tB.m[0], x, (uint64_t)N, 1024, (uint64_t)ldx, 64));
    MMF_TRY(make_tmap_bf16(&tB.m[1], c.Hb, (uint64_t)N, L, L, 64));
    GemmArgs ga = {};
    ga.kb_total = kb_rows;
    ga.n_groups = 2;
    ga.grp[0].M = L; ga.grp[0].N = 1024; ga.grp[0].a_map = 0; ga.grp[0].b_map = 0;
    ga.grp[0].c = g->dW1; ga.grp[0].ldc = 1024;
    ga.grp[0].tma_reduce = 1;
    static const bool l2_hints = [] { const char* e = getenv("MMF_NO_L2_HINTS"); return !(e && e[0] == '1'); }();
    ga.grp[0].b_stream = l2_hints ? 1 : 0;   // x: read once here, evict_first
    MMF_TRY(make_tmap_f32(&tA.m[3], g->dW1, (uint64_t)L, 1024, 1024, 32));
    if (L >= 512) {
      ga.grp[1].M = c.KD; ga.grp[1].N = L; ga.grp[1].a_map = 1; ga.grp[1].b_map = 1;
      ga.grp[1].c = g->dWab; ga.grp[1].ldc = L;
      ga.grp[1].tma_reduce = 1;
      MMF_TRY(make_tmap_f32(&tB.m[3], g->dWab, (uint64_t)c.KD, L, L, 32));
    } else {
      // L = 256 would fill only half of a 512-wide tile: compute dWab^T = H^T dG (M = L, N = KD) and store transposed
      ga.grp[1].M = L; ga.grp[1].N = c.KD; ga.grp[1].a_map = 2; ga.grp[1].b_map = 2; ga.grp[1].trans = 1;
      MMF_TRY(make_tmap_bf16(&tA.m[2], c.Hb, (uint64_t)N, L, L, 64));
      MMF_TRY(make_tmap_bf16(&tB.m[2], c.dG, (uint64_t)N, c.KD, c.KD, 64));
      ga.grp[1].c = g->dWab; ga.grp[1].ldc = L;
    }
    MMF_TRY(launch_gemm2_grouped(tA, tB, ga, st));
  }
  if (flags & MMF_NEED_DX) {
    TMapSet tA = {}, tB = {};
    MMF_TRY(make_tmap_bf16(&tA.m[0], c.dU, (uint64_t)N, L, L, 128));
    MMF_TRY(make_tmap_bf16(&tB.m[0], w->W1, L, 1024, 1024, 64));
    GemmArgs ga = {};
    ga.M = (int)N; ga.N = 1024; ga.kb_total = L / 64; ga.kb_per_split = ga.kb_total;
    ga.a_seg_kb = ga.kb_total; ga.b_seg_n = 1024;
    ga.c_bf16 = dx; ga.ldc = 1024;
    MMF_TRY((launch_gemm2<0, 1, EPI_STORE, 512>(tA, tB, ga, 1, st)));
  }
  return MMF_OK;
}

int mmf_amil_bwd(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                 int flags, uint64_t seed, const float* A_raw, const float* ml, const float* M,
                 const float* dM, const float* dA_raw, const void* H_stash, const MmfAmilGrads* g,
                 void* dx, void* workspace, size_t workspace_bytes, void* stream) {
  (void)H_stash;  // superseded by MMF_STASHED (the stash lives in the workspace)
  if (!g || !g->dW1 || !g->db1 || !g->dWab || !g->dbab || !g->dwc || !g->dbc) return MMF_E_INVALID;
  if (flags & MMF_STASHED) {
    MMF_TRY(check_amil_common(x, N, ldx, w, L, D));
    MMF_TRY(mmf_amil_bwd_gate_hidden_stashed(N, w, L, D, flags, seed, A_raw, ml, M, dM, dA_raw, g, workspace,
                                             workspace_bytes, stream));
  } else {
    MMF_TRY(mmf_amil_bwd_gate(x, N, ldx, w, L, D, flags, seed, A_raw, ml, M, dM, dA_raw, g, workspace,
                              workspace_bytes, stream));
    MMF_TRY(mmf_amil_bwd_hidden(x, N, ldx, w, L, D, flags, A_raw, ml, dM, g, workspace, workspace_bytes, stream));
  }
  return mmf_amil_bwd_wgrad(x, N, ldx, w, L, D, flags, g, dx, workspace, workspace_bytes, stream);
}

// ---- varlen-packed TRAINING of a window of bags (gradient accumulation over `gc` small bags in one launch set) ----
int mmf_amil_window_fwd_train(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                              uint64_t seed, const int32_t* tile_valid, float* A_raw, float* partials, void* workspace,
                              size_t workspace_bytes, float* zero_buf, int64_t zero_count, void* stream) {
  if (!tile_valid) return MMF_E_INVALID;
  return fwd_train_impl(x, R, ldx, w, L, D, flags, seed, A_raw, partials, workspace, workspace_bytes, zero_buf, zero_count,
                        nullptr, stream, tile_valid);
}

int mmf_amil_window_head_nll_step(const float* partials, const int32_t* seg_tile_offsets, int n_bags, int max_tiles,
                                  int L, const float* Wk, const float* bk, int K, const int64_t* Y, const float* c,
                                  float alpha, float eps, float loss_scale, float* M, float* ml, float* hazards, float* S,
                                  int64_t* Y_hat, float* loss, float* dM, float* dWk, float* dbk, void* stream) {
  if (!partials || !seg_tile_offsets || !Wk || !bk || !Y || !c || !M || !ml || !hazards || !S || !loss || !dM)
    return MMF_E_INVALID;
  if (n_bags <= 0 || n_bags > 65535 || max_tiles <= 0 || max_tiles > 4096 || K <= 0 || K > 16) return MMF_E_UNSUPPORTED;
  const int cp = L / (2 * HEAD_CLUSTER);
  if (!(L % (2 * HEAD_CLUSTER) == 0 && cp >= 16 && cp <= 64 && 512 % cp == 0)) return MMF_E_UNSUPPORTED;   // L in {256, 512, 1024}
  return launch_pdl(amil_head_step_cluster_kernel, dim3(HEAD_CLUSTER, n_bags), dim3(512), 0, (cudaStream_t)stream,
                    partials, 0, L, Wk, bk, K, reinterpret_cast<const long long*>(Y), c, alpha, eps, M, ml, hazards, S,
                    reinterpret_cast<long long*>(Y_hat), loss, dM, dWk, dbk, reinterpret_cast<const int*>(seg_tile_offsets),
                    loss_scale);
}

int mmf_amil_window_bwd(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                        uint64_t seed, const float* A_raw, const float* ml, const float* M, const float* dM,
                        const int32_t* tile_bag, const int32_t* tile_valid, const MmfAmilGrads* g, void* dx,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !g->dW1 || !g->db1 || !g->dWab || !g->dbab || !g->dwc || !g->dbc || !tile_bag || !tile_valid) return MMF_E_INVALID;
  if (!(flags & MMF_STASHED)) return MMF_E_UNSUPPORTED;
  MMF_TRY(check_amil_common(x, R, ldx, w, L, D));
  MMF_TRY(bwd_gate_hidden_stashed_impl(R, w, L, D, flags, seed, A_raw, ml, M, dM, nullptr, g, workspace, workspace_bytes,
                                       nullptr, nullptr, stream, tile_bag, tile_valid));
  return mmf_amil_bwd_wgrad(x, R, ldx, w, L, D, flags, g, dx, workspace, workspace_bytes, stream);
}

int mmf_linear_bf16(const void* const* A_segs, int n_segs, int64_t M, int K_per_seg, int64_t lda,
                    const void* W, const float* bias, int N, void* out_bf16, float* out_f32,
                    int64_t ldc, void* stream) {
  if (!A_segs || n_segs < 1 || n_segs > 4 || M <= 0 || !W || N <= 0) return MMF_E_INVALID;
  if ((out_bf16 == nullptr) == (out_f32 == nullptr)) return MMF_E_INVALID;
  if (K_per_seg % 64 || N % 32) return MMF_E_UNSUPPORTED;
  TMapSet tA = {}, tB = {};
  for (int i = 0; i < n_segs; ++i)
    MMF_TRY(make_tmap_bf16(&tA.m[i], A_segs[i], (uint64_t)M, K_per_seg, (uint64_t)lda, 128));
  const int64_t Ktot = (int64_t)n_segs * K_per_seg;
  MMF_TRY(make_tmap_bf16(&tB.m[0], W, N, Ktot, Ktot, 256));
  GemmArgs ga = {};
  ga.M = (int)M; ga.N = N; ga.kb_total = (int)(Ktot / 64); ga.kb_per_split = ga.kb_total;
  ga.a_seg_kb = K_per_seg / 64; ga.b_seg_n = N;
  ga.c_bf16 = out_bf16; ga.c_f32 = out_f32; ga.ldc = ldc; ga.bias = bias;
  return launch_gemm<0, 0, EPI_STORE>(tA, tB, ga, 1, (cudaStream_t)stream);
}

size_t mmf_linear_bf16_wgrad_workspace_bytes(int64_t M, int N) { (void)M; (void)N; return 0; }

int mmf_linear_bf16_wgrad(const void* dY, int64_t M, int N, int64_t lddy, const void* const* X_segs,
                          int n_segs, int K_per_seg, int64_t ldx, float* dW, float* db,
                          void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!dY || !X_segs || n_segs < 1 || n_segs > 4 || M <= 0 || !dW) return MMF_E_INVALID;
  if (N % 128 || K_per_seg % 256) return MMF_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  TMapSet tA = {}, tB = {};
  MMF_TRY(make_tmap_bf16(&tA.m[0], dY, (uint64_t)M, N, (uint64_t)lddy, 64));
  for (int i = 0; i < n_segs; ++i)
    MMF_TRY(make_tmap_bf16(&tB.m[i], X_segs[i], (uint64_t)M, K_per_seg, (uint64_t)ldx, 64));
  const int Ktot = n_segs * K_per_seg;
  GemmArgs ga = {};
  ga.M = N; ga.N = Ktot; ga.kb_total = (int)((M + 63) / 64);
  ga.a_seg_kb = ga.kb_total; ga.b_seg_n = K_per_seg;
  ga.c_f32 = dW; ga.ldc = Ktot;
  const int splits = pick_splits_pair(((N + 255) / 256) * (Ktot / 256), ga.kb_total, &ga.kb_per_split);
  MMF_TRY((launch_gemm2<1, 1, EPI_ATOMIC, 256>(tA, tB, ga, splits, st)));
  if (db) {
    colsum_bf16_kernel<<<(N + 31) / 32, dim3(32, 8), 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dY), M, N, lddy, db, 1);
    MMF_TRY(launch_status());
  }
  return MMF_OK;
}

// ------------------------------ small fp32 kernels --------------------------------------------

int mmf_dense_fwd(const float* x, int64_t ldx, const float* W, const float* b, int B, int in_dim,
                  int out_dim, int act, float* y, int64_t ldy, void* stream) {
  if (!x || !W || !y || B <= 0 || in_dim <= 0 || out_dim <= 0) return MMF_E_INVALID;
  launch_sgemm(B, out_dim, in_dim, LoadRowMajor{x, ldx}, LoadRowMajor{W, in_dim}, EpiBiasAct{y, ldy, b, act},
               (cudaStream_t)stream);
  return launch_status();
}

size_t mmf_dense_fwd_workspace_bytes(int B, int in_dim, int out_dim) {
  if (B <= 0 || in_dim <= 0 || out_dim <= 0) return 0;
  const int splits = dense_fwd_splits(B, in_dim, out_dim);
  return splits > 1 ? sizeof(float) * (size_t)splits * (size_t)B * (size_t)out_dim : 0;
}

int mmf_dense_fwd_ws(const float* x, int64_t ldx, const float* W, const float* b, int B, int in_dim, int out_dim,
                     int act, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !W || !y || B <= 0 || in_dim <= 0 || out_dim <= 0) return MMF_E_INVALID;
  const int splits = dense_fwd_splits(B, in_dim, out_dim);
  if (splits <= 1) return mmf_dense_fwd(x, ldx, W, b, B, in_dim, out_dim, act, y, ldy, stream);
  if (!workspace || workspace_bytes < mmf_dense_fwd_workspace_bytes(B, in_dim, out_dim)) return MMF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = static_cast<float*>(workspace);
  launch_sgemm_slices(B, out_dim, in_dim, LoadRowMajor{x, ldx}, LoadRowMajor{W, in_dim}, ws, splits, st);
  long long blocks = ((long long)B * out_dim + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  splitk_fixup_kernel<<<(int)blocks, 256, 0, st>>>(ws, splits, B, out_dim, b, act, y, ldy);
  return launch_status();
}

int mmf_dense_bwd(const float* x, int64_t ldx, const float* W, int B, int in_dim, int out_dim,
                  int act, const float* y, int64_t ldy, const float* dy, int64_t lddy, float* dx,
                  int64_t lddx, int accumulate_dx, float* dW, float* db, void* stream) {
  if (!x || !W || !y || !dy || B <= 0) return MMF_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx)  // dx[b,i] = sum_o dpre[b,o] W[o,i]
    launch_sgemm(B, in_dim, out_dim, LoadDpre{dy, lddy, y, ldy, act}, LoadColMajor{W, in_dim},
                 EpiStoreAcc{dx, lddx, accumulate_dx}, st);
  if (dW)  // dW[o,i] += sum_b dpre[b,o] x[b,i]
    launch_sgemm(out_dim, in_dim, B, LoadDpreT{dy, lddy, y, ldy, act}, LoadColMajor{x, ldx},
                 EpiStoreAcc{dW, in_dim, 1}, st);
  if (db)
    colsum_functor_kernel<<<(out_dim + 31) / 32, 256, 0, st>>>(B, out_dim, LoadDpre{dy, lddy, y, ldy, act}, db, 1);
  return launch_status();
}

int mmf_kron_enc_fwd(const float* const* o, int m, int E, int B, const float* W, const float* b,
                     int H, float* out, void* stream) {
  return mmf_kron_enc_train_fwd(o, m, E, B, W, b, H, 0, 0, out, stream);
}

namespace {
inline KronElem kron_elem(const float* const* o, int m, int E, uint64_t seed, int dropout) {
  return KronElem{o[0], o[1], m >= 3 ? o[2] : nullptr, m >= 4 ? o[3] : nullptr, m, E, seed, dropout,
                  dropout > 1 ? 65536.0f / (65536.0f - (float)dropout) : 1.0f};
}
}  // namespace

int mmf_kron_enc_train_fwd(const float* const* o, int m, int E, int B, const float* W, const float* b,
                           int H, int dropout, uint64_t seed, float* out, void* stream) {
  return mmf_kron_enc_train_fwd_ws(o, m, E, B, W, b, H, dropout, seed, out, nullptr, 0, stream);
}

size_t mmf_kron_enc_fwd_workspace_bytes(int m, int E, int B, int H) {
  long long kk = 1;
  for (int t = 0; t < m; ++t) kk *= E;
  if (B <= 0 || H <= 0 || kk <= 0 || kk > (1ll << 30)) return 0;
  const int splits = dense_fwd_splits(B, (int)kk, H);
  return splits > 1 ? sizeof(float) * (size_t)splits * (size_t)B * (size_t)H : 0;
}

int mmf_kron_enc_train_fwd_ws(const float* const* o, int m, int E, int B, const float* W, const float* b, int H,
                              int dropout, uint64_t seed, float* out, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (!o || m < 2 || m > 4 || E <= 0 || B <= 0 || !W || !out || dropout < 0 || dropout > 65535) return MMF_E_INVALID;
  KronElem e = kron_elem(o, m, E, seed, dropout);
  if (e.width() > (1ll << 30)) return MMF_E_INVALID;
  const int KK = (int)e.width();
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = dense_fwd_splits(B, KK, H);
  if (splits > 1 && workspace && workspace_bytes >= mmf_kron_enc_fwd_workspace_bytes(m, E, B, H)) {
    // few output tiles, K = E^m (4913 for three modalities): deterministic split-K as mmf_dense_fwd_ws
    float* ws = static_cast<float*>(workspace);
    launch_sgemm_slices(B, H, KK, LoadKronA{e}, LoadRowMajor{W, KK}, ws, splits, st);
    long long blocks = ((long long)B * H + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_fixup_kernel<<<(int)blocks, 256, 0, st>>>(ws, splits, B, H, b, MMF_ACT_RELU, out, H);
    return launch_status();
  }
  launch_sgemm(B, H, KK, LoadKronA{e}, LoadRowMajor{W, KK}, EpiBiasAct{out, H, b, MMF_ACT_RELU}, st);
  return launch_status();
}

size_t mmf_kron_enc_workspace_bytes(int m, int E, int B) {
  size_t kk = 1;
  for (int t = 0; t < m; ++t) kk *= (size_t)E;
  return kk * (size_t)B * sizeof(float);
}

int mmf_kron_enc_bwd(const float* const* o, int m, int E, int B, const float* W, int H,
                     const float* out, const float* dout, float* const* d_o, float* dW, float* db,
                     void* workspace, size_t workspace_bytes, void* stream) {
  return mmf_kron_enc_train_bwd(o, m, E, B, W, H, 0, 0, out, dout, d_o, dW, db, workspace, workspace_bytes, stream);
}

int mmf_kron_enc_train_bwd(const float* const* o, int m, int E, int B, const float* W, int H, int dropout,
                           uint64_t seed, const float* out, const float* dout, float* const* d_o, float* dW,
                           float* db, void* workspace, size_t workspace_bytes, void* stream) {
  if (!o || m < 2 || m > 4 || !W || !out || !dout || !workspace || dropout < 0 || dropout > 65535) return MMF_E_INVALID;
  if (workspace_bytes < mmf_kron_enc_workspace_bytes(m, E, B)) return MMF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  KronElem e = kron_elem(o, m, E, seed, dropout);
  if (e.width() > (1ll << 30)) return MMF_E_INVALID;
  const int KK = (int)e.width();
  float* dkron = reinterpret_cast<float*>(workspace);
  LoadDpre dpre{dout, H, out, H, MMF_ACT_RELU};
  LoadDpreT dpreT{dout, H, out, H, MMF_ACT_RELU};
  if (d_o) {
    // dkron[b,kk] = sum_h dpre[b,h] W[h,kk]
    launch_sgemm(B, KK, H, dpre, LoadColMajor{W, KK}, EpiStoreAcc{dkron, KK, 0}, st);
    kron_contract_kernel<<<B, 256, 8 * E * sizeof(float), st>>>(dkron, e, B, d_o[0], d_o[1],
                                                                m >= 3 ? d_o[2] : nullptr,
                                                                m >= 4 ? d_o[3] : nullptr);
  }
  if (dW)  // dW[h,kk] += sum_b dpre[b,h] kron[b,kk]
    launch_sgemm(H, KK, B, dpreT, LoadKronB{e}, EpiStoreAcc{dW, KK, 1}, st);
  if (db) colsum_functor_kernel<<<(H + 31) / 32, 256, 0, st>>>(B, H, dpre, db, 1);
  return launch_status();
}

int mmf_hazard_head_fwd(const float* M, int B, int Lin, const float* Wk, const float* bk, int K,
                        float* hazards, float* S, int64_t* Y_hat, void* stream) {
  if (!M || !Wk || !bk || !hazards || !S || B <= 0 || K <= 0 || K > 16) return MMF_E_INVALID;
  const int warps_per_block = 4;
  hazard_head_fwd_kernel<<<(B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                           (cudaStream_t)stream>>>(M, B, Lin, Wk, bk, K, hazards, S,
                                                   reinterpret_cast<long long*>(Y_hat));
  return launch_status();
}

int mmf_hazard_head_bwd(const float* M, int B, int Lin, const float* Wk, int K, const float* hazards,
                        const float* S, const float* d_hazards, const float* d_S, float* dM,
                        float* dWk, float* dbk, void* stream) {
  (void)S;
  if (!M || !Wk || !hazards || B <= 0 || K <= 0 || K > 16) return MMF_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  HazardDlogit d{hazards, d_hazards, d_S, K};
  if (dM)  // dM[b,l] = sum_j dlogit[b,j] Wk[j,l]
    launch_sgemm(B, Lin, K, LoadHazA{d}, LoadColMajor{Wk, Lin}, EpiStoreAcc{dM, Lin, 0}, st);
  if (dWk)  // dWk[j,l] += sum_b dlogit[b,j] M[b,l]
    launch_sgemm(K, Lin, B, LoadHazAT{d}, LoadColMajor{M, Lin}, EpiStoreAcc{dWk, Lin, 1}, st);
  if (dbk) colsum_functor_kernel<<<(K + 31) / 32, 256, 0, st>>>(B, K, LoadHazA{d}, dbk, 1);
  return launch_status();
}

int mmf_amil_head_nll_step(const float* partials, int64_t n, int L, const float* Wk, const float* bk, int K,
                           const int64_t* Y, const float* c, float alpha, float eps, float* M, float* ml,
                           float* hazards, float* S, int64_t* Y_hat, float* loss, float* dM, float* dWk,
                           float* dbk, void* stream) {
  if (!partials || !Wk || !bk || !Y || !c || !M || !ml || !hazards || !S || !loss || !dM) return MMF_E_INVALID;
  if (n <= 0 || n > 4096 || L <= 0 || L > 1024 || (L & 1) || K <= 0 || K > 16) return MMF_E_UNSUPPORTED;
  // cluster kernel: L splits into 8 CTAs x (column pairs that divide 512 threads), i.e. L in {256, 512, 1024}
  const int cp = L / (2 * HEAD_CLUSTER);
  if (L % (2 * HEAD_CLUSTER) == 0 && cp >= 16 && cp <= 64 && 512 % cp == 0)
    return launch_pdl(amil_head_step_cluster_kernel, dim3(HEAD_CLUSTER), dim3(512), 0, (cudaStream_t)stream,
                      partials, (int)n, L, Wk, bk, K, reinterpret_cast<const long long*>(Y), c, alpha, eps, M, ml,
                      hazards, S, reinterpret_cast<long long*>(Y_hat), loss, dM, dWk, dbk, (const int*)nullptr, 1.0f);
  return launch_pdl(amil_head_step_kernel, dim3(1), dim3(512), 0, (cudaStream_t)stream, partials, (int)n, L, Wk, bk, K,
                    reinterpret_cast<const long long*>(Y), c, alpha, eps, M, ml, hazards, S,
                    reinterpret_cast<long long*>(Y_hat), loss, dM, dWk, dbk);
}

int mmf_nll_surv_fwd_bwd(const float* hazards, const float* S, const int64_t* Y, const float* c,
                         int B, int K, float alpha, float eps, float* loss, float* d_hazards,
                         float* d_S, void* stream) {
  if (!hazards || !S || !Y || !c || !loss || B <= 0 || K <= 0) return MMF_E_INVALID;
  nll_surv_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(hazards, S, reinterpret_cast<const long long*>(Y), c, B, K,
                                                       alpha, eps, loss, d_hazards, d_S);
  return launch_status();
}

int mmf_ce_surv_fwd_bwd(const float* hazards, const float* S, const int64_t* Y, const float* c, int B, int K,
                        float alpha, float eps, float* loss, float* d_hazards, float* d_S, void* stream) {
  if (!hazards || !S || !Y || !c || !loss || B <= 0 || K <= 0) return MMF_E_INVALID;
  ce_surv_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(hazards, S, reinterpret_cast<const long long*>(Y), c, B, K, alpha,
                                                      eps, loss, d_hazards, d_S);
  return launch_status();
}

int mmf_batchnorm1d_fwd(const float* x, int B, int F, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, int train, float momentum, float eps, float* y, float* save_mean,
                        float* save_invstd, void* stream) {
  if (!x || !y || B <= 0 || F <= 0 || !save_mean || !save_invstd) return MMF_E_INVALID;
  if (train && B < 2) return MMF_E_INVALID;   // nn.BatchNorm1d: "Expected more than 1 value per channel when training"
  if (!train && (!running_mean || !running_var)) return MMF_E_INVALID;
  batchnorm1d_fwd_kernel<<<(F + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, B, F, gamma, beta, running_mean,
                                                                                   running_var, train, momentum, eps, y,
                                                                                   save_mean, save_invstd);
  return launch_status();
}

int mmf_batchnorm1d_bwd(const float* x, const float* dy, int B, int F, const float* gamma, const float* save_mean,
                        const float* save_invstd, int train, float* dx, float* dgamma, float* dbeta, void* stream) {
  if (!x || !dy || !save_mean || !save_invstd || B <= 0 || F <= 0) return MMF_E_INVALID;
  batchnorm1d_bwd_kernel<<<(F + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, dy, B, F, gamma, save_mean,
                                                                                   save_invstd, train, dx, dgamma, dbeta);
  return launch_status();
}

int mmf_highway_mix_fwd(const float* gate, const float* nonlinear, const float* linear, int64_t count, float* y,
                        void* stream) {
  if (!gate || !nonlinear || !linear || !y || count <= 0) return MMF_E_INVALID;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  highway_mix_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gate, nonlinear, linear, count, y);
  return launch_status();
}

int mmf_highway_mix_bwd(const float* gate, const float* nonlinear, const float* linear, const float* dy, int64_t count,
                        float* dgate, float* dnonlinear, float* dlinear, void* stream) {
  if (!gate || !nonlinear || !linear || !dy || !dgate || !dnonlinear || !dlinear || count <= 0) return MMF_E_INVALID;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  highway_mix_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gate, nonlinear, linear, dy, count, dgate,
                                                                         dnonlinear, dlinear);
  return launch_status();
}

size_t mmf_cox_workspace_bytes(int B) { (void)B; return 0; }

int mmf_cox_fwd_bwd(const float* theta, const float* times, const float* c, int B, float* loss,
                    float* dtheta, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!theta || !times || !c || !loss || B <= 0) return MMF_E_INVALID;
  if (B > MMF_COX_MAXB) return MMF_E_UNSUPPORTED;
  cox_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(theta, times, c, B, loss, dtheta);
  return launch_status();
}

size_t mmf_ranking_workspace_bytes(int B) { return 16 + (size_t)B * sizeof(float); }

int mmf_ranking_fwd_bwd(const float* risks, const float* times, const float* c, int B, int phi,
                        int reduction, float* loss, float* drisks, int64_t* n_pairs, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!risks || !times || !c || !loss || !workspace || B < 2) return MMF_E_INVALID;
  if (workspace_bytes < mmf_ranking_workspace_bytes(B)) return MMF_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return MMF_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = reinterpret_cast<double*>(workspace);
  float* g = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 16);
  MMF_TRY(cuda_rc(cudaMemsetAsync(acc, 0, 16, st)));
  const int blocks = (B + 127) / 128;
  ranking_pairs_kernel<<<blocks, 128, 0, st>>>(risks, times, c, B, phi, acc, g);
  ranking_finalize_kernel<<<blocks, 128, 0, st>>>(acc, B, reduction, loss, drisks, g,
                                                  reinterpret_cast<long long*>(n_pairs));
  return launch_status();
}

namespace {
int adam_step_multi_impl(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                         float* const* exp_avg_sq_host, const int64_t* numel_host, int n_tensors, int step,
                         const uint64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                         float grad_scale, float l1_lambda, int zero_grad, float* l1_out, void* stream) {
  if (!params_host || !grads_host || !exp_avg_host || !exp_avg_sq_host || !numel_host || n_tensors <= 0 ||
      (step_dev == nullptr && step < 1))
    return MMF_E_INVALID;
  AdamHyper h = {};
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay; h.grad_scale = grad_scale;
  h.l1_lambda = l1_lambda; h.zero_grad = zero_grad;
  h.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
  if (step_dev == nullptr) {
    h.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    h.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  }
  for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
    AdamTensors T = {};
    T.n = n_tensors - t0 < ADAM_MAX_TENSORS ? n_tensors - t0 : ADAM_MAX_TENSORS;
    long long acc = 0;
    for (int k = 0; k < T.n; ++k) {
      if (!params_host[t0 + k] || !grads_host[t0 + k] || !exp_avg_host[t0 + k] || !exp_avg_sq_host[t0 + k] ||
          numel_host[t0 + k] < 0)
        return MMF_E_INVALID;
      T.p[k] = params_host[t0 + k]; T.g[k] = grads_host[t0 + k]; T.m[k] = exp_avg_host[t0 + k];
      T.v[k] = exp_avg_sq_host[t0 + k];
      T.start[k] = acc; acc += numel_host[t0 + k];
    }
    T.start[T.n] = acc;
    if (acc == 0) continue;
    long long blocks = (acc + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_multi_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(T, h, l1_out);
    MMF_TRY(launch_status());
  }
  return MMF_OK;
}
}  // namespace

int mmf_adam_step_multi(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                        float* const* exp_avg_sq_host, const int64_t* numel_host, int n_tensors, int step, float lr,
                        float beta1, float beta2, float eps, float weight_decay, float grad_scale, float l1_lambda,
                        int zero_grad, float* l1_out, void* stream) {
  return adam_step_multi_impl(params_host, grads_host, exp_avg_host, exp_avg_sq_host, numel_host, n_tensors, step, nullptr,
                              lr, beta1, beta2, eps, weight_decay, grad_scale, l1_lambda, zero_grad, l1_out, stream);
}

int mmf_adam_step_multi_dev(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                            float* const* exp_avg_sq_host, const int64_t* numel_host, int n_tensors,
                            const uint64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float grad_scale, float l1_lambda, int zero_grad, float* l1_out, void* stream) {
  if (!step_dev) return MMF_E_INVALID;
  return adam_step_multi_impl(params_host, grads_host, exp_avg_host, exp_avg_sq_host, numel_host, n_tensors, 0, step_dev,
                              lr, beta1, beta2, eps, weight_decay, grad_scale, l1_lambda, zero_grad, l1_out, stream);
}

int mmf_step_state_advance(uint64_t* state, int n_seeds, void* stream) {
  if (!state || n_seeds < 0 || n_seeds > 1023) return MMF_E_INVALID;
  step_state_advance_kernel<<<(n_seeds + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<unsigned long long*>(state), n_seeds);
  return launch_status();
}

namespace {
int xf_pack(const MmfXfusionMod* mods, int m, int B, int dim, XfMods* P) {
  if (!mods || m < 2 || m > XF_MAX_MOD || B <= 0 || dim <= 0 || dim % XF_KC != 0) return MMF_E_INVALID;
  *P = XfMods{};
  P->m = m; P->B = B; P->dim = dim;
  for (int i = 0; i < m; ++i) {
    const MmfXfusionMod& q = mods[i];
    if (!q.v || !q.Wh || !q.bh || !q.Wz || !q.bz || !q.Wo || !q.bo) return MMF_E_INVALID;
    if ((reinterpret_cast<uintptr_t>(q.v) & 15u) != 0) return MMF_E_ALIGN;
    P->mod[i] = XfMod{q.v, q.Wh, q.bh, q.Wz, q.bz, q.Wo, q.bo};
  }
  return MMF_OK;
}
}  // namespace

int mmf_xfusion_gate_fwd(const MmfXfusionMod* mods_host, int m, int B, int dim, const float* mask, float* h, float* z,
                         float* o, void* stream) {
  XfMods P;
  MMF_TRY(xf_pack(mods_host, m, B, dim, &P));
  if (!h || !z || !o) return MMF_E_INVALID;
  xfusion_gate_fwd_kernel<<<dim3((B + XF_RB - 1) / XF_RB, m), 256, 0, (cudaStream_t)stream>>>(P, mask, h, z, o);
  return launch_status();
}

size_t mmf_xfusion_gate_bwd_workspace_bytes(int m, int B, int dim) {
  if (m < 2 || m > XF_MAX_MOD || B <= 0 || dim <= 0) return 0;
  const size_t Z = (size_t)xf_slices(B);
  return sizeof(float) * ((size_t)m * B * 2 * XF_S + Z * m * XF_SMALL_OUT + Z * m * XF_S * ((size_t)dim * m + dim));
}

int mmf_xfusion_gate_bwd(const MmfXfusionMod* mods_host, int m, int B, int dim, const float* mask, const float* h,
                         const float* z, const float* o, const float* d_o, const MmfXfusionGrads* grads_host,
                         int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  XfMods P;
  MMF_TRY(xf_pack(mods_host, m, B, dim, &P));
  if (!h || !z || !o || !d_o || !grads_host) return MMF_E_INVALID;
  if (!workspace || workspace_bytes < mmf_xfusion_gate_bwd_workspace_bytes(m, B, dim)) return MMF_E_WORKSPACE;
  XfGrads G = {};
  bool any_dv = false;
  for (int i = 0; i < m; ++i) {
    const MmfXfusionGrads& g = grads_host[i];
    if (!g.dWh || !g.dbh || !g.dWz || !g.dbz || !g.dWo || !g.dbo) return MMF_E_INVALID;
    G.dWh[i] = g.dWh; G.dbh[i] = g.dbh; G.dWz[i] = g.dWz; G.dbz[i] = g.dbz; G.dWo[i] = g.dWo; G.dbo[i] = g.dbo;
    G.dv[i] = g.dv;
    any_dv = any_dv || g.dv != nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Z = xf_slices(B);
  float* dhz = static_cast<float*>(workspace);
  float* part_small = dhz + (size_t)m * B * 2 * XF_S;
  float* part_w = part_small + (size_t)Z * m * XF_SMALL_OUT;
  xfusion_gate_bwd_small_kernel<<<dim3(Z, m), 1024, 0, st>>>(P, mask, h, z, o, d_o, dhz, part_small);
  xfusion_gate_bwd_wgrad_kernel<<<dim3(dim * m / 64, m, Z), 256, 0, st>>>(P, dhz, part_w);
  const long long total = (long long)m * (XF_S * ((long long)dim * m + dim) + XF_SMALL_OUT);
  xfusion_gate_bwd_finalize_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(P, part_small, part_w, G, accumulate);
  if (any_dv) xfusion_gate_bwd_dv_kernel<<<dim3((B + XF_RB - 1) / XF_RB, m), 256, 0, st>>>(P, dhz, G);
  return launch_status();
}

namespace {
int snn_pack(const float* x, int B, int in_dim, const MmfSnnLayer* layers, int n, SnnLayers* P) {
  if (!x || !layers || B <= 0 || in_dim <= 0 || n < 1 || n > SNN_MAX_LAYERS) return MMF_E_INVALID;
  *P = SnnLayers{};
  P->n = n; P->B = B; P->d[0] = in_dim;
  for (int l = 0; l < n; ++l) {
    const MmfSnnLayer& q = layers[l];
    if (!q.W || !q.b || !q.y || q.width <= 0 || q.width > SNN_MAX_WIDTH) return MMF_E_INVALID;
    if (q.keep && !(q.p > 0.f && q.p < 1.f)) return MMF_E_INVALID;
    P->d[l + 1] = q.width; P->W[l] = q.W; P->b[l] = q.b; P->keep[l] = q.keep; P->y[l] = q.y;
    const double al = (double)MMF_ALPHA_PRIME, pp = q.keep ? (double)q.p : 0.0;
    const double a = 1.0 / sqrt((1.0 - pp) * (1.0 + pp * al * al));
    P->da[l] = (float)a; P->db[l] = (float)(-a * al * pp);
  }
  return MMF_OK;
}
}  // namespace

int mmf_snn_mlp_fwd(const float* x, int B, int in_dim, const MmfSnnLayer* layers_host, int n_layers, float* out,
                    void* stream) {
  SnnLayers P;
  MMF_TRY(snn_pack(x, B, in_dim, layers_host, n_layers, &P));
  if (!out) return MMF_E_INVALID;
  snn_mlp_fwd_kernel<<<(B + SNN_RB - 1) / SNN_RB, 256, 0, (cudaStream_t)stream>>>(P, x, out);
  return launch_status();
}

int mmf_snn_mlp_bwd(const float* x, int B, int in_dim, const MmfSnnLayer* layers_host, int n_layers, const float* dout,
                    float* const* dW_host, float* const* db_host, int accumulate, float* dx, void* workspace,
                    size_t workspace_bytes, void* stream) {
  SnnLayers P;
  MMF_TRY(snn_pack(x, B, in_dim, layers_host, n_layers, &P));
  if (!dout || !dW_host || !db_host) return MMF_E_INVALID;
  size_t need = 0;
  for (int l = 0; l < n_layers; ++l) need += (size_t)B * (size_t)P.d[l + 1] * sizeof(float);
  if (!workspace || workspace_bytes < need) return MMF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  SnnBwd G = {};
  float* w = static_cast<float*>(workspace);
  for (int l = 0; l < n_layers; ++l) { G.dpre[l] = w; w += (size_t)B * P.d[l + 1]; }
  G.dx = dx;
  snn_mlp_bwd_chain_kernel<<<(B + SNN_RB - 1) / SNN_RB, 256, 0, st>>>(P, dout, G);
  for (int l = 0; l < n_layers; ++l) {
    const int d_in = P.d[l], d_out = P.d[l + 1];
    if (dW_host[l]) {
      LoadPlainT A{G.dpre[l], d_out};
      EpiStoreAcc E{dW_host[l], d_in, accumulate};
      if (l == 0) launch_sgemm(d_out, d_in, B, A, LoadColMajor{x, d_in}, E, st);
      else launch_sgemm(d_out, d_in, B, A, LoadAlphaDropCol{P.y[l - 1], P.keep[l - 1], d_in, P.da[l - 1], P.db[l - 1]}, E, st);
    }
    if (db_host[l])
      colsum_functor_kernel<<<(d_out + 31) / 32, 256, 0, st>>>(B, d_out, ElemPlain{G.dpre[l], d_out}, db_host[l], accumulate);
  }
  return launch_status();
}

int mmf_cindex_counts(const float* risk, const float* times, const float* event, int B, float tied_tol,
                      uint64_t* counts, void* stream) {
  if (!risk || !times || !event || !counts || B <= 0) return MMF_E_INVALID;
  MMF_TRY(cuda_rc(cudaMemsetAsync(counts, 0, 3 * sizeof(uint64_t), (cudaStream_t)stream)));
  cindex_pairs_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(risk, times, event, B, tied_tol,
                                                           reinterpret_cast<unsigned long long*>(counts));
  return launch_status();
}

int mmf_percentile_of_score(const float* ref_scores, int n_ref, const float* query, int n_query, float* out,
                            void* stream) {
  if (!ref_scores || !query || !out || n_ref <= 0 || n_query <= 0) return MMF_E_INVALID;
  if ((long long)n_ref * n_query > (1ll << 36)) return MMF_E_UNSUPPORTED;   // brute-force pair count: keep it under ~1 s
  percentile_of_score_kernel<<<(n_query + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ref_scores, n_ref, query, n_query, out);
  return launch_status();
}

size_t mmf_p2p_flag_bytes(void) { return (size_t)P2P_FLAG_WORDS * sizeof(uint32_t); }

int mmf_p2p_allreduce_sum_f32(void* const* bufs_host, void* const* flags_host, void* multicast_ptr, int world,
                              int rank, int64_t n, int n_ctas, void* stream) {
  if (!bufs_host || !flags_host || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return MMF_E_INVALID;
  if (n <= 0 || (n & 3)) return MMF_E_INVALID;
  if (n_ctas <= 0) n_ctas = 32;
  if (n_ctas > P2P_MAX_CTAS) n_ctas = P2P_MAX_CTAS;
  P2pArgs a = {};
  for (int p = 0; p < world; ++p) {
    if (!bufs_host[p] || !flags_host[p] || (reinterpret_cast<uintptr_t>(bufs_host[p]) & 15u)) return MMF_E_ALIGN;
    a.buf[p] = reinterpret_cast<float*>(bufs_host[p]);
    a.flags[p] = reinterpret_cast<uint32_t*>(flags_host[p]);
  }
  a.n = n; a.world = world; a.rank = rank;
  a.stamps = g_p2p_stamps;
  a.mc = reinterpret_cast<float*>(multicast_ptr);
  if (multicast_ptr && (reinterpret_cast<uintptr_t>(multicast_ptr) & 15u)) return MMF_E_ALIGN;
  // Plain stream-ordered launch (as a programmatic dependent of the wgrad kernel the exchange took 100 us instead of
  // 35 us), as clusters of 2 CTAs: the exchange usually overlaps the next step on a second stream, and its CTAs sit
  // in handshake spins until the slowest rank arrives. Scattered single CTAs would each take one SM out of a
  // different TPC and leave the step's CTA-PAIR kernels (70 pairs of 74 TPCs) short of whole TPCs — a second wave
  // of the wgrad kernel, +80 us per step at 8 GPUs. Pairs of CTAs occupy whole TPCs instead.
  if (n_ctas & 1) ++n_ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas); cfg.blockDim = dim3(P2P_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  // MMF_P2P_PDL=1: also a programmatic dependent of the previous kernel in the stream (in-stream placement right
  // after the wgrad GEMM: with only 2-8 TPCs taken by the exchange the wgrad's 70 CTA pairs still fit in one wave)
  static const bool p2p_pdl = getenv("MMF_P2P_PDL") && getenv("MMF_P2P_PDL")[0] == '1';
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = p2p_pdl ? 2 : 1;
  MMF_TRY(cuda_rc(cudaLaunchKernelEx(&cfg, p2p_allreduce_sum_kernel, a)));
  return launch_status();
}

}  // extern "C"
