/* mmf_b200.h — C ABI of the B200-native attention-MIL / fusion / survival-head hot path.
 *
 * The reference (MultimodalFusion/multimodalfusion) is pure PyTorch: it has no FFI of its own,
 * its boundary is the nn.Module surface.  Each entry point below therefore replaces a *stock
 * ATen op sequence* of the reference; the citation on every function names that sequence
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding and the
 * torch.autograd.Function a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - the caller owns all buffers (including workspaces); the library never allocates device
 *     memory, keeps no global state and is re-entrant;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation;
 *   - return value: MMF_OK (0) or a negative MMF_E_* code; values <= -1000 are -(1000+cudaError_t);
 *   - bf16 arrays are row-major, 16-byte aligned, leading dimension a multiple of 8 elements;
 *   - requires an sm_100a device (B200): tcgen05/TMEM/TMA kernels, no fallback path.
 */
#ifndef MMF_B200_H_
#define MMF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMF_ABI_VERSION 1

enum {
  MMF_OK = 0,
  MMF_E_INVALID = -1,     /* bad argument (null pointer, size, unsupported L/D) */
  MMF_E_ALIGN = -2,       /* pointer / leading dimension alignment */
  MMF_E_DRIVER = -3,      /* cuTensorMapEncodeTiled entry point unavailable */
  MMF_E_TMAP = -4,        /* tensor-map encode failed */
  MMF_E_UNSUPPORTED = -5, /* feature size not compiled in */
  MMF_E_WORKSPACE = -6    /* workspace too small */
};

/* flags for the attention-MIL kernels */
enum {
  MMF_GATED = 1,        /* Attn_Net_Gated (tanh ⊙ sigmoid) vs Attn_Net (tanh only) */
  MMF_DROPOUT_H = 2,    /* train mode: Dropout(0.25) on h = relu(fc(x)) (always on in the reference) */
  MMF_DROPOUT_ATTN = 4, /* train mode with dropout=True: Dropout(0.25) on the tanh / sigmoid outputs */
  MMF_NEED_DX = 8,      /* backward also produces dx (radio path: reduce_dim sits upstream) */
  MMF_STASHED = 16,     /* mmf_amil_bwd*: the forward was mmf_amil_fwd_train on the same workspace — h and the
                           branch activations are read from it instead of being recomputed */
  MMF_PRECISE_FC = 32   /* split-precision fc for small bags: x is bf16 [N, 3072] = [hi | lo | hi] of an fp32 bag
                           (mmf_split_f32_bf16x3, ldx >= 3072) and w->W1_split = [W1_hi | W1_hi | W1_lo]; GEMM1 then
                           computes x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32 accumulators (~16 mantissa bits): the
                           ReLU pattern of a tiny bag matches the fp32 reference instead of flipping units whose
                           pre-activation sits within bf16 rounding of zero. 3x the GEMM1 work: for bags of a few
                           thousand instances (radiology slices), where the step is latency-bound anyway. The
                           backward's dW1 uses x_hi (the first 1024 columns). */
};

/* Dropout seeds: every `uint64_t seed` argument of this header is either a host value below 2^63 or
 * MMF_SEED_DEVICE(ptr): the device address of a uint64 that holds the seed. The kernels then read the seed when they RUN,
 * not when they are launched — a launch captured in a CUDA graph draws a new mask on every replay once the word is moved
 * on by mmf_step_state_advance() at the head of the graph. (The reference draws its masks from torch's generator per call,
 * models/model_attention_mil_path.py:22-26; per-step seeds are the equivalent here.) */
#define MMF_SEED_DEVICE(ptr) ((uint64_t)(uintptr_t)(ptr) | 0x8000000000000000ull)
#define MMF_IN_FEATURES 1024 /* ResNet50-layer3 feature width, fixed by the reference models */
#define MMF_TILE_ROWS 128    /* instances per CTA tile; one (m, l, acc[L]) partial per tile */

int mmf_version(void);
/* DEBUG BUILDS ONLY: a no-op in the release library (no global state). In a library compiled with -DMMF_DEBUG_STAMPS=1
 * (tools/phase_*.py, tools/dp_diag.py): device buffer of gridDim.x*16 uint64 that the tensor-core kernels fill with clock64()
 * phase stamps; NULL disables it. mmf_debug_stamps_enabled() tells which build is loaded. */
void mmf_debug_set_timing_buffer(void* device_u64_buffer);
int mmf_debug_stamps_enabled(void);
/* DEBUG BUILDS ONLY: a no-op in the release library (no device-global state). In a library compiled with
 * -DMMF_DEBUG_TIMELINE=1 (tools/step_timeline.py): device buffer of 16 + 5 * 32768 uint64, zero-filled; every CTA of every
 * hot-path kernel appends one record (kernel id, blockIdx.x, %globaltimer ns at CTA start, after griddepcontrol.wait, at
 * CTA end) at [16 + 5 i ...], record count in [0]. ids: 0 fused forward, 1 head step, 2 fused head + gate + hidden
 * backward, 3 wgrad GEMM, 4 recompute gate, 5 other pair GEMMs. NULL disables. */
void mmf_debug_set_timeline_buffer(void* device_u64_buffer);
/* DEBUG BUILDS ONLY (-DMMF_DEBUG_STAMPS=1; no-op in the release library): device buffer of 8 + 4 * 4000 uint64, zero-filled; CTA 0 of every peer all-reduce appends
 * 4 %globaltimer stamps (kernel start, ready handshake done, data phase done, done handshake done); count in [0]. */
void mmf_debug_set_p2p_stamp_buffer(void* device_u64_buffer);
const char* mmf_error_string(int rc);

/* Weights of fc(1024->L) + attention net, prepared once per optimizer step by the caller.
 * Reference: models/model_attention_mil_path.py:19-29 (fc_WSI, Attn_Net_Gated / Attn_Net),
 * models/model_modules.py:70-110. */
typedef struct MmfAmilWeights {
  const void* W1;        /* bf16 [L,1024]           attention_net_*.0.weight                      */
  const float* b1;       /* f32  [L]                attention_net_*.0.bias                        */
  const void* Wab;       /* bf16 [2D,L] (gated: attention_a.0.weight stacked on attention_b.0.weight)
                                 [D,L]  (un-gated: module.0.weight)                               */
  const void* Wab_packed;/* bf16, same rows regrouped per 128-wide D chunk c:
                            gated: rows [c*256, c*256+128) = Wa[c*128..], next 128 = Wb[c*128..];
                            un-gated: identical to Wab. See mmf_pack_wab().                        */
  const float* bab;      /* f32  [2D] = ba ++ bb   (un-gated: [D])                                */
  const float* wc;       /* f32  [D]               attention_c.weight / module.{2|3}.weight       */
  const float* bc;       /* f32  [1]               attention_c.bias (device scalar)                */
  const void* W1_split;  /* bf16 [L,3072] = [W1_hi | W1_hi | W1_lo], W1_lo = bf16(W1 - W1_hi); NULL unless
                            MMF_PRECISE_FC is used                                                 */
} MmfAmilWeights;

typedef struct MmfAmilGrads {
  float* dW1;  /* f32 [L,1024] */
  float* db1;  /* f32 [L]      */
  float* dWab; /* f32 [2D,L] (un-gated [D,L]), natural row order */
  float* dbab; /* f32 [2D] / [D] */
  float* dwc;  /* f32 [D] */
  float* dbc;  /* f32 [1] */
} MmfAmilGrads;

/* fp32 -> bf16 feature / weight conversion (round-to-nearest-even), n elements. */
int mmf_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* fp32 [n_rows, 1024] (leading dimension ldx) -> bf16 [n_rows, 3072] = [hi | lo | hi], hi = bf16(x), lo = bf16(x - hi):
 * the bag format of MMF_PRECISE_FC. */
int mmf_split_f32_bf16x3(const float* x, int64_t n_rows, int64_t ldx, void* out_bf16, void* stream);

/* Regroups Wab[2D,L] into the per-chunk layout the fused kernel streams with one TMA box. */
int mmf_pack_wab(const void* Wab_bf16, void* Wab_packed_bf16, int L, int D, int gated, void* stream);

/* Number of 128-row tiles (= number of softmax partials) for a bag of N instances. */
int64_t mmf_amil_num_tiles(int64_t N);

/* Fused attention-MIL forward for one bag (or one rank's shard of a bag).
 *   h = relu(x W1^T + b1) [dropout];  s = wc·(tanh(Wa h + ba) ⊙ sigmoid(Wb h + bb)) + bc
 *   per 128-row tile t: (m_t, l_t, acc_t[L]) = (max s, Σ e^{s-m_t}, Σ e^{s-m_t} h)
 * Replaces: nn.Linear+ReLU+Dropout (models/model_attention_mil_path.py:20-21,29),
 *           Attn_Net_Gated.forward / Attn_Net.forward (models/model_modules.py:84-85,105-110),
 *           transpose+softmax+mm (models/model_attention_mil_path.py:53-56) up to the combine.
 *   x        bf16 [N, ldx>=1024]
 *   A_raw    f32 [N]            raw (pre-softmax) attention scores
 *   partials f32 [num_tiles, L+2]  row t = (m_t, l_t, acc_t[0..L))
 *   H_stash  bf16 [N, L] or NULL: when non-NULL the h tile is also written out (consumed by
 *            mmf_amil_bwd with h_stash != NULL, which then skips the fc recompute). */
int mmf_amil_fwd(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                 int flags, uint64_t seed, float* A_raw, float* partials, void* H_stash,
                 void* stream);

/* Cohort inference over ragged bags in two launches (the per-slide forward of create_heatmaps.py /
 * pre_trained_feature.py:116-162 for a whole batch of slides; the reference loops one slide at a time).
 *   x                bf16 [R, ldx]: the bags packed back to back, every bag starting on a 128-row boundary, padding
 *                    rows ZERO; R a multiple of 128
 *   tile_valid       int32 [R/128]: rows of tile t that belong to its bag (1..128; 0 for a pure padding tile)
 *   seg_tile_offsets int32 [n_bags+1]: bag b owns tiles seg_tile_offsets[b] .. seg_tile_offsets[b+1]
 *   outputs          A_raw f32 [R] (padding rows unwritten), partials f32 [R/128, L+2] (scratch), M f32 [n_bags, L],
 *                    ml f32 [n_bags, 2] or NULL, hazards / S f32 [n_bags, K], risk f32 [n_bags] = -sum_k S (or NULL),
 *                    Y_hat int64 [n_bags] or NULL.  Eval mode only (no dropout flags). */
int mmf_amil_infer_varlen(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                          const int32_t* tile_valid, const int32_t* seg_tile_offsets, int n_bags, const float* Wk,
                          const float* bk, int K, float* A_raw, float* partials, float* M, float* ml, float* hazards,
                          float* S, float* risk, int64_t* Y_hat, void* stream);

/* Combines n softmax partials (rows of L+2 floats) into one.
 *   normalize != 0: M[L] = Σ acc_t e^{m_t-m} / l,  ml[2] = (m, l)           (final result)
 *   normalize == 0: out[L+2] = (m, l, Σ acc_t e^{m_t-m})                   (rank-local partial,
 *                   all-gathered across ranks and combined again with normalize = 1)
 * Replaces the tail of F.softmax + torch.mm (models/model_attention_mil_path.py:55-56). */
int mmf_amil_combine(const float* partials, int64_t n, int L, int normalize, float* out_M_or_partial,
                     float* ml, void* stream);

size_t mmf_amil_bwd_workspace_bytes(int64_t N, int L, int D, int flags);

/* Training forward: same outputs as mmf_amil_fwd, and additionally leaves in `workspace` (sized by
 * mmf_amil_bwd_workspace_bytes, 1024-byte aligned, kept alive by the caller until the backward)
 * h as bf16 [N,L] and the pre-dropout branch outputs [tanh | sigmoid] as fp16 [N,2D] — 2L + 4D bytes
 * per instance (2.5 KB big preset). mmf_amil_bwd with MMF_STASHED then skips both recompute GEMMs:
 * its gate stage becomes one HBM-bound elementwise pass. Trades 40 MB of stores per 16k bag for
 * 30 GFLOP of recompute; use mmf_amil_fwd + mmf_amil_bwd (no flag) when memory is the constraint.
 * Replaces the same reference ops as mmf_amil_fwd, with autograd's saved activations made explicit.
 * zero_buf / zero_count (optional, NULL / 0): an fp32 buffer (16-byte aligned, count a multiple of 4) that the
 * kernel clears while its first GEMM runs — the step's gradient accumulators, i.e. optimizer.zero_grad() fused
 * into the forward (a separate fill costs two kernel boundaries per step). It must not alias anything the
 * forward reads.
 * The workspace belongs to ONE step at a time: the training forward treats what the previous step left in it as dead — it
 * overwrites the stash in place and DISCARDS the L2 lines of the dU slot (discard.global.L2: no write-back of data that the
 * previous wgrad has consumed and this step's backward rewrites). Do not start a step's forward on a workspace whose
 * backward is still pending. */
int mmf_amil_fwd_train(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                       int flags, uint64_t seed, float* A_raw, float* partials, void* workspace,
                       size_t workspace_bytes, float* zero_buf, int64_t zero_count, void* stream);

/* ---- fused batch-1 training step (utils/core_utils.py:200-247 for the path / radio AMIL models) ------------------
 * THREE launches: mmf_amil_fwd_train_head (fused forward; also leaves z_i = Wk h_i and the ReLU mask words for the
 * backward), then mmf_amil_bwd_head = [gate + hidden backward whose prologue runs the head] + [grouped wgrad GEMM].
 * Head block: softmax combine, classifier -> sigmoid -> cumprod (models/model_attention_mil_path.py:55-61), nll_surv
 * (utils/loss_utils.py:22-39) and their backward — computed redundantly by every CTA of the gate + hidden kernel from the
 * forward's 12-float head rows (m_t, l_t, Wk.acc_t) while its first tiles are in flight (no head launch); the pooled
 * embedding M and dWk are formed off the critical path inside the same kernel.
 * Inputs: Wk, bk, Wk_split (mmf_pack_head_weights), K <= 8, Y, c, alpha, eps, loss_scale (1/gc of the reference's
 * gradient accumulation: every gradient of the step is scaled by it).
 * Outputs (valid after mmf_amil_bwd_head): M [L], ml [2], hazards / S [K], Y_hat (or NULL), loss [1] (unscaled),
 * dM [L], hs [16] (dlogits, dM.M), dWk [K,L] / dbk [K] ACCUMULATED (or NULL).
 * Limits: N <= 65536 (512 per-tile head rows merged per CTA); larger bags: mmf_amil_fwd_train + mmf_amil_head_nll_step +
 * mmf_amil_bwd. */
typedef struct MmfHeadStep {
  const float* Wk;
  const float* bk;
  const void* Wk_split; /* bf16 [16, L] */
  int K;
  const int64_t* Y;
  const float* c;
  float alpha, eps, loss_scale;
  float* M;
  float* ml;
  float* hazards;
  float* S;
  int64_t* Y_hat;
  float* loss;
  float* dM;
  float* hs;
  float* dWk;
  float* dbk;
} MmfHeadStep;

/* Wk f32 [K, L] -> bf16 [16, L]: rows 0..K-1 = bf16(Wk), rows 8..8+K-1 = bf16(Wk - bf16(Wk)), other rows zero: the B
 * operand of the N = 16 tensor-core side product z_i = Wk h_i of the training forward (hi + lo: fp32-grade z). */
int mmf_pack_head_weights(const float* Wk, int K, int L, void* Wk_split_bf16, void* stream);

/* mmf_amil_fwd_train that additionally leaves z_i = Wk h_i (fp32 [N, 4|8]) and one head row per tile (m_t, l_t, Wk.acc_t)
 * in the workspace (uses head->Wk, head->Wk_split, head->K). */
int mmf_amil_fwd_train_head(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                            int flags, uint64_t seed, float* A_raw, float* partials, void* workspace,
                            size_t workspace_bytes, float* zero_buf, int64_t zero_count, const MmfHeadStep* head,
                            void* stream);

/* Head + backward of mmf_amil_fwd_train_head (same workspace, A_raw and partials): the pooled embedding feeds the
 * linear classifier directly, so dM = Wk^T dlogits and t_i = dM.h_i = dlogits.z_i — the per-row 512-long dot products
 * of the general backward become K FMAs. flags as in the forward (MMF_STASHED implied). Accumulates into g. */
int mmf_amil_bwd_head(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                      uint64_t seed, const float* A_raw, const float* partials, const MmfHeadStep* head,
                      const float* dA_raw, const MmfAmilGrads* g, void* dx, void* workspace, size_t workspace_bytes,
                      void* stream);

/* The head + gate + hidden stage of mmf_amil_bwd_head alone (the wgrad stage is mmf_amil_bwd_wgrad): stage timing, tests. */
int mmf_amil_bwd_gate_hidden_head(int64_t N, const MmfAmilWeights* w, int L, int D, int flags, uint64_t seed,
                                  const float* A_raw, const float* partials, const MmfHeadStep* head,
                                  const float* dA_raw, const MmfAmilGrads* g, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* Backward of mmf_amil_fwd + combine, given dM = dLoss/dM [L] and optionally dA_raw [N].
 * Recomputes h and the attention activations tile by tile (nothing but A_raw, (m,l), M is kept
 * from the forward). Accumulates INTO g (caller zeroes or carries gradient accumulation).
 *   M, ml are the GLOBAL (all-rank) pooled vector and (max, sum) — a rank that owns a shard of the
 *   bag passes the combined values and obtains its shard's contribution to the weight grads.
 *   dx bf16 [N,1024] is written only with MMF_NEED_DX.
 * Replaces autograd through the same reference ops (utils/core_utils.py:242-247 loss.backward()). */
int mmf_amil_bwd(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                 int flags, uint64_t seed, const float* A_raw, const float* ml, const float* M,
                 const float* dM, const float* dA_raw, const void* H_stash, const MmfAmilGrads* g,
                 void* dx, void* workspace, size_t workspace_bytes, void* stream);

/* The three stages of mmf_amil_bwd, individually callable (same workspace; stage k consumes what
 * stages < k left in it). Exposed so that each tensor-core kernel can be timed and tested alone.
 *   gate   : recompute tile kernel -> dG, H in workspace; dwc, dbab, dbc accumulated (recompute mode only)
 *   hidden : dU = (dG Wab + p dM^T) ⊙ relu'(H) in workspace; db1 accumulated
 *   wgrad  : dW1 += dU^T x, dWab += dG^T H, optional dx = dU W1 */
int mmf_amil_bwd_gate(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                      int flags, uint64_t seed, const float* A_raw, const float* ml, const float* M,
                      const float* dM, const float* dA_raw, const MmfAmilGrads* g, void* workspace,
                      size_t workspace_bytes, void* stream);
/* gate + hidden stages of the MMF_STASHED backward in one kernel (the gate backward is the A-operand producer of
 * the dU GEMM); what mmf_amil_bwd(MMF_STASHED) runs: dG and dU left in the workspace for the wgrad stage. */
int mmf_amil_bwd_gate_hidden_stashed(int64_t N, const MmfAmilWeights* w, int L, int D, int flags, uint64_t seed,
                                     const float* A_raw, const float* ml, const float* M, const float* dM,
                                     const float* dA_raw, const MmfAmilGrads* g, void* workspace,
                                     size_t workspace_bytes, void* stream);
int mmf_amil_bwd_hidden(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                        int flags, const float* A_raw, const float* ml, const float* dM,
                        const MmfAmilGrads* g, void* workspace, size_t workspace_bytes, void* stream);
int mmf_amil_bwd_wgrad(const void* x, int64_t N, int64_t ldx, const MmfAmilWeights* w, int L, int D,
                       int flags, const MmfAmilGrads* g, void* dx, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- varlen-packed TRAINING of a window of bags ------------------------------------------------------------------
 * The bags of a gradient-accumulation window (`--gc` bags between optimizer steps, utils/core_utils.py:242-247: loss / gc,
 * backward, step every gc bags) as ONE launch set: the bags are packed into one [R, 1024] buffer, each starting on a
 * 128-row boundary (zero padding; the format of mmf_amil_infer_varlen); tile_valid[t] = rows of tile t that belong to its
 * bag, tile_bag[t] = its bag, seg_tile_offsets[b .. b + 1] = the bag's tile range (device int32 arrays; tile_valid /
 * tile_bag padded to an EVEN number of tiles with valid = 0). Call order on one stream:
 *   mmf_amil_window_fwd_train      fused forward + activation stash of every tile, per-tile softmax partials
 *   mmf_amil_window_head_nll_step  per bag (grid.y): combine, classifier, hazards, nll_surv, loss_scale * its gradient ->
 *                                  M / dM [bags, L], ml [bags, 2], hazards / S [bags, K], loss [bags]; dWk / dbk += (atomic)
 *   mmf_amil_window_bwd            gate + hidden backward with per-tile bag statistics, then the grouped weight gradients:
 *                                  the window's summed gradients accumulate into g (MMF_STASHED required); with
 *                                  MMF_NEED_DX also dx = dU W1 (bf16 [R, 1024]; zero on padding rows) for an upstream layer
 * R is a multiple of 128; workspace as mmf_amil_bwd_workspace_bytes(R, ...). */
int mmf_amil_window_fwd_train(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                              uint64_t seed, const int32_t* tile_valid, float* A_raw, float* partials, void* workspace,
                              size_t workspace_bytes, float* zero_buf, int64_t zero_count, void* stream);
int mmf_amil_window_head_nll_step(const float* partials, const int32_t* seg_tile_offsets, int n_bags, int max_tiles,
                                  int L, const float* Wk, const float* bk, int K, const int64_t* Y, const float* c,
                                  float alpha, float eps, float loss_scale, float* M, float* ml, float* hazards, float* S,
                                  int64_t* Y_hat, float* loss, float* dM, float* dWk, float* dbk, void* stream);
int mmf_amil_window_bwd(const void* x, int64_t R, int64_t ldx, const MmfAmilWeights* w, int L, int D, int flags,
                        uint64_t seed, const float* A_raw, const float* ml, const float* M, const float* dM,
                        const int32_t* tile_bag, const int32_t* tile_valid, const MmfAmilGrads* g, void* dx,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Dense bf16 tensor-core GEMM used either side of the AMIL core (radio reduce_dim and its
 * gradients): C[M,N] = A[M,K] B[N,K]^T + bias (A given as up to 4 K-segments = the modality
 * bags that the reference concatenates, models/model_attention_mil_radio.py:81-82).
 *   out_bf16 / out_f32: exactly one non-NULL. */
int mmf_linear_bf16(const void* const* A_segs, int n_segs, int64_t M, int K_per_seg, int64_t lda,
                    const void* W /*bf16 [N,K]*/, const float* bias /*[N] or NULL*/, int N,
                    void* out_bf16, float* out_f32, int64_t ldc, void* stream);

/* dW[N,K] += dY[M,N]^T X[M,K] (X as up to 4 K-segments), db[N] += colsum(dY). dY bf16 [M,N]. */
int mmf_linear_bf16_wgrad(const void* dY, int64_t M, int N, int64_t lddy, const void* const* X_segs,
                          int n_segs, int K_per_seg, int64_t ldx, float* dW /*[N, n_segs*K_per_seg]*/,
                          float* db /*[N] or NULL*/, void* workspace, size_t workspace_bytes,
                          void* stream);
size_t mmf_linear_bf16_wgrad_workspace_bytes(int64_t M, int N);

/* ---- small fp32 kernels: heads, SNN, Kronecker fusion, survival losses -------------------- */

/* activation codes for mmf_dense_* */
enum { MMF_ACT_NONE = 0, MMF_ACT_RELU = 1, MMF_ACT_SELU = 2, MMF_ACT_SIGMOID = 3, MMF_ACT_TANH = 4 };

/* y[B,out] = act(x[B,in] W[out,in]^T + b) * mask   (mask f32 [B,out] or NULL; carries the
 * inverted-dropout / alpha-dropout scaling when training).
 * Replaces nn.Linear + activation (+Dropout/AlphaDropout) blocks: SNN_Block
 * (models/model_modules.py:64-68), classifier heads, XlinearFusion.reduce/encoder layers. */
int mmf_dense_fwd(const float* x, int64_t ldx, const float* W, const float* b, int B, int in_dim,
                  int out_dim, int act, float* y, int64_t ldy, void* stream);
/* The same layer with a caller-owned workspace: when the output has few 64 x 64 tiles and the k loop is long (the radiology
 * reduce_dim Linear(4096, 1024) on a ~100-slice patient, models/model_attention_mil_radio.py:31,81-82; the fc of a tiny bag)
 * the k loop is split over the whole machine — every slice stores its partial tile to the workspace and a fix-up kernel
 * adds the slices in a FIXED order (deterministic, no atomics) before bias + activation. mmf_dense_fwd_workspace_bytes()
 * returns 0 when the shape is not split (the call is then mmf_dense_fwd). */
size_t mmf_dense_fwd_workspace_bytes(int B, int in_dim, int out_dim);
int mmf_dense_fwd_ws(const float* x, int64_t ldx, const float* W, const float* b, int B, int in_dim, int out_dim,
                     int act, float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);
/* Given y (post-activation) and dy: dpre = dy * act'(y); dx[B,in] (=|+=) dpre W; dW += dpre^T x;
 * db += colsum(dpre).  dx may be NULL. accumulate_dx != 0 adds into dx. */
int mmf_dense_bwd(const float* x, int64_t ldx, const float* W, int B, int in_dim, int out_dim,
                  int act, const float* y, int64_t ldy, const float* dy, int64_t lddy, float* dx,
                  int64_t lddx, int accumulate_dx, float* dW, float* db, void* stream);

/* Kronecker ("Xlinear") fusion encoder1: out[B,H] = relu(W1 · (o_1 ⊗ o_2 [⊗ o_3 [⊗ o_4]]) + b1), m = 2..4, with
 * o_i[B,E] (E = dim+1, last column = 1) never materialising the E^m-wide outer product.
 * Replaces torch.bmm outer products + encoder1 Linear+ReLU (models/model_modules.py:167-173). */
int mmf_kron_enc_fwd(const float* const* o /*HOST array of m device ptrs [B,E]*/, int m, int E, int B,
                     const float* W /*[H,E^m]*/, const float* b, int H, float* out /*[B,H]*/,
                     void* stream);
/* Train-mode form: dropout on the (never materialised) product — element (b, kk) kept iff the counter hash of
 * (seed, stream 3, b, kk) says so (XlinearFusion.post_fusion_dropout, models/model_modules.py:170); the backward
 * regenerates the mask from the same seed. dropout = 0: identical to mmf_kron_enc_fwd / _bwd; dropout = 1: p = 0.25
 * (2-bit fields, scale 1/0.75: the reference's default rate); dropout = 2..65535: p = dropout / 65536 with one 16-bit
 * field per element and scale 65536 / (65536 - dropout) — any rate (the cohort heads build XlinearFusion with 0.7,
 * models/coxranking_models_pretrained.py:107). */
int mmf_kron_enc_train_fwd(const float* const* o, int m, int E, int B, const float* W, const float* b, int H,
                           int dropout, uint64_t seed, float* out, void* stream);
/* The same with a caller-owned workspace for the deterministic split-K form (few output tiles, K = E^m long: see
 * mmf_dense_fwd_ws); mmf_kron_enc_fwd_workspace_bytes() returns 0 when the shape is not split. */
size_t mmf_kron_enc_fwd_workspace_bytes(int m, int E, int B, int H);
int mmf_kron_enc_train_fwd_ws(const float* const* o, int m, int E, int B, const float* W, const float* b, int H,
                              int dropout, uint64_t seed, float* out, void* workspace, size_t workspace_bytes,
                              void* stream);
int mmf_kron_enc_train_bwd(const float* const* o, int m, int E, int B, const float* W, int H, int dropout,
                           uint64_t seed, const float* out, const float* dout, float* const* d_o, float* dW,
                           float* db, void* workspace, size_t workspace_bytes, void* stream);
/* workspace: B * E^m floats (the gradient w.r.t. the outer product, contracted immediately). */
size_t mmf_kron_enc_workspace_bytes(int m, int E, int B);
int mmf_kron_enc_bwd(const float* const* o, int m, int E, int B, const float* W, int H,
                     const float* out, const float* dout, float* const* d_o /*m ptrs [B,E]*/,
                     float* dW, float* db, void* workspace, size_t workspace_bytes, void* stream);

/* Discrete-hazard head: logits = M Wk^T + bk; hazards = sigmoid(logits); S = cumprod(1-hazards);
 * Y_hat = argmax logits.  Replaces models/model_attention_mil_path.py:58-61. */
int mmf_hazard_head_fwd(const float* M, int B, int Lin, const float* Wk, const float* bk, int K,
                        float* hazards, float* S, int64_t* Y_hat, void* stream);
int mmf_hazard_head_bwd(const float* M, int B, int Lin, const float* Wk, int K, const float* hazards,
                        const float* S, const float* d_hazards, const float* d_S, float* dM,
                        float* dWk, float* dbk, void* stream);

/* Batch-1 training-step tail in ONE launch: combine the n tile partials -> M[L], ml[2]; hazard head;
 * nll_surv loss (alpha, eps); gradient dM[L] back to the pooled vector; dWk[K,L], dbk[K] accumulated
 * (may be NULL). Y int64[1], c f32[1] on the device. n <= 4096, L <= 1024, K <= 16.
 * Replaces, per bag of the reference's hot loop (utils/core_utils.py:200-247): softmax+mm tail,
 * classifier/sigmoid/cumprod/topk (models/model_attention_mil_path.py:55-61), nll_loss
 * (utils/loss_utils.py:22-39) and their autograd — about forty ATen launches. */
int mmf_amil_head_nll_step(const float* partials, int64_t n, int L, const float* Wk, const float* bk, int K,
                           const int64_t* Y, const float* c, float alpha, float eps, float* M, float* ml,
                           float* hazards, float* S, int64_t* Y_hat, float* loss, float* dM, float* dWk,
                           float* dbk, void* stream);

/* nll_loss (utils/loss_utils.py:22-39): loss scalar + d_hazards, d_S [B,K]. Y int64 [B], c f32 [B]. */
int mmf_nll_surv_fwd_bwd(const float* hazards, const float* S, const int64_t* Y, const float* c,
                         int B, int K, float alpha, float eps, float* loss, float* d_hazards,
                         float* d_S, void* stream);

/* ce_loss (utils/loss_utils.py:41-56): cross-entropy survival loss, scalar + d_hazards, d_S [B,K]. */
int mmf_ce_surv_fwd_bwd(const float* hazards, const float* S, const int64_t* Y, const float* c, int B, int K,
                        float alpha, float eps, float* loss, float* d_hazards, float* d_S, void* stream);

/* nn.BatchNorm1d over [B,F] (the fcnn / Highway fusion heads: models/coxranking_models_pretrained.py:80-94,
 * models/nll_models_pretrained.py:82-99, models/model_modules.py:13-14,18,26). train != 0: batch statistics, running
 * estimates updated in place (momentum, unbiased variance), B >= 2; train == 0: running statistics. save_mean /
 * save_invstd [F] receive the statistics used (inputs of the backward). gamma / beta may be NULL (affine off). */
int mmf_batchnorm1d_fwd(const float* x, int B, int F, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, int train, float momentum, float eps, float* y, float* save_mean,
                        float* save_invstd, void* stream);
/* dx (nullable) written; dgamma / dbeta (nullable) accumulated. */
int mmf_batchnorm1d_bwd(const float* x, const float* dy, int B, int F, const float* gamma, const float* save_mean,
                        const float* save_invstd, int train, float* dx, float* dgamma, float* dbeta, void* stream);

/* Highway layer mix y = gate * nonlinear + (1 - gate) * linear and its backward (models/model_modules.py:21-25). */
int mmf_highway_mix_fwd(const float* gate, const float* nonlinear, const float* linear, int64_t count, float* y,
                        void* stream);
int mmf_highway_mix_bwd(const float* gate, const float* nonlinear, const float* linear, const float* dy, int64_t count,
                        float* dgate, float* dnonlinear, float* dlinear, void* stream);

/* CoxSurvLoss (utils/loss_utils.py:124-139): loss = -mean_i (theta_i - log Σ_{t_j>=t_i} e^{theta_j})(1-c_i).
 * O(B log B) sort + tie-aware suffix sums instead of the reference's O(B^2) host loop.
 * workspace: mmf_cox_workspace_bytes(B). dtheta may be NULL. */
size_t mmf_cox_workspace_bytes(int B);
int mmf_cox_fwd_bwd(const float* theta, const float* times, const float* c, int B, float* loss,
                    float* dtheta, void* workspace, size_t workspace_bytes, void* stream);

/* ranking_loss (utils/loss_utils.py:58-101): over comparable pairs (t_a < t_b and event_a),
 * loss = -mean|sum phi(r_a - r_b); phi: 0 = sigmoid, 1 = relu; reduction: 0 = mean, 1 = sum.
 * n_pairs (device int64) receives the pair count (0 -> loss 0, zero gradient). */
size_t mmf_ranking_workspace_bytes(int B);
int mmf_ranking_fwd_bwd(const float* risks, const float* times, const float* c, int B, int phi,
                        int reduction, float* loss, float* drisks, int64_t* n_pairs, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- XlinearFusion gated reduction, all modalities in one launch ---------------
 * models/model_modules.py:156-166: per modality i  h_i = relu(Wh_i v_i + bh_i), z_i = sigmoid(Wz_i cat(v_1..v_m) + bz_i),
 * o_i = dropout(relu(Wo_i (z_i * h_i) + bo_i)), then a constant 1 is appended (the Kronecker factor [B, S + 1]).
 * S = dim / scale_dim must be 16 (every configuration of the reference), dim a multiple of 256, m = 2..4.
 * The concatenation is never formed. mask (nullable): [m, B, 16] dropout scale mask (0 or 1 / (1 - p)), made by the caller.
 * Forward outputs: h, z [m, B, 16] (kept for the backward), o [m, B, 17]. */
typedef struct MmfXfusionMod {
  const float* v;                    /* [B, dim] embedding of this modality                          */
  const float* Wh; const float* bh;  /* reduce[i][0][0]: Linear(dim, 16)                              */
  const float* Wz; const float* bz;  /* reduce[i][1][0]: Linear(dim * m, 16)                          */
  const float* Wo; const float* bo;  /* reduce[i][2][0]: Linear(16, 16)                               */
} MmfXfusionMod;
typedef struct MmfXfusionGrads {     /* per modality; all written (accumulate = 0) or added to (accumulate = 1)        */
  float* dWh; float* dbh; float* dWz; float* dbz; float* dWo; float* dbo;
  float* dv;                         /* [B, dim] gradient of the embedding (always overwritten), or NULL if not needed */
} MmfXfusionGrads;
int mmf_xfusion_gate_fwd(const MmfXfusionMod* mods_host, int m, int B, int dim, const float* mask, float* h, float* z,
                         float* o, void* stream);
/* d_o: [m, B, 17] (the constant column's gradient is ignored); workspace: mmf_xfusion_gate_bwd_workspace_bytes(m, B, dim)
 * ([dz | dh] per sample + the partial sums of up to 16 batch slices, added in a fixed order: deterministic). */
size_t mmf_xfusion_gate_bwd_workspace_bytes(int m, int B, int dim);
int mmf_xfusion_gate_bwd(const MmfXfusionMod* mods_host, int m, int B, int dim, const float* mask, const float* h,
                         const float* z, const float* o, const float* d_o, const MmfXfusionGrads* grads_host,
                         int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- self-normalising MLP (SNN_Block x n) in one launch ---------------------------
 * models/model_modules.py:64-68 (Linear -> SELU -> AlphaDropout), models/model_genomic.py:17-25,53-57 (fc_omic), the omics
 * branch of models/model_mm_attention_mil.py. n = 1..4 layers, hidden widths <= 1024, any input width.
 * keep (nullable, per layer): [B, width] keep mask (0 / 1) of that layer's AlphaDropout(p), drawn by the caller; the kernel
 * applies torch's affine a (y m + alpha' (1 - m)) + b. y: [B, width] pre-dropout SELU outputs, kept for the backward.
 * out: [B, width of the last layer] = the network output (after the last layer's dropout). */
typedef struct MmfSnnLayer {
  const float* W;     /* [width, width of the previous layer (input width for layer 0)] */
  const float* b;     /* [width] */
  const float* keep;  /* [B, width] or NULL (eval mode / p = 0) */
  float p;            /* AlphaDropout rate (used when keep != NULL) */
  float* y;           /* [B, width] saved activations (forward: out, backward: in) */
  int width;
} MmfSnnLayer;
int mmf_snn_mlp_fwd(const float* x, int B, int in_dim, const MmfSnnLayer* layers_host, int n_layers, float* out,
                    void* stream);
/* dout: [B, last width]. dW_host / db_host: HOST arrays of n_layers device pointers ([width, in] / [width]), written
 * (accumulate = 0) or added to. dx: [B, in_dim] or NULL. workspace: B * sum(widths) floats (the pre-activation gradients). */
int mmf_snn_mlp_bwd(const float* x, int B, int in_dim, const MmfSnnLayer* layers_host, int n_layers, const float* dout,
                    float* const* dW_host, float* const* db_host, int accumulate, float* dx, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- training-step glue -------------------------------------------------
 * Fused multi-tensor Adam: torch.optim.Adam(lr, weight_decay) as the reference builds it (utils/utils.py:144-151), one
 * launch for all parameter tensors; step >= 1 is the 1-based step count (bias corrections). Per element:
 *   g = grad * grad_scale + l1_lambda * sign(p) + weight_decay * p;  m, v updated;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * l1_lambda folds the reference's l1_reg_all penalty (utils/utils.py:249-257; lambda * sum|W| added to the loss at
 * utils/core_utils.py:218-221,242) into the update; l1_out (nullable, device) += sum |p| over all tensors (pre-update)
 * for logging. zero_grad != 0 also clears the gradients (optimizer.zero_grad(), utils/core_utils.py:246).
 * The *_host arrays are HOST arrays of n_tensors device pointers / element counts. */
int mmf_adam_step_multi(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                        float* const* exp_avg_sq_host, const int64_t* numel_host, int n_tensors, int step, float lr,
                        float beta1, float beta2, float eps, float weight_decay, float grad_scale, float l1_lambda,
                        int zero_grad, float* l1_out, void* stream);

/* The same update with the step count read from the device (graph-captured training steps): *step_dev >= 1 when the kernel
 * runs; the bias corrections are formed in the kernel. */
int mmf_adam_step_multi_dev(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                            float* const* exp_avg_sq_host, const int64_t* numel_host, int n_tensors,
                            const uint64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float grad_scale, float l1_lambda, int zero_grad, float* l1_out, void* stream);
/* Device-resident state of a graph-captured training step (the batch-1 loop of utils/core_utils.py:184-247 replayed as ONE
 * graph launch per patient / cohort step): state[0] = optimizer step count, state[1 .. n_seeds] = dropout seeds. One launch
 * at the head of each replay: state[0] += 1 and every seed moves to the next value of its splitmix64 sequence (< 2^62).
 * Consumers: mmf_adam_step_multi_dev(step_dev = state) and any seed passed as MMF_SEED_DEVICE(state + i). */
int mmf_step_state_advance(uint64_t* state, int n_seeds, void* stream);

/* Concordance index counts as sksurv.concordance_index_censored(event, time, risk, tied_tol) computes them
 * (utils/core_utils.py:258): over comparable pairs (event_i != 0 and t_i < t_j): counts[0] concordant
 * (risk_i > risk_j + tol), counts[1] discordant, counts[2] tied in risk. c-index = (c + 0.5 t) / (c + d + t). */
int mmf_cindex_counts(const float* risk, const float* times, const float* event, int B, float tied_tol,
                      uint64_t* counts, void* stream);

/* Attention scores -> percentiles for heatmaps: out[q] = scipy.stats.percentileofscore(ref_scores, query[q]) (kind='rank')
 * = (left + right + [left < right]) * 50 / n_ref with left = #{ref < x}, right = #{ref <= x}; replaces the per-patch host loops of utils/wsi_utils.py:171-174
 * (to_percentiles: query == ref) and utils/heatmap_utils.py:32-34,99,138 (score2percentile against reference scores). */
int mmf_percentile_of_score(const float* ref_scores, int n_ref, const float* query, int n_query, float* out,
                            void* stream);

/* ---- multi-GPU: SUM all-reduce of a small fp32 buffer over NVLink peer memory ----------------
 * The gradient all-reduce that closes a cohort-data-parallel step (the reference is
 * single-GPU and has no counterpart) as one kernel on the caller's stream: ready handshake, reduce of
 * this rank's slice from all peers + push of the sum to all peers, done handshake. Graph-capturable.
 *   bufs_host[p]  : rank p's buffer (n floats, same n everywhere) as mapped into THIS process (peer memory /
 *                   symmetric memory), p < world <= 8; every rank calls with its own mapping of all buffers
 *   flags_host[p] : rank p's flag block, mmf_p2p_flag_bytes() bytes, zero-initialised once, also peer-mapped
 *   multicast_ptr : NVLS multicast mapping of the same buffer (all ranks) or NULL; when given, the NVSwitch does
 *                   the reduction (multimem.ld_reduce / multimem.st) and each rank moves only its 1/world slice
 *   n             : multiple of 4;  n_ctas <= 64 (0 = default 32); same n_ctas on every rank.
 * All ranks must call it the same number of times (it is a collective). */
size_t mmf_p2p_flag_bytes(void);
int mmf_p2p_allreduce_sum_f32(void* const* bufs_host, void* const* flags_host, void* multicast_ptr, int world,
                              int rank, int64_t n, int n_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMF_B200_H_ */
