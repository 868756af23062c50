"""Torch-facing wrappers over the C ABI: tensors in, tensors out, all on the current stream.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every FLOP of
the path runs in the kernels of ``libmmf_b200.so``.  All functions raise on CPU tensors — there
is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SELU, ACT_SIGMOID, ACT_TANH, MMF_DROPOUT_ATTN,
                   MMF_DROPOUT_H, MMF_GATED, MMF_NEED_DX, MMF_PRECISE_FC, MMF_STASHED, AmilGrads, AmilWeights, HeadStep, check, lib)

IN_FEATURES = 1024
TILE_ROWS = 128


def _require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "multimodalfusion_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback. "
                "Move the model with .relocate()/.cuda() and the inputs with .cuda().")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# formats
# ------------------------------------------------------------------------------------------------
def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (RNE) with the library's cast kernel; bf16 input is returned as is."""
    _require_cuda(x)
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    x = _f32c(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().mmf_cast_f32_to_bf16(_p(x), _p(out), x.numel(), _stream()), "mmf_cast_f32_to_bf16")
    return out


@dataclass
class AmilPrepared:
    """bf16 / packed copies of the fc + attention-net weights, rebuilt when parameters change."""
    L: int
    D: int
    gated: bool
    W1: torch.Tensor          # bf16 [L,1024]
    b1: torch.Tensor          # f32 [L]
    Wab: torch.Tensor         # bf16 [2D,L] / [D,L]
    Wab_packed: torch.Tensor  # bf16 same shape, chunk-interleaved rows
    bab: torch.Tensor         # f32 [2D] / [D]
    wc: torch.Tensor          # f32 [D]
    bc: torch.Tensor          # f32 [1]

    W1_split: Optional[torch.Tensor] = None   # bf16 [L,3072] = [W1_hi | W1_hi | W1_lo] (split-precision fc, small bags)

    def struct(self) -> AmilWeights:
        return AmilWeights(_p(self.W1), _p(self.b1), _p(self.Wab), _p(self.Wab_packed), _p(self.bab),
                           _p(self.wc), _p(self.bc), _p(self.W1_split))


def prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc) -> AmilPrepared:
    """Wa/ba are the tanh branch (or the only branch when Wb is None = un-gated Attn_Net)."""
    _require_cuda(W1, Wa, wc)
    gated = Wb is not None
    L, D = W1.shape[0], Wa.shape[0]
    if W1.shape[1] != IN_FEATURES:
        raise ValueError(f"fc input width must be {IN_FEATURES}, got {W1.shape[1]}")
    W1b = to_bf16(W1)
    if gated:
        Wab32 = torch.cat([_f32c(Wa), _f32c(Wb)], dim=0)
        bab = torch.cat([_f32c(ba), _f32c(bb)], dim=0)
    else:
        Wab32, bab = _f32c(Wa), _f32c(ba)
    Wab = to_bf16(Wab32)
    packed = torch.empty_like(Wab)
    check(lib().mmf_pack_wab(_p(Wab), _p(packed), L, D, int(gated), _stream()), "mmf_pack_wab")
    W1f = _f32c(W1)
    W1lo = to_bf16(W1f - W1b.float())
    return AmilPrepared(L, D, gated, W1b, _f32c(b1), Wab, packed, bab, _f32c(wc).reshape(-1),
                        _f32c(bc).reshape(-1), torch.cat([W1b, W1b, W1lo], dim=1).contiguous())


PRECISE_FC_MAX_ROWS = 4096   # bags up to this size run the split-precision fc (MMF_PRECISE_FC): see split_bag


def split_bag(x: torch.Tensor) -> torch.Tensor:
    """fp32 (or bf16) bag [N,1024] -> bf16 [N,3072] = [hi | lo | hi]: the input format of MMF_PRECISE_FC."""
    _require_cuda(x)
    x = _f32c(x)
    out = torch.empty(x.shape[0], 3 * IN_FEATURES, dtype=torch.bfloat16, device=x.device)
    check(lib().mmf_split_f32_bf16x3(_p(x), x.shape[0], x.stride(0), _p(out), _stream()), "mmf_split_f32_bf16x3")
    return out


# ------------------------------------------------------------------------------------------------
# fused attention-MIL forward / backward
# ------------------------------------------------------------------------------------------------
def amil_flags(gated: bool, dropout_h: bool = False, dropout_attn: bool = False, need_dx: bool = False) -> int:
    return ((MMF_GATED if gated else 0) | (MMF_DROPOUT_H if dropout_h else 0)
            | (MMF_DROPOUT_ATTN if dropout_attn else 0) | (MMF_NEED_DX if need_dx else 0))


def amil_partials(x: torch.Tensor, w: AmilPrepared, flags: int, seed: int = 0,
                  h_stash: Optional[torch.Tensor] = None):
    """Runs the fused tile kernel: returns (A_raw [N] f32, partials [tiles, L+2] f32)."""
    _require_cuda(x)
    _check_bag(x, flags)
    N = x.shape[0]
    tiles = (N + TILE_ROWS - 1) // TILE_ROWS
    A_raw = torch.empty(N, dtype=torch.float32, device=x.device)
    partials = torch.empty(tiles, w.L + 2, dtype=torch.float32, device=x.device)
    ws = w.struct()
    check(lib().mmf_amil_fwd(_p(x), N, x.stride(0), C.byref(ws), w.L, w.D, flags, seed, _p(A_raw),
                             _p(partials), _p(h_stash), _stream()), "mmf_amil_fwd")
    return A_raw, partials


def _check_bag(x: torch.Tensor, flags: int) -> None:
    width = 3 * IN_FEATURES if flags & MMF_PRECISE_FC else IN_FEATURES
    if x.dtype != torch.bfloat16 or x.dim() != 2 or x.shape[1] != width or x.stride(1) != 1:
        raise ValueError(f"x must be a bf16 [N,{width}] tensor with unit inner stride")
    if x.shape[0] == 0:
        raise ValueError("empty bag")


def amil_bwd_workspace(N: int, w: AmilPrepared, flags: int, device) -> torch.Tensor:
    """1024-byte aligned uint8 view sized by mmf_amil_bwd_workspace_bytes."""
    nbytes = lib().mmf_amil_bwd_workspace_bytes(N, w.L, w.D, flags)
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def amil_partials_train(x: torch.Tensor, w: AmilPrepared, flags: int, seed: int = 0,
                        workspace: Optional[torch.Tensor] = None, zero: Optional[torch.Tensor] = None):
    """Training forward: (A_raw, partials, workspace) — the workspace now holds h (bf16) and the branch
    activations (fp16) and must be handed to amil_backward(..., stash=workspace).
    zero: optional contiguous fp32 tensor (numel % 4 == 0) cleared by the kernel while GEMM1 runs — the step's
    flat gradient buffer (fused zero_grad)."""
    _require_cuda(x)
    _check_bag(x, flags)
    N = x.shape[0]
    if workspace is None:
        workspace = amil_bwd_workspace(N, w, flags, x.device)
    tiles = (N + TILE_ROWS - 1) // TILE_ROWS
    A_raw = torch.empty(N, dtype=torch.float32, device=x.device)
    partials = torch.empty(tiles, w.L + 2, dtype=torch.float32, device=x.device)
    ws = w.struct()
    check(lib().mmf_amil_fwd_train(_p(x), N, x.stride(0), C.byref(ws), w.L, w.D, flags, seed, _p(A_raw),
                                   _p(partials), workspace.data_ptr(), workspace.numel(), _p(zero),
                                   0 if zero is None else zero.numel(), _stream()),
          "mmf_amil_fwd_train")
    return A_raw, partials, workspace


@dataclass
class PackedBags:
    """Ragged bags packed for one varlen launch: x bf16 [R,1024] (every bag on a 128-row boundary, padding rows
    zero), tile_valid int32 [R/128], seg_tile_offsets int32 [n+1], row_offsets / sizes (host lists)."""
    x: torch.Tensor
    tile_valid: torch.Tensor
    seg_tile_offsets: torch.Tensor
    row_offsets: list
    sizes: list
    tile_bag: Optional[torch.Tensor] = None      # int32 [tiles padded to an even count]: bag of every tile (training windows)
    tile_valid_even: Optional[torch.Tensor] = None   # tile_valid padded likewise (pad tile: 0 rows)


_PAD_ROWS = {}     # (device, dtype) -> [127, 1024] zeros: the padding rows of pack_bags are views of it


def pack_bags(bags: Sequence[torch.Tensor], index_from: Optional[PackedBags] = None) -> PackedBags:
    """(index_from: reuse the index arrays of a PackedBags of the same bag sizes — another modality of the same patients.)
    Packs CUDA feature bags ([N_i,1024], fp32 or bf16) into the varlen layout of mmf_amil_infer_varlen / the training
    window: ONE concatenation launch (bags interleaved with views of a zero block; + one cast for fp32 bags) and ONE
    host-to-device copy of the four index arrays."""
    _require_cuda(*bags)
    sizes = [int(b.shape[0]) for b in bags]
    if min(sizes) < 1:
        raise ValueError("empty bag")
    tiles = [(n + TILE_ROWS - 1) // TILE_ROWS for n in sizes]
    seg = [0]
    for t in tiles:
        seg.append(seg[-1] + t)
    R = seg[-1] * TILE_ROWS
    dev = bags[0].device
    row_off = [t0 * TILE_ROWS for t0 in seg[:-1]]
    if len({b.dtype for b in bags}) == 1 and all(b.dim() == 2 and b.shape[1] == IN_FEATURES for b in bags):
        key = (dev, bags[0].dtype)
        pad = _PAD_ROWS.get(key)
        if pad is None:
            pad = _PAD_ROWS[key] = torch.zeros(TILE_ROWS - 1, IN_FEATURES, dtype=bags[0].dtype, device=dev)
        pieces = []
        for b, t, n in zip(bags, tiles, sizes):
            pieces.append(b)
            if t * TILE_ROWS > n:
                pieces.append(pad[:t * TILE_ROWS - n])
        x = torch.cat(pieces, dim=0)
        if x.dtype != torch.bfloat16:
            x = to_bf16(x)                                   # fp32 -> bf16 (RNE), the library's cast kernel
    else:
        x = torch.zeros(R, IN_FEATURES, dtype=torch.bfloat16, device=dev)
        for b, r0, n in zip(bags, row_off, sizes):
            x[r0:r0 + n].copy_(b)       # casts fp32 -> bf16 (RNE) on the fly
    if index_from is not None:
        if index_from.sizes != sizes:
            raise ValueError("index_from was packed from bags of other sizes")
        return PackedBags(x, index_from.tile_valid, index_from.seg_tile_offsets, row_off, sizes, index_from.tile_bag,
                          index_from.tile_valid_even)
    n_tiles = seg[-1]
    n_even = n_tiles + (n_tiles & 1)
    nb = len(sizes)
    host = torch.zeros(n_tiles + (nb + 1) + 2 * n_even, dtype=torch.int32)
    tv, sg = host[:n_tiles], host[n_tiles:n_tiles + nb + 1]
    tb, tve = host[n_tiles + nb + 1:n_tiles + nb + 1 + n_even], host[n_tiles + nb + 1 + n_even:]
    tv.fill_(TILE_ROWS)
    sg.copy_(torch.tensor(seg, dtype=torch.int32))
    for i, (t0, t, n) in enumerate(zip(seg, tiles, sizes)):
        tv[t0 + t - 1] = n - (t - 1) * TILE_ROWS
        tb[t0:t0 + t] = i
    tve[:n_tiles] = tv
    d = host.to(dev)
    o1, o2, o3 = n_tiles, n_tiles + nb + 1, n_tiles + nb + 1 + n_even
    return PackedBags(x, d[:o1], d[o1:o2], row_off, sizes, d[o2:o3], d[o3:])


def amil_infer_varlen(packed: PackedBags, w: AmilPrepared, Wk: torch.Tensor, bk: torch.Tensor):
    """Cohort inference in two launches. Returns dict(A_raw [R] (index with packed.row_offsets / sizes), M [n,L],
    hazards [n,K], S [n,K], risk [n], Y_hat [n,1])."""
    _require_cuda(packed.x, Wk)
    R, n = packed.x.shape[0], len(packed.sizes)
    dev = packed.x.device
    K = Wk.shape[0]
    Wk, bk = _f32c(Wk), _f32c(bk)
    out = dict(A_raw=torch.empty(R, dtype=torch.float32, device=dev),
               M=torch.empty(n, w.L, dtype=torch.float32, device=dev),
               hazards=torch.empty(n, K, dtype=torch.float32, device=dev),
               S=torch.empty(n, K, dtype=torch.float32, device=dev),
               risk=torch.empty(n, dtype=torch.float32, device=dev),
               Y_hat=torch.empty(n, 1, dtype=torch.int64, device=dev))
    partials = torch.empty(R // TILE_ROWS, w.L + 2, dtype=torch.float32, device=dev)
    ws = w.struct()
    check(lib().mmf_amil_infer_varlen(_p(packed.x), R, packed.x.stride(0), C.byref(ws), w.L, w.D, amil_flags(w.gated),
                                      _p(packed.tile_valid), _p(packed.seg_tile_offsets), n, _p(Wk), _p(bk), K,
                                      _p(out["A_raw"]), _p(partials), _p(out["M"]), None, _p(out["hazards"]), _p(out["S"]),
                                      _p(out["risk"]), _p(out["Y_hat"]), _stream()), "mmf_amil_infer_varlen")
    return out


def amil_combine(partials: torch.Tensor, L: int, normalize: bool = True):
    """normalize: (M [L], ml [2]);  else one (L+2) partial row (m, l, acc)."""
    _require_cuda(partials)
    partials = partials.contiguous()
    n = partials.shape[0]
    if normalize:
        M = torch.empty(L, dtype=torch.float32, device=partials.device)
        ml = torch.empty(2, dtype=torch.float32, device=partials.device)
        check(lib().mmf_amil_combine(_p(partials), n, L, 1, _p(M), _p(ml), _stream()), "mmf_amil_combine")
        return M, ml
    out = torch.empty(L + 2, dtype=torch.float32, device=partials.device)
    check(lib().mmf_amil_combine(_p(partials), n, L, 0, _p(out), None, _stream()), "mmf_amil_combine")
    return out


def amil_forward(x: torch.Tensor, w: AmilPrepared, flags: int, seed: int = 0):
    A_raw, partials = amil_partials(x, w, flags, seed)
    M, ml = amil_combine(partials, w.L, True)
    return A_raw, M, ml


def amil_backward(x: torch.Tensor, w: AmilPrepared, flags: int, seed: int, A_raw, ml, M, dM,
                  dA_raw=None, grads: Optional[dict] = None, stash: Optional[torch.Tensor] = None):
    """Returns dict(dW1, db1, dWab, dbab, dwc, dbc[, dx]); accumulates into `grads` if given.
    stash = the workspace amil_partials_train filled: skips the recompute GEMMs (MMF_STASHED)."""
    _require_cuda(x, dM)
    N = x.shape[0]
    dev = x.device
    KD = (2 if w.gated else 1) * w.D
    if grads is None:
        grads = dict(
            dW1=torch.zeros(w.L, IN_FEATURES, dtype=torch.float32, device=dev),
            db1=torch.zeros(w.L, dtype=torch.float32, device=dev),
            dWab=torch.zeros(KD, w.L, dtype=torch.float32, device=dev),
            dbab=torch.zeros(KD, dtype=torch.float32, device=dev),
            dwc=torch.zeros(w.D, dtype=torch.float32, device=dev),
            dbc=torch.zeros(1, dtype=torch.float32, device=dev),
        )
    need_dx = bool(flags & MMF_NEED_DX)
    dx = torch.empty(N, IN_FEATURES, dtype=torch.bfloat16, device=dev) if need_dx else None   # (always [N,1024])
    if stash is not None:
        ws_view, flags = stash, flags | MMF_STASHED
    else:
        ws_view = amil_bwd_workspace(N, w, flags, dev)
    g = AmilGrads(_p(grads["dW1"]), _p(grads["db1"]), _p(grads["dWab"]), _p(grads["dbab"]),
                  _p(grads["dwc"]), _p(grads["dbc"]))
    wst = w.struct()
    dM = _f32c(dM).reshape(-1)
    dA = None if dA_raw is None else _f32c(dA_raw).reshape(-1)
    check(lib().mmf_amil_bwd(_p(x), N, x.stride(0), C.byref(wst), w.L, w.D, flags, seed, _p(A_raw), _p(ml),
                             _p(M), _p(dM), _p(dA), None, C.byref(g), _p(dx), ws_view.data_ptr(),
                             ws_view.numel(), _stream()), "mmf_amil_bwd")
    if need_dx:
        grads["dx"] = dx
    return grads


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core linear (radio reduce_dim) and its weight gradient
# ------------------------------------------------------------------------------------------------
def linear_bf16(segs: Sequence[torch.Tensor], W_bf16: torch.Tensor, bias: Optional[torch.Tensor],
                out_dtype=torch.bfloat16) -> torch.Tensor:
    """y[M,N] = cat(segs, dim=1) @ W^T + bias, segments are bf16 [M,Kseg] (never concatenated)."""
    _require_cuda(*segs)
    M, Kseg = segs[0].shape
    N = W_bf16.shape[0]
    for s in segs:
        if s.dtype != torch.bfloat16 or s.shape != (M, Kseg) or s.stride(1) != 1 or s.stride(0) != segs[0].stride(0):
            raise ValueError("segments must be bf16 [M,Kseg] with identical strides")
    out = torch.empty(M, N, dtype=out_dtype, device=segs[0].device)
    arr = _lib.ptr_array([s.data_ptr() for s in segs])
    ob, of = (_p(out), None) if out_dtype == torch.bfloat16 else (None, _p(out))
    check(lib().mmf_linear_bf16(arr, len(segs), M, Kseg, segs[0].stride(0), _p(W_bf16), _p(bias), N, ob, of,
                                N, _stream()), "mmf_linear_bf16")
    return out


def linear_bf16_wgrad(dY: torch.Tensor, segs: Sequence[torch.Tensor], dW: torch.Tensor,
                      db: Optional[torch.Tensor]) -> None:
    """dW[N, sum Kseg] += dY^T cat(segs); db[N] += colsum(dY).  dY bf16 [M,N]."""
    _require_cuda(dY, dW)
    M, N = dY.shape
    Kseg = segs[0].shape[1]
    arr = _lib.ptr_array([s.data_ptr() for s in segs])
    check(lib().mmf_linear_bf16_wgrad(_p(dY), M, N, dY.stride(0), arr, len(segs), Kseg, segs[0].stride(0),
                                      _p(dW), _p(db), None, 0, _stream()), "mmf_linear_bf16_wgrad")


# ------------------------------------------------------------------------------------------------
# small fp32 kernels
# ------------------------------------------------------------------------------------------------
def dense_fwd(x, W, b, act: int) -> torch.Tensor:
    _require_cuda(x, W)
    x, W = _f32c(x), _f32c(W)
    B, in_dim = x.shape
    out = torch.empty(B, W.shape[0], dtype=torch.float32, device=x.device)
    nbytes = lib().mmf_dense_fwd_workspace_bytes(B, in_dim, W.shape[0])   # > 0: deterministic split-K (few tiles, long k loop)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
    check(lib().mmf_dense_fwd_ws(_p(x), in_dim, _p(W), _p(None if b is None else _f32c(b)), B, in_dim, W.shape[0],
                                 act, _p(out), W.shape[0], _p(ws), nbytes, _stream()), "mmf_dense_fwd_ws")
    return out


def dense_bwd(x, W, act: int, y, dy, need_dx=True, need_dw=True, need_db=True):
    x, W, y, dy = _f32c(x), _f32c(W), _f32c(y), _f32c(dy)
    B, in_dim = x.shape
    out_dim = W.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dW = torch.zeros_like(W) if need_dw else None
    db = torch.zeros(out_dim, dtype=torch.float32, device=x.device) if need_db else None
    check(lib().mmf_dense_bwd(_p(x), in_dim, _p(W), B, in_dim, out_dim, act, _p(y), out_dim, _p(dy), out_dim,
                              _p(dx), in_dim, 0, _p(dW), _p(db), _stream()), "mmf_dense_bwd")
    return dx, dW, db


def dense_bwd_into(x, W, y, dy, dW: torch.Tensor, db: torch.Tensor) -> None:
    """Weight / bias gradient of y = x W^T + b ACCUMULATED into the caller's dW [out,in] and db [out] (no dx)."""
    x, W, y, dy = _f32c(x), _f32c(W), _f32c(y), _f32c(dy)
    B, in_dim = x.shape
    out_dim = W.shape[0]
    if not (dW.is_contiguous() and db.is_contiguous() and dW.dtype == torch.float32):
        raise ValueError("dense_bwd_into needs contiguous fp32 gradient tensors")
    check(lib().mmf_dense_bwd(_p(x), in_dim, _p(W), B, in_dim, out_dim, ACT_NONE, _p(y), out_dim, _p(dy), out_dim,
                              None, in_dim, 0, _p(dW), _p(db), _stream()), "mmf_dense_bwd")


def snn_mlp_supported(x, layers) -> bool:
    """Shapes the fused SNN MLP covers: 1-4 blocks, hidden widths <= 1024, a 2-D input."""
    return x.dim() == 2 and 1 <= len(layers) <= 4 and all(W.shape[0] <= 1024 for W, *_ in layers)


def _snn_layers(layers, ys):
    arr = (_lib.SnnLayer * len(layers))()
    for i, ((W, b, keep, p_), y) in enumerate(zip(layers, ys)):
        arr[i] = _lib.SnnLayer(_p(W), _p(b), _p(keep), float(p_), _p(y), W.shape[0])
    return arr


def snn_mlp_fwd(x, layers):
    """SNN_Block x n in ONE launch (models/model_modules.py:64-68). layers: [(W, b, keep mask [B, width] or None, p)].
    Returns (out [B, last width], [y_l]: the pre-dropout SELU outputs the backward needs)."""
    _require_cuda(x)
    x = _f32c(x)
    layers = [(_f32c(W), _f32c(b), None if k is None else _f32c(k), p_) for W, b, k, p_ in layers]
    B = x.shape[0]
    ys = [torch.empty(B, W.shape[0], dtype=torch.float32, device=x.device) for W, *_ in layers]
    out = torch.empty_like(ys[-1])
    check(lib().mmf_snn_mlp_fwd(_p(x), B, x.shape[1], _snn_layers(layers, ys), len(layers), _p(out), _stream()),
          "mmf_snn_mlp_fwd")
    return out, ys


def snn_mlp_bwd(x, layers, ys, dout, need_dx: bool):
    """Returns (dx or None, [(dW, db)])."""
    x, dout = _f32c(x), _f32c(dout)
    layers = [(_f32c(W), _f32c(b), None if k is None else _f32c(k), p_) for W, b, k, p_ in layers]
    B = x.shape[0]
    grads = [(torch.empty_like(W), torch.empty_like(b)) for W, b, *_ in layers]
    dx = torch.empty_like(x) if need_dx else None
    ws = torch.empty(B * sum(W.shape[0] for W, *_ in layers), dtype=torch.float32, device=x.device)
    check(lib().mmf_snn_mlp_bwd(_p(x), B, x.shape[1], _snn_layers(layers, ys), len(layers), _p(dout),
                                _lib.ptr_array([g[0].data_ptr() for g in grads]), _lib.ptr_array([g[1].data_ptr() for g in grads]),
                                0, _p(dx), _p(ws), ws.numel() * 4, _stream()), "mmf_snn_mlp_bwd")
    return dx, grads


def _xf_mods(v_list, params):
    arr = (_lib.XfusionMod * len(v_list))()
    for i, (v, (Wh, bh, Wz, bz, Wo, bo)) in enumerate(zip(v_list, params)):
        arr[i] = _lib.XfusionMod(_p(v), _p(Wh), _p(bh), _p(Wz), _p(bz), _p(Wo), _p(bo))
    return arr


def xfusion_gate_supported(v_list, params) -> bool:
    """Shapes the fused gate kernels cover: 2-4 modalities of equal width, dim % 256 == 0, gate width 16."""
    if not 2 <= len(v_list) <= 4:
        return False
    dim = v_list[0].shape[1]
    return (dim % 256 == 0 and all(v.dim() == 2 and v.shape == v_list[0].shape for v in v_list)
            and all(p_[0].shape == (16, dim) and p_[2].shape == (16, dim * len(v_list)) and p_[4].shape == (16, 16) for p_ in params))


def xfusion_gate_fwd(v_list, params, mask: Optional[torch.Tensor]):
    """The gated reduction of every modality of XlinearFusion in ONE launch (models/model_modules.py:156-166).
    v_list: m x [B, dim]; params: m x (Wh, bh, Wz, bz, Wo, bo); mask: [m, B, 16] dropout scale mask or None.
    Returns (h, z [m, B, 16], o [m, B, 17] with the constant column)."""
    v_list = [_f32c(v) for v in v_list]
    params = [tuple(_f32c(t) for t in p_) for p_ in params]
    _require_cuda(*v_list)
    m, (B, dim) = len(v_list), v_list[0].shape
    h = torch.empty(m, B, 16, dtype=torch.float32, device=v_list[0].device)
    z, o = torch.empty_like(h), torch.empty(m, B, 17, dtype=torch.float32, device=h.device)
    mask = None if mask is None else _f32c(mask)
    check(lib().mmf_xfusion_gate_fwd(_xf_mods(v_list, params), m, B, dim, _p(mask), _p(h), _p(z), _p(o), _stream()),
          "mmf_xfusion_gate_fwd")
    return h, z, o


def xfusion_gate_bwd(v_list, params, mask, h, z, o, d_o, need_dv):
    """Gradients of xfusion_gate_fwd: returns (dv list (None where not needed), m x (dWh, dbh, dWz, dbz, dWo, dbo))."""
    v_list = [_f32c(v) for v in v_list]
    params = [tuple(_f32c(t) for t in p_) for p_ in params]
    m, (B, dim) = len(v_list), v_list[0].shape
    d_o = _f32c(d_o)
    grads = [tuple(torch.empty_like(t) for t in p_) for p_ in params]
    dv = [torch.empty_like(v) if nd else None for v, nd in zip(v_list, need_dv)]
    garr = (_lib.XfusionGrads * m)()
    for i in range(m):
        garr[i] = _lib.XfusionGrads(*[_p(t) for t in grads[i]], _p(dv[i]))
    nbytes = lib().mmf_xfusion_gate_bwd_workspace_bytes(m, B, dim)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=d_o.device)
    check(lib().mmf_xfusion_gate_bwd(_xf_mods(v_list, params), m, B, dim, _p(None if mask is None else _f32c(mask)), _p(h),
                                     _p(z), _p(o), _p(d_o), garr, 0, _p(ws), nbytes, _stream()),
          "mmf_xfusion_gate_bwd")
    return dv, grads


def kron_dropout_code(dropout) -> int:
    """The `dropout` argument of mmf_kron_enc_train_*: False / 0 -> 0; True or p == 0.25 -> 1 (2-bit scheme, the reference's
    default rate); any other rate p -> round(p * 65536) in 2..65535 (16-bit scheme)."""
    if dropout is True:
        return 1
    p_ = float(dropout)
    if p_ <= 0.0:
        return 0
    if p_ == 0.25:
        return 1
    if not p_ < 1.0:
        raise ValueError("dropout rate must be below 1")
    return min(max(int(round(p_ * 65536)), 2), 65535)


def kron_enc_fwd(o_list, W, b, dropout=False, seed: int = 0) -> torch.Tensor:
    """relu(W (o_1 x o_2 [x o_3 [x o_4]]) + b); dropout (True = 0.25, or a rate): train-mode Dropout on the (never
    materialised) product, mask from the counter hash of (seed, stream 3, row, column)."""
    o_list = [_f32c(o) for o in o_list]
    _require_cuda(*o_list)
    B, E = o_list[0].shape
    H = W.shape[0]
    out = torch.empty(B, H, dtype=torch.float32, device=W.device)
    arr = _lib.ptr_array([o.data_ptr() for o in o_list])
    nbytes = lib().mmf_kron_enc_fwd_workspace_bytes(len(o_list), E, B, H)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=W.device) if nbytes else None
    check(lib().mmf_kron_enc_train_fwd_ws(arr, len(o_list), E, B, _p(_f32c(W)), _p(_f32c(b)), H, kron_dropout_code(dropout),
                                          int(seed), _p(out), _p(ws), nbytes, _stream()), "mmf_kron_enc_train_fwd_ws")
    return out


def kron_enc_bwd(o_list, W, out, dout, dropout=False, seed: int = 0):
    o_list = [_f32c(o) for o in o_list]
    W, out, dout = _f32c(W), _f32c(out), _f32c(dout)
    B, E = o_list[0].shape
    H = W.shape[0]
    m = len(o_list)
    d_o = [torch.empty_like(o) for o in o_list]
    dW = torch.zeros_like(W)
    db = torch.zeros(H, dtype=torch.float32, device=W.device)
    nbytes = lib().mmf_kron_enc_workspace_bytes(m, E, B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    arr = _lib.ptr_array([o.data_ptr() for o in o_list])
    darr = _lib.ptr_array([o.data_ptr() for o in d_o])
    check(lib().mmf_kron_enc_train_bwd(arr, m, E, B, _p(W), H, kron_dropout_code(dropout), int(seed), _p(out), _p(dout), darr,
                                       _p(dW), _p(db), _p(ws), nbytes, _stream()), "mmf_kron_enc_train_bwd")
    return d_o, dW, db


def hazard_head_fwd(M, Wk, bk):
    _require_cuda(M, Wk)
    M, Wk, bk = _f32c(M), _f32c(Wk), _f32c(bk)
    B, Lin = M.shape
    K = Wk.shape[0]
    haz = torch.empty(B, K, dtype=torch.float32, device=M.device)
    S = torch.empty_like(haz)
    Y = torch.empty(B, 1, dtype=torch.int64, device=M.device)
    check(lib().mmf_hazard_head_fwd(_p(M), B, Lin, _p(Wk), _p(bk), K, _p(haz), _p(S), _p(Y), _stream()),
          "mmf_hazard_head_fwd")
    return haz, S, Y


def hazard_head_bwd(M, Wk, haz, S, d_haz, d_S):
    M, Wk = _f32c(M), _f32c(Wk)
    B, Lin = M.shape
    K = Wk.shape[0]
    dM = torch.empty_like(M)
    dWk = torch.zeros_like(Wk)
    dbk = torch.zeros(K, dtype=torch.float32, device=M.device)
    check(lib().mmf_hazard_head_bwd(_p(M), B, Lin, _p(Wk), K, _p(haz), _p(S),
                                    _p(None if d_haz is None else _f32c(d_haz)),
                                    _p(None if d_S is None else _f32c(d_S)), _p(dM), _p(dWk), _p(dbk),
                                    _stream()), "mmf_hazard_head_bwd")
    return dM, dWk, dbk


def amil_head_nll_step(partials, Wk, bk, Y, c, alpha: float, eps: float = 1e-7, dWk=None, dbk=None):
    """Fused batch-1 step tail: returns dict(M, ml, hazards, S, Y_hat, loss, dM); dWk/dbk accumulate."""
    _require_cuda(partials, Wk)
    n, Lp2 = partials.shape
    L, K = Lp2 - 2, Wk.shape[0]
    dev = partials.device
    out = dict(M=torch.empty(L, dtype=torch.float32, device=dev), ml=torch.empty(2, dtype=torch.float32, device=dev),
               hazards=torch.empty(1, K, dtype=torch.float32, device=dev),
               S=torch.empty(1, K, dtype=torch.float32, device=dev),
               Y_hat=torch.empty(1, 1, dtype=torch.int64, device=dev),
               loss=torch.empty((), dtype=torch.float32, device=dev),
               dM=torch.empty(L, dtype=torch.float32, device=dev))
    check(lib().mmf_amil_head_nll_step(_p(partials), n, L, _p(Wk), _p(bk), K, _p(Y), _p(c), float(alpha), float(eps),
                                       _p(out["M"]), _p(out["ml"]), _p(out["hazards"]), _p(out["S"]),
                                       _p(out["Y_hat"]), _p(out["loss"]), _p(out["dM"]), _p(dWk), _p(dbk),
                                       _stream()), "mmf_amil_head_nll_step")
    return out


class FusedStepBuffers:
    """Caller-owned buffers of the fused batch-1 training step (mmf_amil_fwd_train_head + mmf_amil_bwd_head):
    the head block's outputs, the z / activation workspace, the score vector and the tile partials. Allocate once
    per (bag size, model) and reuse every step — the step itself then allocates nothing and is CUDA-graph capturable.
    Reference loop: utils/core_utils.py:200-247 (model -> nll_surv -> loss / gc -> backward)."""

    def __init__(self, N: int, w: AmilPrepared, flags: int, K: int, device):
        if not 1 <= K <= 8:
            raise ValueError("the folded head supports 1..8 classes")
        f32 = dict(dtype=torch.float32, device=device)
        self.N, self.K, self.flags = N, K, flags
        self.workspace = amil_bwd_workspace(N, w, flags, device)
        self.A_raw = torch.empty(N, **f32)
        self.partials = torch.empty((N + TILE_ROWS - 1) // TILE_ROWS, w.L + 2, **f32)
        self.M, self.ml = torch.empty(w.L, **f32), torch.empty(2, **f32)
        self.hazards, self.S = torch.empty(1, K, **f32), torch.empty(1, K, **f32)
        self.Y_hat = torch.empty(1, 1, dtype=torch.int64, device=device)
        self.loss = torch.empty((), **f32)
        self.dM, self.hs = torch.empty(w.L, **f32), torch.zeros(16, **f32)
        self.Wk_split = torch.empty(16, w.L, dtype=torch.bfloat16, device=device)
        self._wk_version = None

    def pack_head(self, Wk: torch.Tensor) -> None:
        """bf16 hi / lo split of the classifier weight for the z = Wk h side MMA; redone when Wk changed."""
        ver = (Wk.data_ptr(), Wk._version)
        if ver != self._wk_version:
            check(lib().mmf_pack_head_weights(_p(Wk), self.K, Wk.shape[1], _p(self.Wk_split), _stream()),
                  "mmf_pack_head_weights")
            self._wk_version = ver

    def head_struct(self, Wk, bk, Y, c, alpha, eps, loss_scale, dWk, dbk) -> HeadStep:
        return HeadStep(_p(Wk), _p(bk), _p(self.Wk_split), self.K, _p(Y), _p(c), float(alpha), float(eps),
                        float(loss_scale), _p(self.M), _p(self.ml), _p(self.hazards), _p(self.S), _p(self.Y_hat),
                        _p(self.loss), _p(self.dM), _p(self.hs), _p(dWk), _p(dbk))


def amil_fused_step(x: torch.Tensor, w: AmilPrepared, flags: int, seed: int, buf: FusedStepBuffers, Wk, bk, Y, c,
                    alpha: float, grads: dict, dWk=None, dbk=None, eps: float = 1e-7, loss_scale: float = 1.0,
                    zero: Optional[torch.Tensor] = None, repack_head: bool = True, dx: Optional[torch.Tensor] = None):
    """One batch-1 training step of the path / radio AMIL model in THREE launches: fused forward (+ z = Wk h and the
    ReLU mask words), fused gate + hidden backward whose prologue runs the head (combine, classifier, hazards,
    nll_surv, dlogits, dM, dWk, dbk) and whose phase A is head-projected, grouped wgrad GEMM. N <= 65536. Gradients accumulate into `grads` (dW1, db1, dWab, dbab, dwc,
    dbc), dWk, dbk; `zero` (the flat gradient buffer that holds them all) is cleared by the forward first.
    Outputs are left in `buf` (loss, hazards, S, Y_hat, A_raw, M). dx: optional bf16 [N,1024] output, the gradient w.r.t.
    the bag (a layer sits upstream: the radiology models' reduce_dim) — one more pair-GEMM launch."""
    _require_cuda(x, Wk)
    _check_bag(x, flags)
    N = x.shape[0]
    if N != buf.N:
        raise ValueError("FusedStepBuffers were sized for another bag")
    if repack_head:
        buf.pack_head(Wk)
    head = buf.head_struct(Wk, bk, Y, c, alpha, eps, loss_scale, dWk, dbk)
    wst = w.struct()
    check(lib().mmf_amil_fwd_train_head(_p(x), N, x.stride(0), C.byref(wst), w.L, w.D, flags, seed, _p(buf.A_raw),
                                        _p(buf.partials), buf.workspace.data_ptr(), buf.workspace.numel(), _p(zero),
                                        0 if zero is None else zero.numel(), C.byref(head), _stream()),
          "mmf_amil_fwd_train_head")
    g = AmilGrads(_p(grads["dW1"]), _p(grads["db1"]), _p(grads["dWab"]), _p(grads["dbab"]),
                  _p(grads["dwc"]), _p(grads["dbc"]))
    bflags = flags | MMF_STASHED | (MMF_NEED_DX if dx is not None else 0)
    check(lib().mmf_amil_bwd_head(_p(x), N, x.stride(0), C.byref(wst), w.L, w.D, bflags, seed, _p(buf.A_raw),
                                  _p(buf.partials), C.byref(head), None, C.byref(g), _p(dx), buf.workspace.data_ptr(),
                                  buf.workspace.numel(), _stream()), "mmf_amil_bwd_head")
    return buf.loss


def amil_window_step(packed: PackedBags, w: AmilPrepared, flags: int, seed: int, Wk, bk, Y, c, alpha: float, grads: dict,
                     dWk=None, dbk=None, eps: float = 1e-7, loss_scale: float = 1.0, zero: Optional[torch.Tensor] = None,
                     need_dx: bool = False):
    """fwd + nll_surv + bwd of a WINDOW of bags (the `gc` bags between two optimizer steps, utils/core_utils.py:242-247) in
    one launch set: fused forward of every tile of the packed buffer, one head launch (a cluster per bag), gate + hidden
    backward with per-tile bag statistics, grouped weight gradients. The window's summed gradients (each bag's loss scaled
    by `loss_scale` = 1 / gc) accumulate into `grads` / dWk / dbk; `zero` (the flat gradient buffer) is cleared by the
    forward first. Bags of up to 4096 rows run the split-precision fc like the batch-1 step. Returns dict(loss [n],
    hazards, S [n,K], Y_hat [n,1], M [n,L], A_raw [R] (index with packed.row_offsets / sizes))."""
    _require_cuda(packed.x, Wk)
    n, R, dev = len(packed.sizes), packed.x.shape[0], packed.x.device
    if max(packed.sizes) <= PRECISE_FC_MAX_ROWS:
        x = split_bag(packed.x)                 # (packed.x may be an fp32 bag: the output of an upstream layer)
        flags |= MMF_PRECISE_FC
    else:
        x = to_bf16(packed.x)
    _check_bag(x, flags)
    K, L = Wk.shape[0], w.L
    Wk, bk = _f32c(Wk), _f32c(bk)
    Y = Y.detach().reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    c = c.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    if Y.numel() != n or c.numel() != n:
        raise ValueError("one label / censorship flag per bag")
    f32 = dict(dtype=torch.float32, device=dev)
    ws = amil_bwd_workspace(R, w, flags, dev)
    A_raw = torch.empty(R, **f32)
    partials = torch.empty(R // TILE_ROWS, L + 2, **f32)
    out = dict(M=torch.empty(n, L, **f32), ml=torch.empty(n, 2, **f32), hazards=torch.empty(n, K, **f32),
               S=torch.empty(n, K, **f32), Y_hat=torch.empty(n, 1, dtype=torch.int64, device=dev),
               loss=torch.empty(n, **f32), dM=torch.empty(n, L, **f32), A_raw=A_raw)
    wst = w.struct()
    check(lib().mmf_amil_window_fwd_train(_p(x), R, x.stride(0), C.byref(wst), w.L, w.D, flags, seed, _p(packed.tile_valid_even),
                                          _p(A_raw), _p(partials), ws.data_ptr(), ws.numel(), _p(zero),
                                          0 if zero is None else zero.numel(), _stream()), "mmf_amil_window_fwd_train")
    max_tiles = max((s_ + TILE_ROWS - 1) // TILE_ROWS for s_ in packed.sizes)
    check(lib().mmf_amil_window_head_nll_step(_p(partials), _p(packed.seg_tile_offsets), n, max_tiles, L, _p(Wk), _p(bk), K,
                                              _p(Y), _p(c), float(alpha), float(eps), float(loss_scale), _p(out["M"]),
                                              _p(out["ml"]), _p(out["hazards"]), _p(out["S"]), _p(out["Y_hat"]),
                                              _p(out["loss"]), _p(out["dM"]), _p(dWk), _p(dbk), _stream()),
          "mmf_amil_window_head_nll_step")
    g = AmilGrads(_p(grads["dW1"]), _p(grads["db1"]), _p(grads["dWab"]), _p(grads["dbab"]), _p(grads["dwc"]), _p(grads["dbc"]))
    out["dx"] = torch.empty(R, IN_FEATURES, dtype=torch.bfloat16, device=dev) if need_dx else None   # gradient w.r.t. the bag rows
    check(lib().mmf_amil_window_bwd(_p(x), R, x.stride(0), C.byref(wst), w.L, w.D,
                                    flags | MMF_STASHED | (MMF_NEED_DX if need_dx else 0), seed, _p(A_raw),
                                    _p(out["ml"]), _p(out["M"]), _p(out["dM"]), _p(packed.tile_bag), _p(packed.tile_valid_even),
                                    C.byref(g), _p(out["dx"]), ws.data_ptr(), ws.numel(), _stream()), "mmf_amil_window_bwd")
    return out


def nll_surv(hazards, S, Y, c, alpha: float, eps: float = 1e-7):
    """Returns (loss [scalar tensor], d_hazards, d_S)."""
    _require_cuda(hazards, S)
    hazards, S = _f32c(hazards), _f32c(S)
    B, K = hazards.shape
    Y = Y.detach().reshape(-1).to(device=hazards.device, dtype=torch.int64).contiguous()
    c = c.detach().reshape(-1).to(device=hazards.device, dtype=torch.float32).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=hazards.device)
    dh, dS = torch.empty_like(hazards), torch.empty_like(S)
    check(lib().mmf_nll_surv_fwd_bwd(_p(hazards), _p(S), _p(Y), _p(c), B, K, float(alpha), float(eps), _p(loss),
                                     _p(dh), _p(dS), _stream()), "mmf_nll_surv_fwd_bwd")
    return loss, dh, dS


def cox(theta, times, c):
    _require_cuda(theta)
    theta = _f32c(theta).reshape(-1)
    B = theta.numel()
    times = torch.as_tensor(times).detach().reshape(-1).to(device=theta.device, dtype=torch.float32).contiguous()
    c = torch.as_tensor(c).detach().reshape(-1).to(device=theta.device, dtype=torch.float32).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=theta.device)
    dtheta = torch.empty_like(theta)
    check(lib().mmf_cox_fwd_bwd(_p(theta), _p(times), _p(c), B, _p(loss), _p(dtheta), None, 0, _stream()),
          "mmf_cox_fwd_bwd")
    return loss, dtheta


def ranking(risks, times, c, phi: str = "sigmoid", reduction: str = "mean"):
    _require_cuda(risks)
    risks = _f32c(risks).reshape(-1)
    B = risks.numel()
    times = torch.as_tensor(times).detach().reshape(-1).to(device=risks.device, dtype=torch.float32).contiguous()
    c = torch.as_tensor(c).detach().reshape(-1).to(device=risks.device, dtype=torch.float32).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=risks.device)
    dr = torch.empty_like(risks)
    npairs = torch.empty((), dtype=torch.int64, device=risks.device)
    nbytes = lib().mmf_ranking_workspace_bytes(B)
    ws = torch.empty(nbytes + 8, dtype=torch.uint8, device=risks.device)
    off = (-ws.data_ptr()) % 8
    check(lib().mmf_ranking_fwd_bwd(_p(risks), _p(times), _p(c), B, {"sigmoid": 0, "relu": 1}[phi],
                                    {"mean": 0, "sum": 1}[reduction], _p(loss), _p(dr), _p(npairs),
                                    ws.data_ptr() + off, nbytes, _stream()), "mmf_ranking_fwd_bwd")
    return loss, dr, npairs


# ------------------------------------------------------------------------------------------------
# fcnn / Highway fusion heads: BatchNorm1d, Highway mix, ce_loss
# ------------------------------------------------------------------------------------------------
def batchnorm1d_fwd(x, gamma, beta, running_mean, running_var, train: bool, momentum: float, eps: float):
    """Returns (y, save_mean, save_invstd); running stats are updated in place when train."""
    _require_cuda(x)
    x = _f32c(x)
    B, F = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(F, dtype=torch.float32, device=x.device)
    invstd = torch.empty(F, dtype=torch.float32, device=x.device)
    check(lib().mmf_batchnorm1d_fwd(_p(x), B, F, _p(gamma), _p(beta), _p(running_mean), _p(running_var), int(train),
                                    float(momentum), float(eps), _p(y), _p(mean), _p(invstd), _stream()),
          "mmf_batchnorm1d_fwd")
    return y, mean, invstd


def batchnorm1d_bwd(x, dy, gamma, mean, invstd, train: bool, need_dx=True, need_affine=True):
    x, dy = _f32c(x), _f32c(dy)
    B, F = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dgamma = torch.zeros(F, dtype=torch.float32, device=x.device) if need_affine else None
    dbeta = torch.zeros(F, dtype=torch.float32, device=x.device) if need_affine else None
    check(lib().mmf_batchnorm1d_bwd(_p(x), _p(dy), B, F, _p(gamma), _p(mean), _p(invstd), int(train), _p(dx), _p(dgamma),
                                    _p(dbeta), _stream()), "mmf_batchnorm1d_bwd")
    return dx, dgamma, dbeta


def highway_mix_fwd(g, n, l):
    _require_cuda(g)
    g, n, l = _f32c(g), _f32c(n), _f32c(l)
    y = torch.empty_like(g)
    check(lib().mmf_highway_mix_fwd(_p(g), _p(n), _p(l), g.numel(), _p(y), _stream()), "mmf_highway_mix_fwd")
    return y


def highway_mix_bwd(g, n, l, dy):
    dy = _f32c(dy)
    dg, dn, dl = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
    check(lib().mmf_highway_mix_bwd(_p(g), _p(n), _p(l), _p(dy), g.numel(), _p(dg), _p(dn), _p(dl), _stream()),
          "mmf_highway_mix_bwd")
    return dg, dn, dl


def ce_surv(hazards, S, Y, c, alpha: float, eps: float = 1e-7):
    """utils/loss_utils.py:41-56. Returns (loss, d_hazards, d_S)."""
    _require_cuda(hazards, S)
    hazards, S = _f32c(hazards), _f32c(S)
    B, K = hazards.shape
    Y = Y.detach().reshape(-1).to(device=hazards.device, dtype=torch.int64).contiguous()
    c = c.detach().reshape(-1).to(device=hazards.device, dtype=torch.float32).contiguous()
    loss = torch.empty((), dtype=torch.float32, device=hazards.device)
    dh, dS = torch.empty_like(hazards), torch.empty_like(S)
    check(lib().mmf_ce_surv_fwd_bwd(_p(hazards), _p(S), _p(Y), _p(c), B, K, float(alpha), float(eps), _p(loss), _p(dh),
                                    _p(dS), _stream()), "mmf_ce_surv_fwd_bwd")
    return loss, dh, dS
