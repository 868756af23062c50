"""ctypes binding of the C-ABI library (include/mmf_b200.h).

The shared object is built in-tree (``multimodalfusion_b200/libmmf_b200.so``) by
:func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``).  There is no CPU or PyTorch
fallback: if the library is missing, or a kernel is asked to run without a CUDA device, the
call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
# MMF_LIB_PATH: an alternative build of the same library (A/B timing of compile-time kernel variants); must be in-tree
DEFAULT_LIB_PATH = os.path.join(_PKG_DIR, "libmmf_b200.so")      # what build() writes
LIB_PATH = os.environ.get("MMF_LIB_PATH") or DEFAULT_LIB_PATH      # what lib() loads
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "mmf_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

# error codes / flags mirrored from the header
MMF_OK = 0
MMF_GATED, MMF_DROPOUT_H, MMF_DROPOUT_ATTN, MMF_NEED_DX, MMF_STASHED, MMF_PRECISE_FC = 1, 2, 4, 8, 16, 32
SEED_DEVICE_BIT = 1 << 63     # MMF_SEED_DEVICE(ptr): the seed argument is the device address of the seed word
ACT_NONE, ACT_RELU, ACT_SELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4


class MmfError(RuntimeError):
    pass


def _sources():
    return sorted(
        os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))
    ) + [HEADER_PATH]


def needs_build() -> bool:
    if not os.path.exists(DEFAULT_LIB_PATH):
        return True
    t = os.path.getmtime(DEFAULT_LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/capi.cu for sm_100a into the in-tree shared object."""
    if not force and not needs_build():
        return DEFAULT_LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise MmfError("nvcc not found: cannot build libmmf_b200.so")
    tmp = DEFAULT_LIB_PATH + ".tmp"
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, os.path.join(_CSRC, "capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise MmfError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, DEFAULT_LIB_PATH)
    if verbose:
        print(r.stderr)
    return DEFAULT_LIB_PATH


class AmilWeights(C.Structure):
    _fields_ = [
        ("W1", C.c_void_p), ("b1", C.c_void_p), ("Wab", C.c_void_p), ("Wab_packed", C.c_void_p),
        ("bab", C.c_void_p), ("wc", C.c_void_p), ("bc", C.c_void_p), ("W1_split", C.c_void_p),
    ]


class AmilGrads(C.Structure):
    _fields_ = [
        ("dW1", C.c_void_p), ("db1", C.c_void_p), ("dWab", C.c_void_p), ("dbab", C.c_void_p),
        ("dwc", C.c_void_p), ("dbc", C.c_void_p),
    ]


class HeadStep(C.Structure):
    """MmfHeadStep (include/mmf_b200.h): the head block of the fused batch-1 training step."""
    _fields_ = [
        ("Wk", C.c_void_p), ("bk", C.c_void_p), ("Wk_split", C.c_void_p), ("K", C.c_int),
        ("Y", C.c_void_p), ("c", C.c_void_p), ("alpha", C.c_float), ("eps", C.c_float), ("loss_scale", C.c_float),
        ("M", C.c_void_p), ("ml", C.c_void_p), ("hazards", C.c_void_p), ("S", C.c_void_p), ("Y_hat", C.c_void_p),
        ("loss", C.c_void_p), ("dM", C.c_void_p), ("hs", C.c_void_p), ("dWk", C.c_void_p), ("dbk", C.c_void_p),
    ]


class XfusionMod(C.Structure):
    """MmfXfusionMod (include/mmf_b200.h): one modality of the fused XlinearFusion gate."""
    _fields_ = [("v", C.c_void_p), ("Wh", C.c_void_p), ("bh", C.c_void_p), ("Wz", C.c_void_p), ("bz", C.c_void_p),
                ("Wo", C.c_void_p), ("bo", C.c_void_p)]


class XfusionGrads(C.Structure):
    _fields_ = [("dWh", C.c_void_p), ("dbh", C.c_void_p), ("dWz", C.c_void_p), ("dbz", C.c_void_p), ("dWo", C.c_void_p),
                ("dbo", C.c_void_p), ("dv", C.c_void_p)]


class SnnLayer(C.Structure):
    """MmfSnnLayer (include/mmf_b200.h): one Linear -> SELU -> AlphaDropout block of the fused SNN MLP."""
    _fields_ = [("W", C.c_void_p), ("b", C.c_void_p), ("keep", C.c_void_p), ("p", C.c_float), ("y", C.c_void_p),
                ("width", C.c_int)]


_vp, _i, _i64, _sz, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_uint64
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the single source of truth the symbol test checks against the header
SIGNATURES = {
    "mmf_version": (_i, []),
    "mmf_debug_set_timing_buffer": (None, [_vp]),
    "mmf_debug_stamps_enabled": (_i, []),
    "mmf_debug_set_timeline_buffer": (None, [_vp]),
    "mmf_debug_set_p2p_stamp_buffer": (None, [_vp]),
    "mmf_error_string": (C.c_char_p, [_i]),
    "mmf_cast_f32_to_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "mmf_pack_wab": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mmf_split_f32_bf16x3": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "mmf_amil_num_tiles": (_i64, [_i64]),
    "mmf_amil_fwd": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp]),
    "mmf_amil_infer_varlen": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmf_amil_combine": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp]),
    "mmf_amil_bwd_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "mmf_amil_fwd_train": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _sz, _vp, _i64,
                                _vp]),
    "mmf_pack_head_weights": (_i, [_vp, _i, _i, _vp, _vp]),
    "mmf_amil_fwd_train_head": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _sz, _vp,
                                     _i64, C.POINTER(HeadStep), _vp]),
    "mmf_amil_bwd_head": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, C.POINTER(HeadStep), _vp,
                               C.POINTER(AmilGrads), _vp, _vp, _sz, _vp]),
    "mmf_amil_bwd_gate_hidden_head": (_i, [_i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, C.POINTER(HeadStep),
                                           _vp, C.POINTER(AmilGrads), _vp, _sz, _vp]),
    "mmf_amil_bwd_gate_hidden_stashed": (_i, [_i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp, _vp,
                                              C.POINTER(AmilGrads), _vp, _sz, _vp]),
    "mmf_amil_bwd": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp,
                          _vp, _vp, C.POINTER(AmilGrads), _vp, _vp, _sz, _vp]),
    "mmf_amil_bwd_gate": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp, _vp,
                               C.POINTER(AmilGrads), _vp, _sz, _vp]),
    "mmf_amil_bwd_hidden": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _vp, _vp, _vp,
                                 C.POINTER(AmilGrads), _vp, _sz, _vp]),
    "mmf_amil_bwd_wgrad": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, C.POINTER(AmilGrads), _vp, _vp,
                                _sz, _vp]),
    "mmf_amil_window_fwd_train": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp, _sz, _vp,
                                       _i64, _vp]),
    "mmf_amil_window_head_nll_step": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp]),
    "mmf_amil_window_bwd": (_i, [_vp, _i64, _i64, C.POINTER(AmilWeights), _i, _i, _i, _u64, _vp, _vp, _vp, _vp, _vp, _vp,
                                 C.POINTER(AmilGrads), _vp, _vp, _sz, _vp]),
    "mmf_linear_bf16": (_i, [_PP, _i, _i64, _i, _i64, _vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "mmf_linear_bf16_wgrad_workspace_bytes": (_sz, [_i64, _i]),
    "mmf_linear_bf16_wgrad": (_i, [_vp, _i64, _i, _i64, _PP, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "mmf_dense_fwd": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp]),
    "mmf_dense_fwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "mmf_dense_fwd_ws": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _sz, _vp]),
    "mmf_dense_bwd": (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _vp, _vp]),
    "mmf_kron_enc_fwd": (_i, [_PP, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "mmf_kron_enc_workspace_bytes": (_sz, [_i, _i, _i]),
    "mmf_kron_enc_bwd": (_i, [_PP, _i, _i, _i, _vp, _i, _vp, _vp, _PP, _vp, _vp, _vp, _sz, _vp]),
    "mmf_kron_enc_train_fwd": (_i, [_PP, _i, _i, _i, _vp, _vp, _i, _i, _u64, _vp, _vp]),
    "mmf_kron_enc_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mmf_kron_enc_train_fwd_ws": (_i, [_PP, _i, _i, _i, _vp, _vp, _i, _i, _u64, _vp, _vp, _sz, _vp]),
    "mmf_kron_enc_train_bwd": (_i, [_PP, _i, _i, _i, _vp, _i, _i, _u64, _vp, _vp, _PP, _vp, _vp, _vp, _sz, _vp]),
    "mmf_hazard_head_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "mmf_hazard_head_bwd": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmf_amil_head_nll_step": (_i, [_vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp, _vp]),
    "mmf_nll_surv_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "mmf_ce_surv_fwd_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "mmf_batchnorm1d_fwd": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "mmf_batchnorm1d_bwd": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "mmf_highway_mix_fwd": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mmf_highway_mix_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mmf_cox_workspace_bytes": (_sz, [_i]),
    "mmf_cox_fwd_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmf_adam_step_multi": (_i, [_PP, _PP, _PP, _PP, C.POINTER(C.c_int64), _i, _i, _f, _f, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "mmf_snn_mlp_fwd": (_i, [_vp, _i, _i, C.POINTER(SnnLayer), _i, _vp, _vp]),
    "mmf_snn_mlp_bwd": (_i, [_vp, _i, _i, C.POINTER(SnnLayer), _i, _vp, _PP, _PP, _i, _vp, _vp, _sz, _vp]),
    "mmf_xfusion_gate_fwd": (_i, [C.POINTER(XfusionMod), _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mmf_xfusion_gate_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "mmf_xfusion_gate_bwd": (_i, [C.POINTER(XfusionMod), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(XfusionGrads), _i,
                                  _vp, _sz, _vp]),
    "mmf_adam_step_multi_dev": (_i, [_PP, _PP, _PP, _PP, C.POINTER(C.c_int64), _i, _vp, _f, _f, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "mmf_step_state_advance": (_i, [_vp, _i, _vp]),
    "mmf_percentile_of_score": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "mmf_cindex_counts": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp]),
    "mmf_p2p_flag_bytes": (_sz, []),
    "mmf_p2p_allreduce_sum_f32": (_i, [_PP, _PP, _vp, _i, _i, _i64, _i, _vp]),
    "mmf_ranking_workspace_bytes": (_sz, [_i]),
    "mmf_ranking_fwd_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded library. Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MmfError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback for the CUDA path)")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def require_debug_stamps() -> None:
    """The phase-stamp tools need a library compiled with -DMMF_DEBUG_STAMPS=1 (the release build keeps no global state)."""
    if not lib().mmf_debug_stamps_enabled():
        raise MmfError("this libmmf_b200.so is a release build: `python tools/ab_variant.py build stamps -DMMF_DEBUG_STAMPS=1` "
                       "and run the tool with MMF_LIB_PATH pointing at the variant")


def check(rc: int, what: str = "") -> None:
    if rc != MMF_OK:
        msg = lib().mmf_error_string(rc).decode()
        raise MmfError(f"{what or 'mmf call'} failed: {msg} (rc={rc})")


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr
