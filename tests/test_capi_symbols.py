"""CPU: the C-ABI library loads without a GPU and exports every symbol include/mmf_b200.h declares;
the ctypes signature table covers exactly that set; host-only entry points behave."""
import ctypes
import os
import re

import pytest

from multimodalfusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/mmf_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)


def test_host_only_entry_points():
    lib = _lib.lib()
    assert lib.mmf_version() == 1
    assert lib.mmf_error_string(0) == b"ok"
    assert b"workspace" in lib.mmf_error_string(-6)
    assert lib.mmf_amil_num_tiles(1) == 1 and lib.mmf_amil_num_tiles(128) == 1 and lib.mmf_amil_num_tiles(129) == 2
    big = lib.mmf_amil_bwd_workspace_bytes(16384, 512, 384, 1)
    # H [N,L] + dG [N,2D] + dU [N,L] bf16 dominate
    assert big >= 16384 * (512 + 768 + 512) * 2
    assert lib.mmf_amil_bwd_workspace_bytes(0, 512, 384, 1) == 0
    assert lib.mmf_ranking_workspace_bytes(512) == 16 + 512 * 4
    assert lib.mmf_kron_enc_workspace_bytes(3, 17, 512) == 17 ** 3 * 512 * 4


def test_argument_validation_happens_before_any_device_work():
    """NULL / unsupported arguments are rejected on the host (no GPU needed)."""
    lib = _lib.lib()
    w = _lib.AmilWeights()
    assert lib.mmf_amil_fwd(None, 10, 1024, ctypes.byref(w), 256, 256, 1, 0, None, None, None, None) == -1
    assert lib.mmf_amil_combine(None, 1, 256, 1, None, None, None) == -1
    assert lib.mmf_cox_fwd_bwd(None, None, None, 4, None, None, None, 0, None) == -1
    assert lib.mmf_ranking_fwd_bwd(None, None, None, 1, 0, 0, None, None, None, None, 0, None) == -1
    # Kronecker encoder: 2..4 factors; anything else is refused before a launch (pointers are never dereferenced)
    fake = ctypes.c_void_p(16)
    arr = (ctypes.c_void_p * 5)(16, 16, 16, 16, 16)
    for m in (1, 5):
        assert lib.mmf_kron_enc_fwd(arr, m, 17, 4, fake, fake, 256, fake, None) == -1
        assert lib.mmf_kron_enc_bwd(arr, m, 17, 4, fake, 256, fake, fake, arr, fake, fake, fake, 1 << 30, None) == -1
    assert lib.mmf_kron_enc_bwd(arr, 4, 17, 4, fake, 256, fake, fake, arr, fake, fake, fake, 16, None) == -6   # workspace
    assert lib.mmf_kron_enc_workspace_bytes(4, 17, 33) == 17 ** 4 * 33 * 4
    assert lib.mmf_hazard_head_fwd(fake, 2, 256, fake, fake, 17, fake, fake, None, None) == -1               # K <= 16


def test_built_for_sm_100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
