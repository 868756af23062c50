"""CPU-side audit of the built library's SASS (cuobjdump): no global load may be scheduled ahead of griddepcontrol.wait
in a kernel launched with programmatic dependent launch — nvcc hoists ld.global.nc loads above the wait unless their
address is re-derived after it (pdl_fresh in csrc/mmf_ptx.cuh). Found on the GPU as an intermittent stale read of the
forward's partials by the head kernel (round 2)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_global_load_before_griddepcontrol_wait():
    import sass_pdl_audit
    from multimodalfusion_b200 import _lib
    if not os.path.exists(_lib.DEFAULT_LIB_PATH):
        pytest.skip("library not built")
    kernels, bad = sass_pdl_audit.audit(_lib.DEFAULT_LIB_PATH)
    assert kernels >= 10, "the PDL kernels were not found in the SASS (mnemonic change?)"
    assert not bad, f"global loads ahead of griddepcontrol.wait: {bad[:5]}"
