"""multimodalfusion_b200 — B200-native (sm_100a) attention-MIL / fusion / survival-head hot path.

Drop-in mirrors of the reference's model and loss modules (same constructors, forward kwargs,
return tuples and state_dict keys) whose arithmetic runs in hand-written CUDA kernels behind a
C ABI (include/mmf_b200.h).  See DESIGN.md and INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from ._lib import MmfError, build, lib  # noqa: F401

__version__ = "0.1.0"
