"""CUDA-graph replay of whole training steps (model -> loss -> backward -> optimizer) for the launch-bound loops.

The reference's small-bag loops (utils/core_utils.py:184-247: one radiology patient of ~100 slices per optimizer step;
the cohort heads of models/coxranking_models_pretrained.py on a few hundred embeddings) are tens of launches of a few
microseconds each: driven from Python they cost ~16 us of host time per launch. A step whose shapes repeat is captured ONCE
into a CUDA graph and replayed with a single launch. Two things an eager loop draws on the host every step have to live on
the device for that to be the same training run:

  * the optimizer step count (Adam's bias corrections)   -> ``StepState`` word 0, read by ``mmf_adam_step_multi_dev``;
  * the dropout seeds of the library's kernels            -> ``StepState`` words 1.., passed as ``MMF_SEED_DEVICE(ptr)``;

``mmf_step_state_advance`` is the first node of every captured graph: the count goes up by one and every seed moves to
the next value of its splitmix64 sequence, so every replay trains with fresh masks. (ATen dropout inside the step uses
torch's own graph-safe Philox state.) No CPU fallback: graphs exist on CUDA devices only.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, Optional

import torch

from ._lib import SEED_DEVICE_BIT, check, lib

# the StepState of the capture in progress (one capture at a time per process: torch captures in global mode; the autograd
# worker threads that run a captured backward see the same state as the capturing thread)
_capturing: Optional["StepState"] = None


class StepState:
    """Device-resident step count + dropout seed words of graph-captured steps (include/mmf_b200.h: mmf_step_state_advance)."""
    MAX_SEEDS = 63

    def __init__(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("StepState lives on a CUDA device (no CPU fallback)")
        # word 0: step count; words 1..: seed words, started from torch's CPU generator (reproducible under manual_seed)
        init = torch.cat([torch.zeros(1, dtype=torch.int64), torch.randint(0, 2 ** 62, (self.MAX_SEEDS,))])
        self.buf = init.to(device)
        self.n_seeds = 0     # seed words handed out by the capture in progress (every capture starts again at word 1:
                             # the graphs that share a StepState never run concurrently)

    @property
    def step_ptr(self) -> int:
        return self.buf.data_ptr()

    def set_step(self, step: int) -> None:
        self.buf[0:1].fill_(int(step))

    def new_seed(self) -> int:
        """The next seed word of this capture; returns the MMF_SEED_DEVICE-encoded address to hand to the kernels in
        place of a host seed. Capture time only."""
        if self.n_seeds >= self.MAX_SEEDS:
            raise RuntimeError("StepState: more than 63 dropout call sites in one captured step")
        self.n_seeds += 1
        return SEED_DEVICE_BIT | (self.buf.data_ptr() + 8 * self.n_seeds)

    def seed_values(self):
        """Current seed words (host copy; tests)."""
        return self.buf[1:1 + self.n_seeds].tolist()

    def advance(self) -> None:
        check(lib().mmf_step_state_advance(self.buf.data_ptr(), self.MAX_SEEDS, torch.cuda.current_stream().cuda_stream),
              "mmf_step_state_advance")


def current_state() -> Optional[StepState]:
    """The StepState of the capture in progress, or None (eager execution)."""
    return _capturing


@contextlib.contextmanager
def device_step_state(state: StepState):
    global _capturing
    prev = _capturing
    _capturing = state
    try:
        yield state
    finally:
        _capturing = prev


def _invalidate_prepared(modules, keep: Optional[list] = None) -> None:
    # the bf16 / packed weight copies are rebuilt INSIDE the capture (every replay follows an optimizer step), and the
    # fused step's per-size buffers are allocated inside it (graph pool); what a capture leaves in those caches is moved
    # into `keep` (alive as long as the graph) and never handed to eager code
    for mod in modules:
        for m in mod.modules():
            if getattr(m, "_mmf_prep", None) is not None:
                if keep is not None:
                    keep.append(m._mmf_prep)
                m._mmf_prep = None
            f = getattr(m, "_fused", None)
            if isinstance(f, dict) and f.get("bufs"):
                if keep is not None:
                    keep.append(f["bufs"])
                f["bufs"] = {}


def _detached(out):
    # outputs are handed out detached: a returned loss that still carries its autograd graph keeps the AccumulateGrad nodes
    # of the EAGER step alive (bound to the eager stream), and the captured backward would then have to synchronise the
    # capturing stream with that stream — which CUDA refuses (cudaErrorStreamCaptureImplicit)
    if isinstance(out, torch.Tensor):
        return out.detach()
    if isinstance(out, (tuple, list)):
        return type(out)(_detached(o) for o in out)
    if isinstance(out, dict):
        return {k: _detached(v) for k, v in out.items()}
    return out


class GraphedStep:
    """``fn()`` = one whole training step on STATIC tensors (inputs the caller overwrites in place between calls; `fn`
    returns the tensors to read back). The first call runs `fn` eagerly (a real step: lazy allocations, kernel
    attributes) and then captures it; every later call is one graph launch.

    `optimizers`: the FusedAdam instances stepped inside `fn` (their step count moves to the device);
    `modules`: the models whose cached weight copies must be rebuilt inside the graph;
    `state` / `pool`: share one StepState / one graph memory pool between the graphs of a family (one per bag size).
    """

    def __init__(self, fn: Callable[[], object], optimizers=(), modules=(), state: Optional[StepState] = None, pool=None):
        self.fn, self.optimizers, self.modules = fn, tuple(optimizers), tuple(modules)
        self.state, self.pool = state, pool
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.outputs = None
        self.replays = 0

    def _capture(self):
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.state is None:
            self.state = StepState(dev)
        for opt in self.optimizers:
            opt.flush_graph_steps()
        steps = {opt.host_step() for opt in self.optimizers}
        if len(steps) > 1:
            raise RuntimeError("GraphedStep: the optimizers of one graph must share their step count")
        if steps:
            self.state.set_step(steps.pop())
        _invalidate_prepared(self.modules)
        self.state.n_seeds = 0
        g = torch.cuda.CUDAGraph()
        kw = {} if self.pool is None else {"pool": self.pool}
        with torch.cuda.graph(g, **kw), device_step_state(self.state):
            self.state.advance()
            self.outputs = _detached(self.fn())
        self.graph = g
        self._keep = []
        _invalidate_prepared(self.modules, self._keep)

    def __call__(self):
        if self.graph is None:
            out = _detached(self.fn())
            self._capture()
            return out
        self.graph.replay()
        self.replays += 1
        for opt in self.optimizers:
            opt.note_graph_step()
        _invalidate_prepared(self.modules)     # the replay changed the parameters behind the version counters
        return self.outputs


class GraphedStepFamily:
    """One GraphedStep per shape key (e.g. the slice count of a radiology patient), sharing a StepState and a memory pool."""

    def __init__(self, max_graphs: int = 256):
        self.max_graphs = max_graphs
        self.graphs: Dict[object, GraphedStep] = {}
        self.state: Optional[StepState] = None
        self.pool = None

    def get(self, key, make: Callable[[], GraphedStep]) -> GraphedStep:
        g = self.graphs.get(key)
        if g is None:
            if len(self.graphs) >= self.max_graphs:
                self.graphs.pop(next(iter(self.graphs)))
            if self.state is None:
                self.state = StepState(torch.device("cuda", torch.cuda.current_device()))
                self.pool = torch.cuda.graph_pool_handle()
            g = self.graphs[key] = make()
            g.state, g.pool = self.state, self.pool
        return g
