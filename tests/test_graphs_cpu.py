"""CPU: host-side bookkeeping of the graph-replayed training steps (multimodalfusion_b200/graphs.py, utils/optim.py) —
what does not need a device: the optimizer's lazily synchronised step count, output detaching, the no-CPU-fallback
guards, and the host restatement of the device seed sequence against the header's definition."""
import pytest
import torch

from multimodalfusion_b200 import graphs
from multimodalfusion_b200.utils.optim import FusedAdam


def test_fused_adam_graph_step_bookkeeping():
    p = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
    opt = FusedAdam(p, lr=1e-3)
    assert opt.host_step() == 0
    for q in p:       # what an eager step would have created
        opt.state[q] = {"step": 3, "exp_avg": torch.zeros_like(q), "exp_avg_sq": torch.zeros_like(q)}
    opt.note_graph_step(); opt.note_graph_step()
    assert opt._graph_steps == 2
    assert opt.host_step() == 5 and opt._graph_steps == 0            # flushed into every parameter's entry
    assert all(st["step"] == 5 for st in opt.state.values())
    opt.note_graph_step()
    sd = opt.state_dict()                                             # a checkpoint sees the replayed steps
    assert all(st["step"] == 6 for st in sd["state"].values())


def test_outputs_are_detached_recursively():
    w = torch.ones(2, requires_grad=True)
    out = graphs._detached(((w * 2).sum(), [w + 1, {"k": w * 3}], 7, None))
    assert not out[0].requires_grad and not out[1][0].requires_grad and not out[1][1]["k"].requires_grad
    assert out[2] == 7 and out[3] is None and isinstance(out[1], list) and isinstance(out, tuple)


def test_step_state_lives_on_the_device_only():
    with pytest.raises(RuntimeError):
        graphs.StepState("cpu")
    assert graphs.current_state() is None


def test_capture_context_is_restored_on_error():
    sentinel = object()
    with pytest.raises(ValueError):
        with graphs.device_step_state(sentinel):
            assert graphs.current_state() is sentinel
            raise ValueError("boom")
    assert graphs.current_state() is None


def test_seed_device_encoding_matches_the_header():
    from multimodalfusion_b200._lib import HEADER_PATH, SEED_DEVICE_BIT
    assert SEED_DEVICE_BIT == 0x8000000000000000
    assert "0x8000000000000000ull" in open(HEADER_PATH).read()
