"""CPU stand-ins for the CUDA autograd.Functions, so that the drop-in modules' PYTHON GLUE (module order, kwargs, fusion
order, reshapes, which weights feed which op) can be executed and differentiated on a machine without a GPU and compared
with the reference goldens. Test infrastructure only: every stand-in is the oracle's fp32 restatement of the op the
kernel implements (oracle/amil_oracle.py); nothing in the product package imports this file, and the GPU tests never
use it — they call the real kernels through the C ABI."""
import contextlib

import torch

from multimodalfusion_b200 import autograd as A
from multimodalfusion_b200._lib import ACT_NONE, ACT_RELU, ACT_SELU, ACT_SIGMOID, ACT_TANH
from multimodalfusion_b200.models import model_modules as MM
from oracle import amil_oracle as O

_ACT = {ACT_NONE: lambda t: t, ACT_RELU: torch.relu, ACT_SELU: torch.selu, ACT_SIGMOID: torch.sigmoid, ACT_TANH: torch.tanh}


def _dense(x, W, b, act):
    y = x @ W.t()
    return _ACT[act](y if b is None else y + b)


def _kron(W, b, *o_list):
    fused = o_list[0]
    for o in o_list[1:]:
        fused = (fused[:, :, None] * o[:, None, :]).flatten(1)
    return torch.relu(fused @ W.t() + b)


def _hazard(M, Wk, bk):
    return O.hazard_head(M, Wk, bk)


def _segmented_linear(W, b, *segs):
    return torch.cat([s.float() for s in segs], dim=1) @ W.t() + b


def _batchnorm(x, gamma, beta, running_mean, running_var, train, momentum, eps):
    assert not train, "stand-in covers eval mode"
    return O.batchnorm1d_eval(x, gamma, beta, running_mean, running_var, eps)


def _highway_mix(g, n, l):
    return g * n + (1 - g) * l


def _pooled(seq, x, training, group=None):
    assert not training and group is None, "stand-in covers eval mode, single process"
    fc, attn = seq[0], seq[3]
    s, h, _, _ = O.fc_attention(x.float(), fc.weight, fc.bias, *attn.amil_weights())
    M, _, _ = O.softmax_pool(s, h)
    return s.reshape(1, -1), M.reshape(1, -1)


def _ste_bf16(t):
    """bf16 rounding with a straight-through gradient: the value the tensor cores see, the gradient of the identity."""
    return t + (t.to(torch.bfloat16).float() - t).detach()


def _segmented_linear_bf16(W, b, *segs):
    y = torch.cat([_ste_bf16(s.float()) for s in segs], dim=1) @ _ste_bf16(W).t() + b
    return _ste_bf16(y)                       # the GEMM writes the bf16 bag the fused AMIL kernel consumes


def _pooled_bf16(seq, x, training, group=None):
    """The fused kernel's operand precision: bf16 x / W1 / Wa / Wb, bf16 h tile, fp32 accumulation and biases."""
    assert not training and group is None
    fc, attn = seq[0], seq[3]
    Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
    h = _ste_bf16(torch.relu(_ste_bf16(x.float()) @ _ste_bf16(fc.weight).t() + fc.bias))
    a = torch.tanh(h @ _ste_bf16(Wa).t() + ba)
    g = torch.sigmoid(h @ _ste_bf16(Wb).t() + bb) if Wb is not None else torch.ones_like(a)
    s = (a * g) @ wc.reshape(-1) + bc.reshape(())
    M, _, _ = O.softmax_pool(s, h)
    return s.reshape(1, -1), M.reshape(1, -1)


@contextlib.contextmanager
def oracle_kernels(bf16_operands: bool = False):
    """Route Dense / KronEncoder / HazardHead / SegmentedLinearBf16 / BatchNorm1dFn / HighwayMix / the fused AMIL pooling
    through fp32 torch ops. bf16_operands=True gives the tensor-core GEMMs (fused AMIL, segmented reduce_dim) the operand
    precision of the kernels, so that the GPU tests' tolerances can be dry-run on the CPU."""
    saved = [(A.Dense, A.Dense.__dict__.get("apply")), (A.KronEncoder, A.KronEncoder.__dict__.get("apply")),
             (A.HazardHead, A.HazardHead.__dict__.get("apply")),
             (A.SegmentedLinearBf16, A.SegmentedLinearBf16.__dict__.get("apply")),
             (A.BatchNorm1dFn, A.BatchNorm1dFn.__dict__.get("apply")), (A.HighwayMix, A.HighwayMix.__dict__.get("apply"))]
    pooled = MM.AmilBranch.__dict__["pooled"]
    A.Dense.apply = staticmethod(_dense)
    A.KronEncoder.apply = staticmethod(_kron)
    A.HazardHead.apply = staticmethod(_hazard)
    A.SegmentedLinearBf16.apply = staticmethod(_segmented_linear_bf16 if bf16_operands else _segmented_linear)
    A.BatchNorm1dFn.apply = staticmethod(_batchnorm)
    A.HighwayMix.apply = staticmethod(_highway_mix)
    MM.AmilBranch.pooled = staticmethod(_pooled_bf16 if bf16_operands else _pooled)
    try:
        yield
    finally:
        for cls, orig in saved:
            if orig is None:
                delattr(cls, "apply")
            else:
                cls.apply = orig
        MM.AmilBranch.pooled = pooled
