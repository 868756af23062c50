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


@contextlib.contextmanager
def oracle_kernels():
    """Route Dense / KronEncoder / HazardHead / SegmentedLinearBf16 / BatchNorm1dFn / HighwayMix / the fused AMIL pooling
    through fp32 torch ops."""
    saved = [(A.Dense, A.Dense.__dict__.get("apply")), (A.KronEncoder, A.KronEncoder.__dict__.get("apply")),
             (A.HazardHead, A.HazardHead.__dict__.get("apply")),
             (A.SegmentedLinearBf16, A.SegmentedLinearBf16.__dict__.get("apply")),
             (A.BatchNorm1dFn, A.BatchNorm1dFn.__dict__.get("apply")), (A.HighwayMix, A.HighwayMix.__dict__.get("apply"))]
    pooled = MM.AmilBranch.__dict__["pooled"]
    A.Dense.apply = staticmethod(_dense)
    A.KronEncoder.apply = staticmethod(_kron)
    A.HazardHead.apply = staticmethod(_hazard)
    A.SegmentedLinearBf16.apply = staticmethod(_segmented_linear)
    A.BatchNorm1dFn.apply = staticmethod(_batchnorm)
    A.HighwayMix.apply = staticmethod(_highway_mix)
    MM.AmilBranch.pooled = staticmethod(_pooled)
    try:
        yield
    finally:
        for cls, orig in saved:
            if orig is None:
                delattr(cls, "apply")
            else:
                cls.apply = orig
        MM.AmilBranch.pooled = pooled
