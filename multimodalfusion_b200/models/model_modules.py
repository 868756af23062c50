"""Building blocks with the reference's names, constructor signatures and state_dict keys
(reference: models/model_modules.py:64-178), executing on the libmmf_b200 kernels.

The parameter containers are ordinary ``nn.Linear`` modules created in the same order as the
reference creates them, so a seeded construction consumes the RNG identically and checkpoints load
either way.  ``forward`` never calls ``nn.Linear.forward``: it hands the parameters to the fused
kernels through :mod:`multimodalfusion_b200.autograd`.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..autograd import BatchNorm1dFn, Dense, HighwayMix, KronEncoder, KronEncoderTrain, SnnMlp, XfusionGate
from .._lib import ACT_NONE, ACT_RELU, ACT_SELU, ACT_SIGMOID, ACT_TANH


def _seed_from_torch() -> int:
    """Per-call dropout seed drawn from torch's CPU generator (reproducible under manual_seed). While a step is being
    captured into a CUDA graph (multimodalfusion_b200.graphs) the seed is a device word instead, advanced per replay."""
    from ..graphs import current_state
    st = current_state()
    if st is not None:
        return st.new_seed()
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def batchnorm1d_forward(bn: nn.BatchNorm1d, x: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm1d semantics on the library's kernel (batch statistics + running-stat update in training mode,
    running statistics in eval mode); the module only holds the parameters / buffers."""
    train = bn.training or not bn.track_running_stats
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return BatchNorm1dFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, train, momentum, bn.eps)


def fcnn_forward(seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """Linear -> BatchNorm1d -> ReLU -> Dropout(.7) [-> Linear] containers of the early/late-fcnn heads
    (models/coxranking_models_pretrained.py:80-86, models/nll_models_pretrained.py:82-91)."""
    h = Dense.apply(x, seq[0].weight, seq[0].bias, ACT_NONE)
    h = torch.relu(batchnorm1d_forward(seq[1], h))
    h = seq[3](h)                                   # Dropout(0.7): elementwise mask only
    if len(seq) > 4:
        h = Dense.apply(h, seq[4].weight, seq[4].bias, ACT_NONE)
    return h


class Highway(nn.Module):
    """BatchNorm -> Dropout(.7) -> num_layers x [gate * f(nonlinear) + (1 - gate) * linear] -> BatchNorm
    (models/model_modules.py:5-27; same submodule names and creation order). Only f = relu is used by the
    reference heads and is what the fused Dense kernel provides."""

    def __init__(self, size, num_layers, f):
        super().__init__()
        self.num_layers = num_layers
        self.nonlinear = nn.ModuleList([nn.Linear(size, size) for _ in range(num_layers)])
        self.linear = nn.ModuleList([nn.Linear(size, size) for _ in range(num_layers)])
        self.gate = nn.ModuleList([nn.Linear(size, size) for _ in range(num_layers)])
        if f is not F.relu and f is not torch.relu:
            raise NotImplementedError("Highway: only f = relu is on the accelerated path (all the reference's heads use it)")
        self.f = f
        self.bn1 = nn.BatchNorm1d(size)
        self.bn2 = nn.BatchNorm1d(size)
        self.dropout1 = nn.Dropout(0.7)

    def forward(self, x):
        x = batchnorm1d_forward(self.bn1, x.float())
        x = self.dropout1(x)
        for layer in range(self.num_layers):
            gate = Dense.apply(x, self.gate[layer].weight, self.gate[layer].bias, ACT_SIGMOID)
            nonlinear = Dense.apply(x, self.nonlinear[layer].weight, self.nonlinear[layer].bias, ACT_RELU)
            linear = Dense.apply(x, self.linear[layer].weight, self.linear[layer].bias, ACT_NONE)
            x = HighwayMix.apply(gate, nonlinear, linear)
        return batchnorm1d_forward(self.bn2, x)


class ResidualBlock(nn.Module):
    """fc1 -> bn1 -> ReLU -> fc2 -> bn2 -> (+ x) -> ReLU (models/model_modules.py:28-49; same submodule names)."""

    def __init__(self, size):
        super().__init__()
        self.fc1 = nn.Linear(size, size)
        self.bn1 = nn.BatchNorm1d(size)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Linear(size, size)
        self.bn2 = nn.BatchNorm1d(size)

    def forward(self, x):
        x = x.float()
        out = Dense.apply(x, self.fc1.weight, self.fc1.bias, ACT_NONE)
        out = torch.relu(batchnorm1d_forward(self.bn1, out))
        out = Dense.apply(out, self.fc2.weight, self.fc2.bias, ACT_NONE)
        return torch.relu(batchnorm1d_forward(self.bn2, out) + x)


class Residual(nn.Module):
    """n_layer ResidualBlocks (models/model_modules.py:51-59)."""

    def __init__(self, size, n_layer):
        super().__init__()
        self.n_layer = n_layer
        self.blocks = nn.ModuleList([ResidualBlock(size) for _ in range(n_layer)])

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        return x


def SNN_Block(dim1, dim2, dropout=0.25):
    """Linear -> SELU -> AlphaDropout container (models/model_modules.py:64-68); executed by
    :func:`snn_block_forward`."""
    return nn.Sequential(nn.Linear(dim1, dim2), nn.SELU(), nn.AlphaDropout(p=dropout, inplace=False))


def snn_block_forward(block: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    y = Dense.apply(x, block[0].weight, block[0].bias, ACT_SELU)
    if block.training and block[2].p > 0:
        y = F.alpha_dropout(y, block[2].p, True)  # elementwise mask only; the GEMM+SELU is ours
    return y


def snn_forward(blocks, x: torch.Tensor) -> torch.Tensor:
    """A chain of SNN_Blocks (``fc_omic`` of models/model_genomic.py:22-25 and of the multimodal model) on a [B, d] input:
    one fused launch on the GPU (csrc/snn_mlp.cuh; the AlphaDropout keep masks are drawn here, one Bernoulli draw per block),
    block by block otherwise (CPU glue tests, shapes the fused kernel does not cover)."""
    blocks = list(blocks)
    layers = [(b_[0].weight, b_[0].bias) for b_ in blocks]
    if x.is_cuda and x.dim() == 2 and ops.snn_mlp_supported(x, layers):
        ps, keeps = [], []
        for b_ in blocks:
            p_ = float(b_[2].p) if (b_.training and b_[2].p > 0) else 0.0
            ps.append(p_)
            keeps.append(torch.empty(x.shape[0], b_[0].weight.shape[0], dtype=torch.float32, device=x.device).bernoulli_(1 - p_)
                         if p_ > 0 else None)
        return SnnMlp.apply(x.float(), tuple(ps), tuple(keeps), *[t for wb in layers for t in wb])
    for b_ in blocks:
        x = snn_block_forward(b_, x)
    return x


class _AttnBase(nn.Module):
    def amil_weights(self):
        """(Wa, ba, Wb, bb, wc, bc) with Wb = bb = None for the un-gated net."""
        raise NotImplementedError

    def forward(self, x):
        """Stand-alone use (scores for an already computed h): returns (A [N, n_classes], x)."""
        Wa, ba, Wb, bb, wc, bc = self.amil_weights()
        a = Dense.apply(x, Wa, ba, ACT_TANH)
        if self.training and self.use_dropout:
            a = F.dropout(a, 0.25, True)
        if Wb is not None:
            b = Dense.apply(x, Wb, bb, ACT_SIGMOID)
            if self.training and self.use_dropout:
                b = F.dropout(b, 0.25, True)
            a = a * b
        return Dense.apply(a, wc, bc, ACT_NONE), x


class Attn_Net(_AttnBase):
    """Un-gated attention net: Linear(L,D) -> Tanh [-> Dropout] -> Linear(D,n_classes)
    (models/model_modules.py:70-85; keys ``module.0.*`` and ``module.2.*`` / ``module.3.*``)."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        layers = [nn.Linear(L, D), nn.Tanh()]
        if dropout:
            layers.append(nn.Dropout(0.25))
        layers.append(nn.Linear(D, n_classes))
        self.module = nn.Sequential(*layers)
        self.use_dropout = bool(dropout)

    def amil_weights(self):
        first, last = self.module[0], self.module[-1]
        return first.weight, first.bias, None, None, last.weight, last.bias


class Attn_Net_Gated(_AttnBase):
    """Gated attention net: tanh(Wa h) ⊙ sigmoid(Wb h) -> Linear(D,n_classes)
    (models/model_modules.py:87-110; keys ``attention_{a,b}.0.*``, ``attention_c.*``)."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        branch_a = [nn.Linear(L, D), nn.Tanh()]
        branch_b = [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            branch_a.append(nn.Dropout(0.25))
            branch_b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*branch_a)
        self.attention_b = nn.Sequential(*branch_b)
        self.attention_c = nn.Linear(D, n_classes)
        self.use_dropout = bool(dropout)

    def amil_weights(self):
        return (self.attention_a[0].weight, self.attention_a[0].bias, self.attention_b[0].weight,
                self.attention_b[0].bias, self.attention_c.weight, self.attention_c.bias)


class AmilBranch:
    """Mixin-style helper: runs ``nn.Sequential(Linear(1024,L), ReLU, Dropout, Attn_Net*)`` as ONE
    fused kernel and caches the bf16 / packed weight copies until a parameter changes."""

    @staticmethod
    def _params(seq: nn.Sequential):
        fc, attn = seq[0], seq[3]
        Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
        if wc.shape[0] != 1:
            raise NotImplementedError("fused attention-MIL pooling supports n_classes=1 attention heads")
        return (fc.weight, fc.bias, Wa, ba, Wb, bb, wc, bc)

    @staticmethod
    def prepared(seq: nn.Sequential):
        """bf16 / packed copies of the fc + attention weights, rebuilt when a parameter changed (version counters)."""
        params = AmilBranch._params(seq)
        key = tuple((p.data_ptr(), p._version) for p in params if p is not None)
        cache = getattr(seq, "_mmf_prep", None)
        if cache is None or cache[0] != key:
            cache = (key, ops.prepare_amil_weights(*params))
            seq._mmf_prep = cache
        return cache[1]

    # Bags of up to this many instances run on the library's fp32 functor-SGEMM kernels (exact reference arithmetic):
    # below one 128-row MMA tile the tensor cores are mostly idle and the step is launch-latency-bound either way, while
    # a pooled embedding that is exact to fp32 keeps every ReLU unit of a downstream fusion head on the reference's
    # side (a multimodal patient has ~1300 of them; an embedding off by 1e-4 flips one in a few patients, and one
    # flipped unit near the top moves every upstream gradient by a few per cent). Radiology bags of 17-40 slices.
    TINY_BAG_FP32_ROWS = 64

    @staticmethod
    def _pooled_fp32(seq: nn.Sequential, x: torch.Tensor, training: bool):
        """fc -> ReLU -> Dropout -> (gated) attention -> softmax -> A.h with one Dense kernel per layer (fp32)."""
        fc, attn = seq[0], seq[3]
        Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
        h = Dense.apply(x.float(), fc.weight, fc.bias, ACT_RELU)
        if training:
            h = torch.nn.functional.dropout(h, 0.25, True)
        q = Dense.apply(h, Wa, ba, ACT_TANH)
        if training and attn.use_dropout:
            q = torch.nn.functional.dropout(q, 0.25, True)
        if Wb is not None:
            g = Dense.apply(h, Wb, bb, ACT_SIGMOID)
            if training and attn.use_dropout:
                g = torch.nn.functional.dropout(g, 0.25, True)
            q = q * g
        A_raw = Dense.apply(q, wc, bc, ACT_NONE).t()                       # [1, N]
        A = torch.softmax(A_raw, dim=1)
        M = Dense.apply(A, h.t().contiguous(), None, ACT_NONE)             # [1, L] = A . h
        return A_raw, M

    @staticmethod
    def pooled(seq: nn.Sequential, x: torch.Tensor, training: bool, group=None):
        from ..autograd import AmilPool
        attn = seq[3]
        params = AmilBranch._params(seq)
        if group is None and AmilPool.precise_small_bags and 0 < x.shape[0] <= AmilBranch.TINY_BAG_FP32_ROWS:
            return AmilBranch._pooled_fp32(seq, x, training)
        prep = AmilBranch.prepared(seq)
        flags = ops.amil_flags(prep.gated, dropout_h=training, dropout_attn=training and attn.use_dropout)
        seed = _seed_from_torch() if training else 0
        return AmilPool.apply(x, *params, prep, flags, seed, group)


class XlinearFusion(nn.Module):
    """Kronecker ("Xlinear") late fusion (models/model_modules.py:113-178): per-modality gated
    reduction to ``dim//scale_dim`` (+1 constant), outer product across modalities, two encoders
    with an optional skip of the raw embeddings. The outer product is formed inside the
    ``encoder1`` kernel and never materialised in the forward."""

    def __init__(self, skip=1, use_bilinear=0, gate=1, dim=256, scale_dim=16, num_modalities=4,
                 mmhid1=256, mmhid2=256, dropout_rate=0.25):
        super().__init__()
        self.skip, self.use_bilinear, self.gate, self.num_modalities = skip, use_bilinear, gate, num_modalities
        full, small = dim, dim // scale_dim
        skip_dim = full * num_modalities if skip else 0
        blocks = []
        for _ in range(num_modalities):
            lin_h = nn.Sequential(nn.Linear(full, small), nn.ReLU())
            lin_z = (nn.Bilinear(full, full, small) if use_bilinear
                     else nn.Sequential(nn.Linear(full * num_modalities, small)))
            lin_o = nn.Sequential(nn.Linear(small, small), nn.ReLU(), nn.Dropout(p=dropout_rate))
            blocks.append(nn.ModuleList([lin_h, lin_z, lin_o] if gate else [lin_h, lin_o]))
        self.reduce = nn.ModuleList(blocks)
        self.post_fusion_dropout = nn.Dropout(p=dropout_rate)
        self.encoder1 = nn.Sequential(nn.Linear((small + 1) ** num_modalities, mmhid1), nn.ReLU(),
                                      nn.Dropout(p=dropout_rate))
        self.encoder2 = nn.Sequential(nn.Linear(mmhid1 + skip_dim, mmhid2), nn.ReLU(),
                                      nn.Dropout(p=dropout_rate))

    KRON_IN_KERNEL_MAX_ROWS = 128    # train mode: batches up to this size generate the post-fusion dropout mask in-kernel

    def forward(self, v_list: list):
        if self.use_bilinear:
            raise NotImplementedError("use_bilinear=1 (nn.Bilinear gate) is not on the accelerated path")
        if not self.gate:
            # the reference indexes reduce[i][2] unconditionally (models/model_modules.py:163), so
            # gate=0 cannot run there either
            raise NotImplementedError("gate=0 is not runnable in the reference (IndexError on reduce[i][2])")
        if len(v_list) not in (2, 3, 4):
            raise NotImplementedError("Kronecker fusion kernel supports 2, 3 or 4 modalities")
        v_list = [v.float() for v in v_list]
        params = [(blk[0][0].weight, blk[0][0].bias, blk[1][0].weight, blk[1][0].bias, blk[2][0].weight, blk[2][0].bias)
                  for blk in self.reduce[:len(v_list)]]
        if v_list[0].is_cuda and ops.xfusion_gate_supported(v_list, params):
            # every modality's h / z / o chain, the dropout on o and the constant column in ONE launch (was 3 m Dense
            # launches + 4 m ATen launches forward and ~12 m backward); the mask of all modalities is one ATen draw
            p_o = self.reduce[0][2][2].p
            mask = None
            if self.training and p_o > 0:
                mask = torch.empty(len(v_list), v_list[0].shape[0], 16, dtype=torch.float32,
                                   device=v_list[0].device).bernoulli_(1 - p_o).div_(1 - p_o)
            o_all = XfusionGate.apply(len(v_list), mask, *v_list, *[t for p_ in params for t in p_])
            return self._encode(list(o_all.unbind(0)), v_list)
        v_cat = torch.cat(v_list, dim=1)
        o_list = []
        for v, blk in zip(v_list, self.reduce):
            h = Dense.apply(v, blk[0][0].weight, blk[0][0].bias, ACT_RELU)
            z = Dense.apply(v_cat, blk[1][0].weight, blk[1][0].bias, ACT_SIGMOID)
            o = Dense.apply(z * h, blk[2][0].weight, blk[2][0].bias, ACT_RELU)
            o = blk[2][2](o)
            o_list.append(torch.cat([o, torch.ones(o.shape[0], 1, dtype=o.dtype, device=o.device)], dim=1))
        return self._encode(o_list, v_list)

    def _encode(self, o_list, v_list):
        p_f = float(self.post_fusion_dropout.p) if self.training else 0.0
        if p_f > 0 and o_list[0].shape[0] <= self.KRON_IN_KERNEL_MAX_ROWS:
            # patients / small batches: the mask is generated inside the encoder kernels (counter hash; 2-bit fields at the
            # reference's default rate 0.25, 16-bit fields at any other rate, e.g. the cohort heads' 0.7) and the
            # [B, 17^m] product is not materialised — three launches forward + backward instead of ~15
            out = KronEncoderTrain.apply(self.encoder1[0].weight, self.encoder1[0].bias, _seed_from_torch(), p_f, *o_list)
        elif p_f > 0:
            # large cohorts: forming and hashing every product element inside three GEMM operand loaders costs more than
            # the 10 MB product it saves (B = 512, three modalities: 718 vs 598 us of GPU time per cohort step,
            # gpurun_out/r2v / r2w_cfg3_profile.log) — the product is materialised once and masked by ATen
            fused = o_list[0]
            for o in o_list[1:]:
                fused = (fused.unsqueeze(2) * o.unsqueeze(1)).flatten(1)
            fused = self.post_fusion_dropout(fused)
            out = Dense.apply(fused, self.encoder1[0].weight, self.encoder1[0].bias, ACT_RELU)
        else:
            out = KronEncoder.apply(self.encoder1[0].weight, self.encoder1[0].bias, *o_list)
        out = self.encoder1[2](out)
        if self.skip:
            out = torch.cat([out] + v_list, dim=1)
        out = Dense.apply(out, self.encoder2[0].weight, self.encoder2[0].bias, ACT_RELU)
        return self.encoder2[2](out)
