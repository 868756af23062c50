"""torch.autograd.Functions over the C-ABI kernels (the `wrapped as torch.autograd.Functions over a
thin C-ABI extension` layer of the north star).  Forward and backward both run in
``libmmf_b200.so``; autograd only routes gradients between them.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import MMF_NEED_DX, MMF_PRECISE_FC


class AmilPool(torch.autograd.Function):
    """x[N,1024] -> (A_raw[1,N], M[1,L]): fc+ReLU(+dropout) -> (gated) attention scores -> softmax
    pooling, fused (reference: models/model_attention_mil_path.py:52-56).

    With ``group`` set, x is this rank's shard of the bag: the per-rank (m, l, acc) partial is
    all-gathered (L+2 floats per rank) and combined on every rank, so M is replicated; the backward
    then yields this rank's contribution to the weight gradients (sum over ranks = full gradient).

    ``use_stash`` (class switch): True = the training forward stashes h / branch activations (2.5 KB per
    instance) and the backward runs no recompute GEMM; False = the north star's recompute backward.

    ``precise_small_bags`` (class switch): bags of up to ``ops.PRECISE_FC_MAX_ROWS`` instances run the fc layer in split
    precision (x and W1 as bf16 hi + lo pairs, three tensor-core products, ~16 mantissa bits): with plain bf16 operands
    a ReLU unit whose pre-activation sits within bf16 rounding of zero flips against the fp32 reference, and one flip
    moves a whole row of dW1 by ~1/N of its magnitude — visible only on tiny bags (radiology slices, small WSIs),
    exactly where 3x the GEMM1 work costs nothing (the step is latency-bound there).
    """
    use_stash = True
    precise_small_bags = True

    @staticmethod
    def forward(ctx, x, W1, b1, Wa, ba, Wb, bb, wc, bc, prep, flags: int, seed: int, group=None):
        ctx.set_materialize_grads(False)
        # (an instance-sharded bag runs the plain bf16 path on every rank: the choice must not depend on the size of the
        #  local shard, or the ranks of one bag — and the whole-bag run it is compared with — would mix two arithmetics)
        if group is None and AmilPool.precise_small_bags and 0 < x.shape[0] <= ops.PRECISE_FC_MAX_ROWS:
            xb = ops.split_bag(x)
            flags |= MMF_PRECISE_FC
        else:
            xb = ops.to_bf16(x)
        if x.requires_grad:
            flags |= MMF_NEED_DX
        # training (some parameter or x needs a gradient): stash h and the branch activations for the
        # backward instead of recomputing both GEMMs; inference keeps the N x L intermediates on chip
        train = any(ctx.needs_input_grad[:9]) and AmilPool.use_stash
        stash = None
        ctx.empty_shard = group is not None and xb.shape[0] == 0
        if ctx.empty_shard:
            # a rank that owns no rows of a sharded bag contributes the neutral partial (m = -inf, l = 0)
            from .parallel import all_gather_combine, empty_partial
            M, ml = all_gather_combine(empty_partial(prep.L, xb.device), lambda g: ops.amil_combine(g, prep.L, True), group)
            ctx.prep, ctx.flags, ctx.gated = prep, flags, Wb is not None
            return torch.empty(1, 0, device=xb.device), M.view(1, -1)
        if train:
            A_raw, partials, stash = ops.amil_partials_train(xb, prep, flags, seed)
        else:
            A_raw, partials = ops.amil_partials(xb, prep, flags, seed)
        if group is None:
            M, ml = ops.amil_combine(partials, prep.L, True)
        else:
            from .parallel import all_gather_combine
            local = ops.amil_combine(partials, prep.L, False)
            M, ml = all_gather_combine(local, lambda g: ops.amil_combine(g, prep.L, True), group)
        ctx.save_for_backward(xb, A_raw, M, ml)
        ctx.stash = stash
        ctx.prep, ctx.flags, ctx.seed = prep, flags, seed
        ctx.x_dtype = x.dtype
        ctx.gated = Wb is not None
        return A_raw.view(1, -1), M.view(1, -1)

    @staticmethod
    def backward(ctx, dA, dM):
        prep = ctx.prep
        if ctx.empty_shard:
            return (None,) * 13   # no rows here: this rank adds nothing to the SUM all-reduce of the weight grads
        xb, A_raw, M, ml = ctx.saved_tensors
        if dM is None:
            dM = torch.zeros(prep.L, dtype=torch.float32, device=xb.device)
        g = ops.amil_backward(xb, prep, ctx.flags, ctx.seed, A_raw, ml, M, dM, dA, stash=ctx.stash)
        ctx.stash = None
        D = prep.D
        dx = g["dx"].to(ctx.x_dtype) if (ctx.flags & MMF_NEED_DX) else None
        if ctx.gated:
            dWa, dWb, dba, dbb = g["dWab"][:D], g["dWab"][D:], g["dbab"][:D], g["dbab"][D:]
        else:
            dWa, dWb, dba, dbb = g["dWab"], None, g["dbab"], None
        return (dx, g["dW1"], g["db1"], dWa, dba, dWb, dbb, g["dwc"].view(1, D), g["dbc"],
                None, None, None, None)


class HazardHead(torch.autograd.Function):
    """M[B,L] -> (hazards, S, Y_hat)  (models/model_attention_mil_path.py:58-61)."""

    @staticmethod
    def forward(ctx, M, Wk, bk):
        ctx.set_materialize_grads(False)
        haz, S, Y = ops.hazard_head_fwd(M, Wk, bk)
        ctx.save_for_backward(M, Wk, haz, S)
        ctx.mark_non_differentiable(Y)
        return haz, S, Y

    @staticmethod
    def backward(ctx, d_haz, d_S, _dY):
        M, Wk, haz, S = ctx.saved_tensors
        if d_haz is None and d_S is None:
            return None, None, None
        dM, dWk, dbk = ops.hazard_head_bwd(M, Wk, haz, S, d_haz, d_S)
        return dM, dWk, dbk


class Dense(torch.autograd.Function):
    """y = act(x W^T + b), fp32 (SNN blocks, classifier heads, Xlinear reduce / encoder layers)."""

    @staticmethod
    def forward(ctx, x, W, b, act: int):
        y = ops.dense_fwd(x, W, b, act)
        ctx.save_for_backward(x, W, y)
        ctx.act = act
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dx, dW, db = ops.dense_bwd(x, W, ctx.act, y, dy, need_dx=ctx.needs_input_grad[0],
                                   need_dw=ctx.needs_input_grad[1],
                                   need_db=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, dW, db, None


class KronEncoder(torch.autograd.Function):
    """relu(W (o_1 ⊗ o_2 [⊗ o_3 [⊗ o_4]]) + b) without materialising the outer product in the forward
    (models/model_modules.py:167-173)."""

    @staticmethod
    def forward(ctx, W, b, *o_list):
        out = ops.kron_enc_fwd(o_list, W, b)
        ctx.save_for_backward(W, out, *o_list)
        return out

    @staticmethod
    def backward(ctx, dout):
        W, out, *o_list = ctx.saved_tensors
        d_o, dW, db = ops.kron_enc_bwd(o_list, W, out, dout)
        return (dW, db, *d_o)


class KronEncoderTrain(torch.autograd.Function):
    """KronEncoder with train-mode Dropout(p) on the outer product (XlinearFusion.post_fusion_dropout,
    models/model_modules.py:170): the mask comes from the counter hash of (seed, stream 3, row, column) inside the
    kernels — forward and backward — so the [B, 17^m] product is not materialised in training either."""

    @staticmethod
    def forward(ctx, W, b, seed: int, p: float, *o_list):
        out = ops.kron_enc_fwd(o_list, W, b, dropout=p, seed=seed)
        ctx.save_for_backward(W, out, *o_list)
        ctx.seed, ctx.p = seed, p
        return out

    @staticmethod
    def backward(ctx, dout):
        W, out, *o_list = ctx.saved_tensors
        d_o, dW, db = ops.kron_enc_bwd(o_list, W, out, dout, dropout=ctx.p, seed=ctx.seed)
        return (dW, db, None, None, *d_o)


class SnnMlp(torch.autograd.Function):
    """SNN_Block x n (Linear -> SELU -> AlphaDropout, models/model_modules.py:64-68) as one forward launch and one backward
    chain launch + the weight-gradient GEMMs (csrc/snn_mlp.cuh). apply(x, ps, keeps, W_0, b_0, W_1, b_1, ...): ps = tuple of
    the blocks' dropout rates, keeps = tuple of keep masks [B, width] (or None: eval mode)."""

    @staticmethod
    def forward(ctx, x, ps, keeps, *wb):
        layers = [(wb[2 * i], wb[2 * i + 1], keeps[i], ps[i]) for i in range(len(ps))]
        out, ys = ops.snn_mlp_fwd(x, layers)
        ctx.ps, ctx.keeps = ps, keeps
        ctx.save_for_backward(x, *ys, *wb)
        return out

    @staticmethod
    def backward(ctx, dout):
        n = len(ctx.ps)
        x, *rest = ctx.saved_tensors
        ys, wb = rest[:n], rest[n:]
        layers = [(wb[2 * i], wb[2 * i + 1], ctx.keeps[i], ctx.ps[i]) for i in range(n)]
        dx, grads = ops.snn_mlp_bwd(x, layers, ys, dout, ctx.needs_input_grad[0])
        return (dx, None, None, *[g for pair in grads for g in pair])


class XfusionGate(torch.autograd.Function):
    """Per-modality gated reduction of XlinearFusion for ALL modalities in one forward launch and two or three backward
    launches (models/model_modules.py:156-166; csrc/xfusion_gate.cuh). apply(m, mask, v_1..v_m, then (Wh, bh, Wz, bz, Wo, bo)
    per modality) -> o [m, B, 17] (o[i] = cat(dropout(relu(Wo (z * h) + bo)), 1))."""

    @staticmethod
    def forward(ctx, m: int, mask, *tensors):
        v_list = tensors[:m]
        params = [tensors[m + 6 * i: m + 6 * i + 6] for i in range(m)]
        h, z, o = ops.xfusion_gate_fwd(v_list, params, mask)
        ctx.m, ctx.mask = m, mask
        ctx.save_for_backward(h, z, o, *tensors)
        return o

    @staticmethod
    def backward(ctx, d_o):
        m = ctx.m
        h, z, o, *tensors = ctx.saved_tensors
        v_list = tensors[:m]
        params = [tensors[m + 6 * i: m + 6 * i + 6] for i in range(m)]
        dv, grads = ops.xfusion_gate_bwd(v_list, params, ctx.mask, h, z, o, d_o, ctx.needs_input_grad[2:2 + m])
        return (None, None, *dv, *[g for gs in grads for g in gs])


class SegmentedLinearBf16(torch.autograd.Function):
    """y = cat(segs, 1) @ W^T + b on the bf16 tensor-core GEMM, the modality bags are read in place
    (radio reduce_dim: models/model_attention_mil_radio.py:81-82). Output bf16 feeds AmilPool."""

    @staticmethod
    def forward(ctx, W, b, *segs):
        segs_b = [ops.to_bf16(s) for s in segs]
        Wb = ops.to_bf16(W)
        y = ops.linear_bf16(segs_b, Wb, b.detach().float().contiguous(), torch.bfloat16)
        ctx.save_for_backward(*segs_b)
        ctx.w_shape = W.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        segs_b = ctx.saved_tensors
        dW = torch.zeros(ctx.w_shape, dtype=torch.float32, device=dy.device)
        db = torch.zeros(ctx.w_shape[0], dtype=torch.float32, device=dy.device)
        ops.linear_bf16_wgrad(ops.to_bf16(dy), segs_b, dW, db)
        return (dW, db, *([None] * len(segs_b)))


def reduce_dim_forward(weight, bias, bags):
    """``reduce_dim(cat(bags, 1))`` of the radiology models (models/model_attention_mil_radio.py:81-82). Bags of up to
    ``ops.PRECISE_FC_MAX_ROWS`` slices (every real patient: <= 155) run in fp32 on the functor SGEMM kernel — the same
    arithmetic as the reference, so the bag that enters the split-precision fc is not rounded to bf16 in between;
    larger bags use the segmented bf16 tensor-core GEMM (no concat, bf16 output)."""
    from ._lib import ACT_NONE
    if AmilPool.precise_small_bags and bags[0].shape[0] <= ops.PRECISE_FC_MAX_ROWS:
        return Dense.apply(torch.cat([b.float() for b in bags], dim=1), weight, bias, ACT_NONE)
    return SegmentedLinearBf16.apply(weight, bias, *bags)


class NllSurv(torch.autograd.Function):
    """utils/loss_utils.py:22-39; loss and both input gradients come from one kernel."""

    @staticmethod
    def forward(ctx, hazards, S, Y, c, alpha: float, eps: float):
        loss, dh, dS = ops.nll_surv(hazards, S, Y, c, alpha, eps)
        ctx.save_for_backward(dh, dS)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dh, dS = ctx.saved_tensors
        return dh * dl, dS * dl, None, None, None, None


class CoxLoss(torch.autograd.Function):
    """utils/loss_utils.py:124-139 (sort + scan instead of the O(B^2) host loop)."""

    @staticmethod
    def forward(ctx, theta, times, c):
        loss, dtheta = ops.cox(theta, times, c)
        ctx.save_for_backward(dtheta)
        ctx.shape = theta.shape
        return loss

    @staticmethod
    def backward(ctx, dl):
        (dtheta,) = ctx.saved_tensors
        return (dtheta * dl).view(ctx.shape), None, None


class RankingLoss(torch.autograd.Function):
    """utils/loss_utils.py:58-101."""

    @staticmethod
    def forward(ctx, risks, times, c, phi: str, reduction: str):
        loss, dr, _ = ops.ranking(risks, times, c, phi, reduction)
        ctx.save_for_backward(dr)
        ctx.shape = risks.shape
        return loss

    @staticmethod
    def backward(ctx, dl):
        (dr,) = ctx.saved_tensors
        return (dr * dl).view(ctx.shape), None, None, None, None


class BatchNorm1dFn(torch.autograd.Function):
    """nn.BatchNorm1d on [B,F] (fcnn / Highway fusion heads). running_mean / running_var are updated in place in
    training mode, exactly as the module does."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, train: bool, momentum: float, eps: float):
        y, mean, invstd = ops.batchnorm1d_fwd(x, gamma, beta, running_mean, running_var, train, momentum, eps)
        ctx.save_for_backward(x, gamma, mean, invstd)
        ctx.train = train
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, invstd = ctx.saved_tensors
        dx, dgamma, dbeta = ops.batchnorm1d_bwd(x, dy, gamma, mean, invstd, ctx.train,
                                                need_dx=ctx.needs_input_grad[0],
                                                need_affine=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        return dx, dgamma, dbeta, None, None, None, None, None


class HighwayMix(torch.autograd.Function):
    """y = gate * nonlinear + (1 - gate) * linear (models/model_modules.py:21-25)."""

    @staticmethod
    def forward(ctx, g, n, l):
        y = ops.highway_mix_fwd(g, n, l)
        ctx.save_for_backward(g, n, l)
        return y

    @staticmethod
    def backward(ctx, dy):
        g, n, l = ctx.saved_tensors
        return ops.highway_mix_bwd(g.contiguous(), n.contiguous(), l.contiguous(), dy)


class CeSurv(torch.autograd.Function):
    """utils/loss_utils.py:41-56."""

    @staticmethod
    def forward(ctx, hazards, S, Y, c, alpha: float, eps: float):
        loss, dh, dS = ops.ce_surv(hazards, S, Y, c, alpha, eps)
        ctx.save_for_backward(dh, dS)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dh, dS = ctx.saved_tensors
        return dh * dl, dS * dl, None, None, None, None
