"""Shared implementation of ``fused_step`` for the models whose pooled embedding feeds a linear classifier directly
(path and radiology attention-MIL survival models): model -> nll_surv -> backward of utils/core_utils.py:200-247 as the
library's three-launch step (mmf_amil_fwd_train_head + mmf_amil_bwd_head), gradients written into the parameters' ``.grad``."""
import torch

from .. import ops
from .model_modules import AmilBranch, _seed_from_torch


def enable(model, seq, classifier):
    """ONE flat fp32 gradient buffer for the fc / attention / classifier parameters; every ``.grad`` becomes a view of it
    (the layout the step's kernels accumulate into). Returns the flat buffer."""
    fc, attn = seq[0], seq[3]
    Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
    order = [fc.weight, fc.bias, Wa] + ([Wb] if Wb is not None else []) + [ba] + ([bb] if bb is not None else []) \
        + [wc, bc, classifier.weight, classifier.bias]
    dev = fc.weight.device
    if dev.type != "cuda":
        raise RuntimeError("enable_fused_step needs the model on a CUDA device (no CPU fallback)")
    n = sum(p.numel() for p in order)
    flat = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=dev)
    o = 0
    for p_ in order:
        p_.grad = flat[o:o + p_.numel()].view_as(p_)
        o += p_.numel()
    L, D = fc.weight.shape[0], Wa.shape[0]
    KD = D * (2 if Wb is not None else 1)
    state = dict(flat=flat, L=L, D=D, KD=KD, gated=Wb is not None, bufs={})
    g, o = {}, 0
    for name, cnt, shape in (("dW1", L * 1024, (L, 1024)), ("db1", L, (L,)), ("dWab", KD * L, (KD, L)),
                             ("dbab", KD, (KD,)), ("dwc", D, (D,)), ("dbc", 1, (1,))):
        g[name] = flat[o:o + cnt].view(shape)
        o += cnt
    K = classifier.weight.shape[0]
    state.update(grads=g, dWk=flat[o:o + K * L].view(K, L), dbk=flat[o + K * L:o + K * L + K])
    model._fused = state
    return flat


def run(model, seq, classifier, bag, Y, c, alpha, loss_scale, accumulate, eps, need_dx=False):
    """One step on `bag` ([N,1024], fp32 or bf16). Returns (hazards, S, Y_hat, A_raw [1,N], loss, dx or None)."""
    f = model._fused
    prep = AmilBranch.prepared(seq)
    N = bag.shape[0]
    if N > 65536:
        raise NotImplementedError("fused_step merges at most 512 per-tile head rows per CTA (N <= 65536)")
    attn = seq[3]
    flags = ops.amil_flags(prep.gated, dropout_h=model.training, dropout_attn=model.training and attn.use_dropout)
    if N <= ops.PRECISE_FC_MAX_ROWS:      # small bag: split-precision fc (see autograd.AmilPool)
        x = ops.split_bag(bag)
        flags |= ops.MMF_PRECISE_FC
    else:
        x = ops.to_bf16(bag)
    seed = _seed_from_torch() if model.training else 0
    K = classifier.weight.shape[0]
    buf = f["bufs"].get(N)
    if buf is None:
        if len(f["bufs"]) >= 4:     # bags come in many sizes: keep a few workspaces, not one per size
            f["bufs"].pop(next(iter(f["bufs"])))
        buf = f["bufs"][N] = ops.FusedStepBuffers(N, prep, flags, K, x.device)
    Yd = Y.detach().reshape(-1).to(device=x.device, dtype=torch.int64)
    cd = c.detach().reshape(-1).to(device=x.device, dtype=torch.float32)
    dx = torch.empty(N, 1024, dtype=torch.bfloat16, device=x.device) if need_dx else None
    ops.amil_fused_step(x, prep, flags, seed, buf, classifier.weight.detach(), classifier.bias.detach(),
                        Yd, cd, alpha, f["grads"], dWk=f["dWk"], dbk=f["dbk"], eps=eps, loss_scale=loss_scale,
                        zero=None if accumulate else f["flat"], dx=dx)
    return buf.hazards, buf.S, buf.Y_hat, buf.A_raw.view(1, -1), buf.loss, dx


def graphed(model, optimizer, bags: dict, Y, c, alpha=0.0, loss_scale=1.0, eps=1e-7):
    """``model.fused_step(...)`` + ``optimizer.step()`` of one patient as ONE CUDA-graph launch (multimodalfusion_b200.graphs):
    the batch-1 loop of utils/core_utils.py:184-247 with `--gc 1`. One graph per (bag size, dtype, hyper-parameters), captured
    the first time the size is seen (that patient's step runs eagerly) and replayed afterwards; the bags, Y and c are copied
    into the graph's static inputs (one stack kernel + two small copies). The dropout seeds and Adam's step count live on the
    device and move on with every replay. Returns fused_step's tuple — static buffers, overwritten by the next call."""
    from ..graphs import GraphedStep, GraphedStepFamily
    names = list(bags)
    x0 = bags[names[0]]
    key = (tuple(names), x0.shape[0], x0.dtype, bool(model.training), float(alpha), float(loss_scale), float(eps), id(optimizer))
    fam = getattr(model, "_graph_family", None)
    if fam is None:
        fam = model._graph_family = GraphedStepFamily()
    entry = fam.graphs.get(key)
    if entry is None:
        sx = torch.empty((len(names),) + tuple(x0.shape), dtype=x0.dtype, device=x0.device)
        sY = torch.zeros(1, dtype=torch.int64, device=x0.device)
        sc = torch.zeros(1, dtype=torch.float32, device=x0.device)

        def fn():
            out = model.fused_step(Y=sY, c=sc, alpha=alpha, loss_scale=loss_scale, accumulate=False, eps=eps,
                                   **{n: sx[i] for i, n in enumerate(names)})
            optimizer.step(zero_grad=False)     # (the next step clears the gradients inside its forward kernel)
            return out

        entry = fam.get(key, lambda: GraphedStep(fn, optimizers=(optimizer,), modules=(model,)))
        entry.static = (sx, sY, sc)
    sx, sY, sc = entry.static
    if len(names) == 1:
        sx[0].copy_(x0)
    else:
        torch.stack([bags[n] for n in names], out=sx)
    sY.copy_(Y.reshape(-1), non_blocking=True)
    sc.copy_(c.reshape(-1), non_blocking=True)
    return entry()
