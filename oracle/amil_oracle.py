"""CPU oracle for the attention-MIL -> fusion -> survival-head path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multimodalfusion_b200/`` imports this file; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the checker and the timed CPU baseline, never as a product path.

It is a plain PyTorch fp32 (optionally fp64) restatement of what the reference computes with
stock ATen ops; every function cites the reference lines it follows
(paths relative to the MultimodalFusion/multimodalfusion tree).

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against *outputs of the reference itself*: ``oracle/make_goldens.py`` imports the
reference modules in the build container, runs them on seeded inputs and commits the results as
``tests/golden/*.pt``; ``tests/test_oracle.py`` checks this restatement against those fixtures
(and, when /root/reference is mounted, against the live reference modules).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

TILE = 128


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back (what the kernels see for x, W1, Wa, Wb, h)."""
    return t.to(torch.bfloat16).to(t.dtype)


# ------------------------------------------------------------------------------------------------
# dropout bits — bit-exact restatement of drop_row_state / drop_bits4 (csrc/amil_tile.cuh)
# ------------------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def dropout_scale_mask(seed: int, stream: int, rows: int, cols: int) -> torch.Tensor:
    """[rows, cols] tensor of {0, 1/0.75}: the inverted-dropout (p = 0.25) factor the kernels apply.
    stream 0 = h, 1 = tanh branch, 2 = sigmoid branch. One 32-bit hash per (row, 16-column group),
    2 bits per column, dropped iff the field is 0 (drop_row_state / drop_bits16 / drop_keep)."""
    m32 = np.uint64(0xFFFFFFFF)
    s0 = _mix32(np.array([(seed & 0xFFFFFFFF) ^ ((stream * 0x9E3779B9) & 0xFFFFFFFF)], dtype=np.uint64))
    r = np.arange(rows, dtype=np.uint64)
    row_state = _mix32((s0 + r * np.uint64(0x85EBCA6B) + np.uint64((seed >> 32) & 0xFFFFFFFF)) & m32)
    cg = np.arange((cols + 15) // 16, dtype=np.uint64)
    bits = _mix32(row_state[:, None] ^ ((cg[None, :] * np.uint64(0xC2B2AE35)) & m32))
    j = np.arange(16, dtype=np.uint64)
    fields = (bits[:, :, None] >> (np.uint64(2) * j[None, None, :])) & np.uint64(3)
    keep = (fields != 0).reshape(rows, -1)[:, :cols]
    return torch.from_numpy(keep.astype(np.float32) / 0.75)


def dropout_scale_mask_p(seed: int, stream: int, rows: int, cols: int, p: float) -> torch.Tensor:
    """[rows, cols] inverted-dropout factor at ANY rate p (csrc/small_kernels.cuh KronElem::keep_scale, drop > 1): one
    32-bit hash per (row, column pair), 16 bits per column, dropped iff the field is below t = round(p * 65536); the
    kept elements are scaled by 65536 / (65536 - t)."""
    t = min(max(int(round(p * 65536)), 2), 65535)
    m32 = np.uint64(0xFFFFFFFF)
    s0 = _mix32(np.array([(seed & 0xFFFFFFFF) ^ ((stream * 0x9E3779B9) & 0xFFFFFFFF)], dtype=np.uint64))
    r = np.arange(rows, dtype=np.uint64)
    row_state = _mix32((s0 + r * np.uint64(0x85EBCA6B) + np.uint64((seed >> 32) & 0xFFFFFFFF)) & m32)
    cp = np.arange((cols + 1) // 2, dtype=np.uint64)
    w = _mix32(row_state[:, None] ^ ((cp[None, :] * np.uint64(0xC2B2AE35)) & m32))
    j = np.arange(2, dtype=np.uint64)
    fields = (w[:, :, None] >> (np.uint64(16) * j[None, None, :])) & np.uint64(0xFFFF)
    keep = (fields >= np.uint64(t)).reshape(rows, -1)[:, :cols]
    return torch.from_numpy(keep.astype(np.float32) * np.float32(65536.0 / (65536.0 - t)))


# ------------------------------------------------------------------------------------------------
# attention-MIL forward   (models/model_attention_mil_path.py:20-21,29,50-61;
#                          models/model_modules.py:84-85 un-gated, :105-110 gated)
# ------------------------------------------------------------------------------------------------
def fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc, *, h_scale=None, a_scale=None, g_scale=None,
                 round_h: bool = False):
    """h = relu(x W1^T + b1) [* h_scale]; s = wc · (tanh(Wa h + ba) ⊙ sigmoid(Wb h + bb)) + bc.
    Wb None -> un-gated. *_scale are inverted-dropout factors ({0, 1/(1-p)} tensors; None in eval).
    Returns (s [N], h, a, g): a, g are the branch activations BEFORE their dropout scaling."""
    h = torch.relu(x @ W1.t() + b1)
    if h_scale is not None:
        h = h * h_scale
    if round_h:
        h = bf16_round(h)
    a = torch.tanh(h @ Wa.t() + ba)
    g = torch.sigmoid(h @ Wb.t() + bb) if Wb is not None else torch.ones_like(a)
    ad = a * a_scale if a_scale is not None else a
    gd = g * g_scale if (g_scale is not None and Wb is not None) else g
    s = (ad * gd) @ wc.reshape(-1) + bc.reshape(())
    return s, h, a, g


def softmax_pool(s: torch.Tensor, h: torch.Tensor):
    """A = softmax(s over the bag); M = A h   (models/model_attention_mil_path.py:53-56)."""
    m = s.max()
    e = torch.exp(s - m)
    l = e.sum()
    return (e / l) @ h, m, l


def tile_partials(s: torch.Tensor, h: torch.Tensor, tile: int = TILE) -> torch.Tensor:
    """Per-tile online-softmax partial rows (m_t, l_t, acc_t[L]) exactly as the tile kernel emits."""
    rows = []
    for r0 in range(0, s.numel(), tile):
        st, ht = s[r0:r0 + tile], h[r0:r0 + tile]
        m = st.max()
        e = torch.exp(st - m)
        rows.append(torch.cat([m.reshape(1), e.sum().reshape(1), e @ ht]))
    return torch.stack(rows)


def combine_partials(parts: torch.Tensor, normalize: bool = True):
    """Associative combine of (m, l, acc) rows."""
    m = parts[:, 0].max()
    w = torch.exp(parts[:, 0] - m)
    l = (parts[:, 1] * w).sum()
    acc = (parts[:, 2:] * w[:, None]).sum(0)
    if normalize:
        return acc / l, m, l
    return torch.cat([m.reshape(1), l.reshape(1), acc])


def hazard_head(M, Wk, bk):
    """logits -> hazards = sigmoid, S = cumprod(1 - hazards), Y_hat = argmax
    (models/model_attention_mil_path.py:58-61)."""
    logits = M @ Wk.t() + bk
    hazards = torch.sigmoid(logits)
    S = torch.cumprod(1 - hazards, dim=1)
    return hazards, S, logits.argmax(dim=1, keepdim=True)


# ------------------------------------------------------------------------------------------------
# analytic backward — the spec of the backward kernels
# ------------------------------------------------------------------------------------------------
def amil_backward(x, W1, Wa, Wb, wc, s, h, a, g, M, m, l, dM, dA_raw=None, *, drop_h=False,
                  a_scale=None, g_scale=None, need_dx=False) -> Dict[str, torch.Tensor]:
    """Gradients of (A_raw, M) w.r.t. the fc / attention parameters given dM [L] and dA_raw [N].
    a, g: unscaled branch activations from fc_attention; h already carries its dropout scaling."""
    p = torch.exp(s - m) / l
    ds = p * (h @ dM - dM @ M)
    if dA_raw is not None:
        ds = ds + dA_raw
    return _amil_backward_from_ds(x, W1, Wa, Wb, wc, h, a, g, p, ds, dM, (h > 0).to(h.dtype), drop_h=drop_h,
                                  a_scale=a_scale, g_scale=g_scale, need_dx=need_dx)


def _amil_backward_from_ds(x, W1, Wa, Wb, wc, h, a, g, p, ds, dM, relu_gate, *, drop_h=False, a_scale=None,
                           g_scale=None, need_dx=False) -> Dict[str, torch.Tensor]:
    """Everything downstream of the score gradient ds [N]: gate backward, weight gradients of the
    attention branches, dh -> ReLU / dropout gate -> weight gradients of the fc layer."""
    ka = a_scale if a_scale is not None else torch.ones_like(a)
    kg = g_scale if (g_scale is not None and Wb is not None) else torch.ones_like(g)
    ad, gd = a * ka, g * kg
    dq = ds[:, None] * wc.reshape(1, -1)
    out = {"dwc": (ds[:, None] * ad * gd).sum(0), "dbc": ds.sum().reshape(1)}
    da = dq * gd * ka * (1 - a * a)
    if Wb is not None:
        dg = dq * ad * kg * g * (1 - g)
        dG = torch.cat([da, dg], 1)
        Wab = torch.cat([Wa, Wb], 0)
    else:
        dG, Wab = da, Wa
    out["dWab"] = dG.t() @ h
    out["dbab"] = dG.sum(0)
    dh = p[:, None] * dM[None, :] + dG @ Wab
    du = dh * relu_gate * ((1.0 / 0.75) if drop_h else 1.0)
    out["dW1"] = du.t() @ x
    out["db1"] = du.sum(0)
    if need_dx:
        out["dx"] = du @ W1
    return out


def head_projection(h, Wk):
    """z_i = Wk h_i [N, K] and the ReLU mask [N, L] — what a training forward can emit while the H tile is resident
    (head-projected phase A; K = n_classes = 4-8 floats per instance)."""
    return h @ Wk.t(), h > 0


def amil_backward_head_projected(x, W1, Wa, Wb, wc, s, h, a, g, M, m, l, dlogits, Wk, z, relu_mask, dA_raw=None,
                                 **kw) -> Dict[str, torch.Tensor]:
    """amil_backward for the case where the pooled embedding feeds a LINEAR classifier directly
    (models/model_attention_mil_path.py:56-58: logits = M Wk^T + bk), restated the way the round-2 kernel is meant to
    compute it: dM = Wk^T dlogits, hence t_i = dM.h_i = dlogits.(Wk h_i) = dlogits.z_i — K FMAs per instance from the
    forward's z instead of an L-long dot product over a re-read of the H tile — and the ReLU gate is the forward's
    stored mask. h still enters the weight-gradient GEMM (dWab = dG^T h) as before. Equal to amil_backward up to fp32
    re-association (tests/test_oracle.py::test_head_projected_backward_equals_the_general_one)."""
    dM = Wk.t() @ dlogits
    p = torch.exp(s - m) / l
    ds = p * (z @ dlogits - dM @ M)
    if dA_raw is not None:
        ds = ds + dA_raw
    return _amil_backward_from_ds(x, W1, Wa, Wb, wc, h, a, g, p, ds, dM, relu_mask.to(h.dtype), **kw)


# ------------------------------------------------------------------------------------------------
# survival losses
# ------------------------------------------------------------------------------------------------
def nll_surv_loss(hazards, S, Y, c, alpha=0.4, eps=1e-7):
    """Discrete-time survival NLL (utils/loss_utils.py:22-39)."""
    B = Y.numel()
    Y = Y.reshape(B, 1).long()
    c = c.reshape(B, 1).to(hazards.dtype)
    S_pad = torch.cat([torch.ones_like(c), S], dim=1)
    s_prev = S_pad.gather(1, Y).clamp(min=eps)
    h_y = hazards.gather(1, Y).clamp(min=eps)
    s_y = S_pad.gather(1, Y + 1).clamp(min=eps)
    unc = -(1 - c) * (s_prev.log() + h_y.log())
    cen = -c * s_y.log()
    return ((1 - alpha) * (cen + unc) + alpha * unc).mean()


def cox_loss(theta, times, c):
    """Cox partial likelihood with the reference's risk-set definition t_j >= t_i and mean over all B
    (utils/loss_utils.py:124-139), evaluated with a vectorised mask instead of the host double loop."""
    theta = theta.reshape(-1)
    times = torch.as_tensor(times, dtype=theta.dtype).reshape(-1)
    c = torch.as_tensor(c, dtype=theta.dtype).reshape(-1)
    R = (times[None, :] >= times[:, None]).to(theta.dtype)
    return -torch.mean((theta - torch.log((torch.exp(theta)[None, :] * R).sum(1))) * (1 - c))


def cox_loss_sorted(theta, times, c):
    """Same value via sort + tie-aware prefix sums — the algorithm of the CUDA kernel (App. A.4)."""
    theta = theta.reshape(-1).double()
    times = torch.as_tensor(times).reshape(-1).double()
    c = torch.as_tensor(c).reshape(-1).double()
    order = torch.argsort(times, descending=True, stable=True)
    ts, th, cs = times[order], theta[order], c[order]
    mx = th.max()
    pre = torch.cumsum(torch.exp(th - mx), 0)
    B = ts.numel()
    end = torch.arange(B)
    for k in range(B - 2, -1, -1):
        if ts[k + 1] == ts[k]:
            end[k] = end[k + 1]
    E = pre[end]
    return (-((1 - cs) * (th - mx - torch.log(E))).sum() / B).to(torch.float32)


def ranking_loss(risks, times, c, phi="sigmoid", reduction="mean"):
    """Pairwise ranking loss over comparable pairs (utils/loss_utils.py:58-101), vectorised."""
    risks = risks.reshape(-1)
    if risks.numel() == 1:
        raise NotImplementedError("Batch size must be at least 2")
    times = torch.as_tensor(times, dtype=risks.dtype).reshape(-1)
    ev = (1 - torch.as_tensor(c, dtype=risks.dtype).reshape(-1)) != 0
    comparable = (times[:, None] < times[None, :]) & ev[:, None]   # (i risky, j safe)
    n = int(comparable.sum())
    if n == 0:
        return torch.zeros(1)
    r = risks[:, None] - risks[None, :]
    f = torch.sigmoid(r) if phi == "sigmoid" else torch.relu(r)
    tot = (f * comparable.to(f.dtype)).sum()
    return -(tot / n) if reduction == "mean" else -tot


# ------------------------------------------------------------------------------------------------
# SNN / Kronecker fusion restatements
# ------------------------------------------------------------------------------------------------
def snn_forward(x, layers: List[Tuple[torch.Tensor, torch.Tensor]]):
    """Stack of Linear + SELU (eval mode: AlphaDropout is identity)
    (models/model_modules.py:64-68, models/model_genomic.py:22-25)."""
    for W, b in layers:
        x = torch.selu(x @ W.t() + b)
    return x


def xfusion_gate(v_list, reduce_params, o_scale=None):
    """The per-modality gated reduction of XlinearFusion (models/model_modules.py:156-166), restated for the fused gate
    kernels (csrc/xfusion_gate.cuh): returns o [m, B, S + 1]. reduce_params[i] = ((Wh,bh),(Wz,bz),(Wo,bo));
    o_scale: optional [m, B, S] inverted-dropout scale mask of reduce[i][2][2] (nn.Dropout on o)."""
    v_cat = torch.cat(v_list, dim=1)
    outs = []
    for i, (v, ((Wh, bh), (Wz, bz), (Wo, bo))) in enumerate(zip(v_list, reduce_params)):
        h = torch.relu(v @ Wh.t() + bh)
        z = torch.sigmoid(v_cat @ Wz.t() + bz)
        o = torch.relu((z * h) @ Wo.t() + bo)
        if o_scale is not None:
            o = o * o_scale[i]
        outs.append(torch.cat([o, torch.ones(o.shape[0], 1, dtype=o.dtype, device=o.device)], dim=1))
    return torch.stack(outs)


def xfusion_forward(v_list, reduce_params, enc1, enc2, skip=True, fused_scale=None):
    """XlinearFusion.forward in eval mode, gate=1, use_bilinear=0 (models/model_modules.py:156-178).
    reduce_params[i] = ((Wh,bh),(Wz,bz),(Wo,bo)); enc1/enc2 = (W,b). fused_scale: optional [B, E^m] inverted-dropout
    scale mask of post_fusion_dropout (models/model_modules.py:170; dropout_scale_mask(seed, 3, B, E^m))."""
    o_list = list(xfusion_gate(v_list, reduce_params).unbind(0))     # (the gate restatement is pinned by the same goldens)
    fused = o_list[0]
    for o in o_list[1:]:
        fused = (fused[:, :, None] * o[:, None, :]).flatten(1)
    if fused_scale is not None:
        fused = fused * fused_scale
    out = torch.relu(fused @ enc1[0].t() + enc1[1])
    if skip:
        out = torch.cat([out] + list(v_list), dim=1)
    return torch.relu(out @ enc2[0].t() + enc2[1])


# ------------------------------------------------------------------------------------------------
# fcnn / Highway fusion heads, ce_loss — eval-mode restatements
# ------------------------------------------------------------------------------------------------
def batchnorm1d_eval(x, weight, bias, running_mean, running_var, eps=1e-5):
    """nn.BatchNorm1d in eval mode (models/model_modules.py:13-14, fcnn heads)."""
    return (x - running_mean) / torch.sqrt(running_var + eps) * weight + bias


def batchnorm1d_train(x, weight, bias, eps=1e-5):
    """nn.BatchNorm1d in training mode: batch statistics, biased variance. Returns (y, mean, unbiased var)."""
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias, mean, x.var(0, unbiased=True)


def _bn(sd, prefix, x):
    return batchnorm1d_eval(x, sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"],
                            sd[prefix + ".running_var"])


def fcnn_eval(sd, prefix, x, last=True):
    """Linear -> BatchNorm1d -> ReLU -> Dropout (identity) [-> Linear] (models/coxranking_models_pretrained.py:80-86)."""
    h = x @ sd[prefix + ".0.weight"].t() + sd[prefix + ".0.bias"]
    h = torch.relu(_bn(sd, prefix + ".1", h))
    if last and (prefix + ".4.weight") in sd:
        h = h @ sd[prefix + ".4.weight"].t() + sd[prefix + ".4.bias"]
    return h


def highway_eval(sd, prefix, x, num_layers):
    """Highway.forward in eval mode with f = relu (models/model_modules.py:17-27)."""
    x = _bn(sd, prefix + ".bn1", x)
    for i in range(num_layers):
        lin = lambda nm: x @ sd[f"{prefix}.{nm}.{i}.weight"].t() + sd[f"{prefix}.{nm}.{i}.bias"]
        gate = torch.sigmoid(lin("gate"))
        x = gate * torch.relu(lin("nonlinear")) + (1 - gate) * lin("linear")
    return _bn(sd, prefix + ".bn2", x)


def fusion_head2_eval(sd, kind, train_type, mode, n_layers, h_radio, h_path, h_omic):
    """multimodal_pretrained.forward for the fcnn / highway train types in eval mode
    (models/coxranking_models_pretrained.py:134-169, models/nll_models_pretrained.py:140-176).
    Returns the risk (cox) or the hazard logits (nll)."""
    r, p, o = "radio" in mode, "path" in mode, "omic" in mode
    if train_type.startswith("late"):
        if train_type == "late-fcnn":
            outs = {"radio": fcnn_eval(sd, "layer_MRI", h_radio), "path": fcnn_eval(sd, "layer_WSI", h_path),
                    "omic": fcnn_eval(sd, "layer_omic", h_omic)}
            cw, cb = sd["classifier.0.weight"], sd["classifier.0.bias"]
        else:
            outs = {"radio": highway_eval(sd, "highway_radio", h_radio, n_layers),
                    "path": highway_eval(sd, "highway_path", h_path, n_layers),
                    "omic": highway_eval(sd, "highway_omic", h_omic, n_layers)}
            cw, cb = sd["classifier.weight"], sd["classifier.bias"]
        order = (["radio", "path", "omic"] if r and p and o else ["radio", "path"] if r and p
                 else ["radio", "omic"] if r and o else ["omic", "path"])
        MM = torch.cat([outs[k] for k in order], dim=1)
        out = MM @ cw.t() + cb
        return out.squeeze() if kind == "cox" else out
    order = ([h_radio, h_path, h_omic] if r and p and o else [h_radio, h_path] if r and p
             else [h_radio, h_omic] if r and o else [h_omic, h_path])
    MM = torch.cat(order, dim=1)
    if train_type == "early-fcnn":
        return fcnn_eval(sd, "classifier", MM)
    MM = highway_eval(sd, "highway", MM, n_layers)
    return MM @ sd["classifier.weight"].t() + sd["classifier.bias"]


def ce_surv_loss(hazards, S, Y, c, alpha=0.4, eps=1e-7):
    """utils/loss_utils.py:41-56."""
    B = len(Y)
    Y = Y.view(B, 1)
    c = c.view(B, 1).float()
    S_padded = torch.cat([torch.ones_like(c), S], 1)
    reg = -(1 - c) * (torch.log(torch.gather(S_padded, 1, Y) + eps) + torch.log(torch.gather(hazards, 1, Y).clamp(min=eps)))
    sy = torch.gather(S, 1, Y).clamp(min=eps)
    ce = -c * torch.log(sy) - (1 - c) * torch.log(1 - sy)
    return ((1 - alpha) * ce + alpha * reg).mean()


def concordance_index(risk, times, event, tied_tol=1e-8):
    """Harrell's C as sksurv.metrics.concordance_index_censored computes it (utils/core_utils.py:258; scikit-survival
    is not vendored in the reference tree nor installed here — pinned dependency of its env.yml — so this restates its
    published algorithm, metrics.py:_get_comparable / _estimate_concordance_index): comparable pairs = (i has an event,
    t_i < t_j) plus (i has an event, t_j == t_i, j censored); ties in risk within tied_tol count 1/2.
    `sksurv_comparable_pairs` below is a second, literal restatement of _get_comparable (sort + tie groups) used to
    cross-check this one on tied-time data."""
    risk, times, event = map(lambda a: np.asarray(a, dtype=np.float64), (risk, times, event))
    conc = disc = tied = 0
    for i in range(len(times)):
        if not event[i]:
            continue
        mask = (times > times[i]) | ((times == times[i]) & (event == 0))
        d = risk[i] - risk[mask]
        conc += int((d > tied_tol).sum())
        tied += int((np.abs(d) <= tied_tol).sum())
        disc += int((d < -tied_tol).sum())
    tot = conc + disc + tied
    return (conc + 0.5 * tied) / tot if tot else float("nan")


def sksurv_comparable_pairs(times, event):
    """Literal restatement of sksurv.metrics._get_comparable: walk the samples in time order, group equal times;
    an event j of a group is comparable to every later sample and to the censored samples of its own group.
    Returns the set of ordered pairs (j, k)."""
    times = np.asarray(times, dtype=np.float64)
    event = np.asarray(event).astype(bool)
    order = np.argsort(times, kind="stable")
    n = len(times)
    pairs = set()
    i = 0
    while i < n:
        end = i + 1
        while end < n and times[order[end]] == times[order[i]]:
            end += 1
        for j in range(i, end):
            if event[order[j]]:
                for k in range(end, n):
                    pairs.add((int(order[j]), int(order[k])))
                for k in range(i, end):
                    if not event[order[k]]:
                        pairs.add((int(order[j]), int(order[k])))
        i = end
    return pairs
