"""GPU parity of the fused batch-1 training step (mmf_amil_fwd_train_head + mmf_amil_bwd_head: three launches) —
forward with the head folded into its tail, head-projected backward — against (a) the modular kernels it replaces
(mmf_amil_fwd_train -> mmf_amil_head_nll_step -> mmf_amil_bwd) and (b) the oracle (bf16 operands, tight bars; fp32
reference at the metric shape with dropout + stash, north-star bars).
Reference step: utils/core_utils.py:200-247 on models/model_attention_mil_path.py:50-72, utils/loss_utils.py:22-39."""
import math

import pytest
import torch

from helpers import rel_err
from oracle import amil_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu

TOL_FWD_REF, TOL_GRAD_REF = 1e-2, 2e-2
TOL_FWD_TIGHT, TOL_GRAD_TIGHT = 4e-3, 8e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda")


def bfr(t):
    return None if t is None else t.to(torch.bfloat16).float()


def close(a, ref, tol):
    """rel_err < tol; a reference that is exactly zero (a one-instance bag has p = 1, ds = p (t - dM.M) = 0) is
    compared absolutely: t - dM.M is then a rounding residual of two different summation orders."""
    ref = ref.detach().float().cpu()
    if ref.abs().max().item() < 1e-6:
        return a.detach().float().cpu().abs().max().item() < 1e-6
    return rel_err(a, ref) < tol


def _rand(L, D, gated, K, seed):
    g = torch.Generator().manual_seed(seed)
    W1 = torch.randn(L, 1024, generator=g) * math.sqrt(2.0 / (1024 + L))
    b1 = torch.randn(L, generator=g) * 0.05
    Wa = torch.randn(D, L, generator=g) * math.sqrt(2.0 / (L + D))
    ba = torch.randn(D, generator=g) * 0.05
    Wb = torch.randn(D, L, generator=g) * math.sqrt(2.0 / (L + D)) if gated else None
    bb = torch.randn(D, generator=g) * 0.05 if gated else None
    wc = torch.randn(1, D, generator=g) * math.sqrt(2.0 / (D + 1))
    bc = torch.randn(1, generator=g) * 0.05
    Wk = torch.randn(K, L, generator=g) * math.sqrt(2.0 / (L + K))
    bk = torch.randn(K, generator=g) * 0.1
    return (W1, b1, Wa, ba, Wb, bb, wc, bc), Wk, bk


def _grad_bufs(L, D, gated, K, dev):
    KD = (2 if gated else 1) * D
    sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
    flat = torch.randn((sum(sizes) + 3) // 4 * 4, device=dev)   # junk: the forward must clear it
    vs, o = [], 0
    for sz in sizes:
        vs.append(flat[o:o + sz]); o += sz
    grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
    return flat, grads, vs[6].view(K, L), vs[7]


FUSED_SHAPES = [
    # N, L, D, gated, drop flags, K, y, c
    (1, 256, 256, True, 0, 4, 0, 0.0), (127, 256, 256, True, 0, 4, 3, 1.0), (129, 256, 256, True, 2, 4, 2, 0.0),
    (1000, 256, 256, True, 6, 4, 1, 1.0), (300, 512, 384, True, 2, 4, 2, 0.0), (1500, 512, 384, True, 6, 8, 5, 0.0),
    (200, 256, 256, False, 0, 4, 0, 1.0), (300, 512, 384, False, 2, 2, 1, 0.0), (500, 256, 384, True, 2, 1, 0, 0.0),
    (4097, 512, 384, True, 2, 4, 3, 0.0),
]


@pytest.mark.parametrize("N,L,D,gated,drop,K,y,c_val", FUSED_SHAPES)
def test_fused_step_matches_modular_kernels_and_oracle(dev, N, L, D, gated, drop, K, y, c_val):
    from multimodalfusion_b200 import ops
    seed = 0xF00D + N
    W, Wk, bk = _rand(L, D, gated, K, N + L + K)
    W1, b1, Wa, ba, Wb, bb, wc, bc = W
    x = cases.features(N, 700 + N)
    xb = x.to(dev).to(torch.bfloat16)
    prep = ops.prepare_amil_weights(*[None if t is None else t.to(dev) for t in W])
    Wkd, bkd = Wk.to(dev), bk.to(dev)
    flags = ops.amil_flags(gated) | drop
    Y, c = torch.tensor([y], device=dev), torch.tensor([c_val], device=dev)
    alpha, scale = 0.15, 0.5

    # ---- fused: 3 launches -------------------------------------------------------------------------------------
    flat, grads, dWk, dbk = _grad_bufs(L, D, gated, K, dev)
    buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
    ops.amil_fused_step(xb, prep, flags, seed, buf, Wkd, bkd, Y, c, alpha, grads, dWk=dWk, dbk=dbk, loss_scale=scale,
                        zero=flat)
    torch.cuda.synchronize()

    # ---- modular: forward(train) -> head step kernel -> general stashed backward ---------------------------------
    flat2, grads2, dWk2, dbk2 = _grad_bufs(L, D, gated, K, dev)
    A_raw, parts, ws = ops.amil_partials_train(xb, prep, flags, seed, zero=flat2)
    t = ops.amil_head_nll_step(parts, Wkd, bkd, Y, c, alpha, dWk=dWk2, dbk=dbk2)
    ops.amil_backward(xb, prep, flags, seed, A_raw, t["ml"], t["M"], t["dM"] * scale, grads=grads2, stash=ws)
    assert torch.equal(buf.A_raw, A_raw), "the side MMA / mask words / folded head must not change the scores"
    assert rel_err(buf.M, t["M"]) < 1e-5 and rel_err(buf.ml, t["ml"]) < 1e-5
    assert rel_err(buf.hazards, t["hazards"]) < 1e-5 and rel_err(buf.S, t["S"]) < 1e-5
    assert torch.equal(buf.Y_hat, t["Y_hat"])
    assert abs(buf.loss.item() - t["loss"].item()) < 1e-5 * max(1.0, abs(t["loss"].item()))
    assert rel_err(buf.dM, t["dM"] * scale) < 1e-5
    assert rel_err(dWk, dWk2 * scale) < 1e-5 and rel_err(dbk, dbk2 * scale) < 1e-5
    for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc"):
        # same kernels downstream of t_i; t_i = dlogits·z_i (hi + lo bf16 split of Wk) vs dM·h_i in fp32
        assert close(grads[k], grads2[k], 2e-3), k

    # ---- oracle, bf16 operands -------------------------------------------------------------------------------
    hs = O.dropout_scale_mask(seed, 0, N, L) if drop & 2 else None
    as_ = O.dropout_scale_mask(seed, 1, N, D) if drop & 4 else None
    gs = O.dropout_scale_mask(seed, 2, N, D) if drop & 4 else None
    s, h, a, g = O.fc_attention(x, bfr(W1), b1, bfr(Wa), ba, bfr(Wb), bb, wc, bc, h_scale=hs, a_scale=as_, g_scale=gs,
                                round_h=True)
    Mo, m, l = O.softmax_pool(s, h)
    Mr = Mo.reshape(1, -1).clone().requires_grad_(True)
    Wr, br = Wk.clone().requires_grad_(True), bk.clone().requires_grad_(True)
    hz, S, Yh = O.hazard_head(Mr, Wr, br)
    loss = O.nll_surv_loss(hz, S, Y.cpu(), c.cpu(), alpha=alpha)
    (loss * scale).backward()
    assert rel_err(buf.A_raw, s) < TOL_FWD_TIGHT and rel_err(buf.M, Mo) < 1e-3
    assert rel_err(buf.hazards, hz) < 2e-3 and abs(buf.loss.item() - loss.item()) < 2e-3 * max(1.0, abs(loss.item()))
    go = O.amil_backward(x, bfr(W1), bfr(Wa), bfr(Wb), wc, s, h, a, g, Mo, m, l, Mr.grad.reshape(-1),
                         drop_h=bool(drop & 2), a_scale=as_, g_scale=gs)
    for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc"):
        assert close(grads[k], go[k], TOL_GRAD_TIGHT), k
    assert rel_err(dWk, Wr.grad) < TOL_GRAD_TIGHT and rel_err(dbk, br.grad) < TOL_GRAD_TIGHT


def test_fused_step_z_and_mask_words(dev):
    """What the forward leaves for the backward: z_i = Wk h_i (fp32-grade through the hi + lo bf16 split) and the
    ReLU mask words, checked against the stashed H tile the same launch wrote."""
    from multimodalfusion_b200 import ops
    from multimodalfusion_b200._lib import lib
    N, L, D, K = 777, 512, 384, 4
    W, Wk, bk = _rand(L, D, True, K, 3)
    xb = cases.features(N, 5).to(dev).to(torch.bfloat16)
    prep = ops.prepare_amil_weights(*[t.to(dev) for t in W])
    flags = ops.amil_flags(True) | 2
    flat, grads, dWk, dbk = _grad_bufs(L, D, True, K, dev)
    buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
    buf.pack_head(Wk.to(dev))
    wst = prep.struct()
    import ctypes as C
    from multimodalfusion_b200._lib import check
    head = buf.head_struct(Wk.to(dev), bk.to(dev), torch.tensor([1], device=dev), torch.tensor([0.0], device=dev), 0.0,
                           1e-7, 1.0, dWk, dbk)
    check(lib().mmf_amil_fwd_train_head(xb.data_ptr(), N, 1024, C.byref(wst), L, D, flags, 11, buf.A_raw.data_ptr(),
                                        buf.partials.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(),
                                        flat.data_ptr(), flat.numel(), C.byref(head),
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    # workspace layout (csrc/capi.cu bwd_layout): H | dG | dU | cs | dbc | db1 | mask | z, 1 KiB aligned slots
    def up(v):
        return (v + 1023) // 1024 * 1024
    tiles = (N + 255) // 256 * 2
    KD, ncols = 2 * D, 3 * D
    o_H = 0
    o_dG = up(N * L * 2)
    o_dU = up(o_dG + N * KD * 2)
    o_cs = up(o_dU + N * L * 2)
    o_dbc = up(o_cs + tiles * 4 * ncols * 4)
    o_db1 = up(o_dbc + tiles * 4 * 4)
    o_mask = up(o_db1 + tiles * 4 * L * 4)
    o_z = up(o_mask + N * (L // 32) * 4)
    ws = buf.workspace
    H = ws[o_H:o_H + N * L * 2].view(torch.bfloat16).view(N, L).float()
    mask = ws[o_mask:o_mask + N * (L // 32) * 4].view(torch.int32).view(N, L // 32)
    z = ws[o_z:o_z + N * 4 * 4].view(torch.float32).view(N, 4)
    bits = ((mask.unsqueeze(-1) >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(N, L).bool()
    assert torch.equal(bits, H > 0)
    z_ref = H.double() @ Wk.to(dev).double().t()
    assert (z.double() - z_ref).abs().max().item() < 2e-5 * z_ref.abs().max().item()


@pytest.mark.parametrize("N,L,D", [(10000, 256, 256), (16384, 512, 384)])
def test_fused_step_full_size_with_dropout_vs_fp32_oracle(dev, N, L, D):
    """BASELINE config 1 and the metric shape through the fused 3-launch step WITH train-mode dropout on h (what
    bench.py times), against the fp32 oracle (fp32 weights, fp32 h) at the north-star bars."""
    from multimodalfusion_b200 import ops
    K, seed = 4, 0x5EED
    W, Wk, bk = _rand(L, D, True, K, N)
    W1, b1, Wa, ba, Wb, bb, wc, bc = W
    x = cases.features(N, 1234)
    xb = x.to(dev).to(torch.bfloat16)
    prep = ops.prepare_amil_weights(*[t.to(dev) for t in W])
    flags = ops.amil_flags(True, dropout_h=True)
    Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
    flat, grads, dWk, dbk = _grad_bufs(L, D, True, K, dev)
    buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
    for _ in range(2):   # twice: buffers, ticket and the fused zero_grad are reusable as they are
        ops.amil_fused_step(xb, prep, flags, seed, buf, Wk.to(dev), bk.to(dev), Y, c, 0.0, grads, dWk=dWk, dbk=dbk,
                            zero=flat)
    hs = O.dropout_scale_mask(seed, 0, N, L)
    s32, h32, a32, g32 = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc, h_scale=hs)
    M32, m32, l32 = O.softmax_pool(s32, h32)
    Mr = M32.reshape(1, -1).clone().requires_grad_(True)
    Wr, br = Wk.clone().requires_grad_(True), bk.clone().requires_grad_(True)
    hz, S, _ = O.hazard_head(Mr, Wr, br)
    loss = O.nll_surv_loss(hz, S, Y.cpu(), c.cpu(), alpha=0.0)
    loss.backward()
    assert rel_err(buf.A_raw, s32) < TOL_FWD_REF and rel_err(buf.M, M32) < TOL_FWD_REF
    assert rel_err(buf.hazards, hz) < TOL_FWD_REF and abs(buf.loss.item() - loss.item()) < TOL_FWD_REF * abs(loss.item())
    go = O.amil_backward(x, W1, Wa, Wb, wc, s32, h32, a32, g32, M32, m32, l32, Mr.grad.reshape(-1), drop_h=True)
    for k in ("dW1", "db1", "dWab", "dbab", "dwc"):
        assert rel_err(grads[k], go[k]) < TOL_GRAD_REF, k
    assert rel_err(dWk, Wr.grad) < TOL_GRAD_REF and rel_err(dbk, br.grad) < TOL_GRAD_REF


@pytest.mark.parametrize("N,gate,bf16_bags", [(17, True, False), (80, True, False), (155, True, False), (131, False, False),
                                              (155, True, True), (96, False, True)])
def test_radio_fused_step_matches_autograd_path(dev, N, gate, bf16_bags):
    """MIL_Attention_fc_surv_radio.fused_step (reduce_dim + the fused 3-launch step with dx + reduce_dim wgrad, no
    autograd graph) == model(**bags) -> NLLSurvLoss -> backward through the autograd.Functions (eval mode: no dropout)."""
    from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
    from multimodalfusion_b200.utils import NLLSurvLoss
    torch.manual_seed(21)
    model = MIL_Attention_fc_surv_radio(gate_radio=gate, dropout=True, n_classes=4).to(dev).eval()
    bags = {m: cases.features(N, 300 + i).to(dev) for i, m in enumerate(model.modalities)}
    if bf16_bags:      # stored bf16 features: reduce_dim's weight gradient runs on the tensor cores (exact products)
        bags = {m: b.to(torch.bfloat16) for m, b in bags.items()}
    Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
    hz, S, Y_hat, A = model(**bags)
    loss = NLLSurvLoss(alpha=0.15)(hazards=hz, S=S, Y=Y, c=c)
    (loss * 0.5).backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    model.enable_fused_step()
    for rep in range(2):     # the second call must not accumulate
        hz2, S2, Y2, A2, loss2 = model.fused_step(Y=Y, c=c, alpha=0.15, loss_scale=0.5, **bags)
    assert rel_err(hz2, hz) < 2e-3 and rel_err(A2, A) < 2e-3 and abs(loss2.item() - loss.item()) < 2e-3 * max(1, abs(loss.item()))
    assert torch.equal(Y2.cpu(), Y_hat.cpu())
    for n, p in model.named_parameters():
        if n.endswith("attention_c.bias"):
            continue   # sum_i ds_i: exactly 0 in exact arithmetic, rounding residue on both sides
        assert close(p.grad, ref[n], 2e-2), n
    # accumulation over a gc window of two patients
    model.fused_step(Y=Y, c=c, alpha=0.15, loss_scale=0.5, accumulate=True, **bags)
    for n, p in model.named_parameters():
        if not n.endswith("attention_c.bias"):
            assert close(p.grad, 2 * ref[n], 2e-2), n
