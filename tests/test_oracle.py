"""CPU: pins the oracle (oracle/amil_oracle.py) to the reference's own outputs (tests/golden), and
checks the algorithmic restatements the kernels implement (tile partials + combine, analytic
backward, sort+scan Cox) against the plain formulation."""
import pytest
import torch

from helpers import (amil_weights, build_head_model, build_omic_model, build_path_model, build_radio_model,
                     rel_err, xfusion_params)
from oracle import amil_oracle as O
from oracle import cases

TOL = 2e-5  # fp32 reassociation between the reference's ATen ops and the restatement


@pytest.mark.parametrize("name", list(cases.PATH_CASES))
def test_path_forward_backward_matches_reference(goldens, name):
    cfg, gold = cases.PATH_CASES[name], goldens["path"][name]
    model = build_path_model(cfg)
    # seeded construction reproduces the reference's weights exactly
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    W1, b1, Wa, ba, Wb, bb, wc, bc = amil_weights(model.attention_net_WSI)
    x = cases.path_bag(cfg)
    s, h, a, g = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc)
    M, m, l = O.softmax_pool(s, h)
    Wk, bk = model.classifier.weight.detach(), model.classifier.bias.detach()
    Mr = M.reshape(1, -1).clone().requires_grad_(True)
    hazards, S, Y_hat = O.hazard_head(Mr, Wk, bk)
    Y, c = cases.labels(cfg)
    loss = O.nll_surv_loss(hazards, S, Y, c, alpha=cfg["alpha"])
    assert rel_err(s, gold["A_raw"]) < TOL
    assert rel_err(M, gold["M"]) < TOL
    assert rel_err(hazards, gold["hazards"]) < TOL and rel_err(S, gold["S"]) < TOL
    assert torch.equal(Y_hat, gold["Y_hat"])
    assert abs(loss.item() - gold["loss"].item()) < TOL * max(1.0, abs(gold["loss"].item()))
    # tile partials + combine == plain softmax pooling
    M2, m2, l2 = O.combine_partials(O.tile_partials(s, h))
    assert rel_err(M2, M) < 1e-5 and abs(m2 - m) < 1e-6
    # analytic backward == autograd of the reference
    loss.backward()
    grads = O.amil_backward(x, W1, Wa, Wb, wc, s, h, a, g, M, m, l, Mr.grad.reshape(-1))
    pre = "attention_net_WSI."
    D = Wa.shape[0]
    if cfg["gated"]:
        named = {pre + "0.weight": grads["dW1"], pre + "0.bias": grads["db1"],
                 pre + "3.attention_a.0.weight": grads["dWab"][:D], pre + "3.attention_b.0.weight": grads["dWab"][D:],
                 pre + "3.attention_a.0.bias": grads["dbab"][:D], pre + "3.attention_b.0.bias": grads["dbab"][D:],
                 pre + "3.attention_c.weight": grads["dwc"].reshape(1, -1), pre + "3.attention_c.bias": grads["dbc"]}
    else:
        last = "3" if cfg["dropout"] else "2"
        named = {pre + "0.weight": grads["dW1"], pre + "0.bias": grads["db1"],
                 pre + "3.module.0.weight": grads["dWab"], pre + "3.module.0.bias": grads["dbab"],
                 pre + f"3.module.{last}.weight": grads["dwc"].reshape(1, -1),
                 pre + f"3.module.{last}.bias": grads["dbc"]}
    for k, v in named.items():
        cases.check_fingerprint(v, gold["grads"][k], 2e-4, f"grad {k}", atol=1e-6)


@pytest.mark.parametrize("name", list(cases.RADIO_CASES))
def test_radio_forward_matches_reference(goldens, name):
    cfg, gold = cases.RADIO_CASES[name], goldens["radio"][name]
    model = build_radio_model(cfg)
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    bags = cases.radio_bags(cfg)
    x = torch.cat([bags[m] for m in model.modalities], 1) @ model.reduce_dim.weight.detach().t() \
        + model.reduce_dim.bias.detach()
    W = amil_weights(model.attention_net_radio)
    s, h, a, g = O.fc_attention(x, *W)
    M, m, l = O.softmax_pool(s, h)
    hazards, S, _ = O.hazard_head(M.reshape(1, -1), model.classifier.weight.detach(), model.classifier.bias.detach())
    assert rel_err(s, gold["A_raw"]) < TOL and rel_err(M, gold["M"]) < TOL
    assert rel_err(hazards, gold["hazards"]) < TOL


@pytest.mark.parametrize("name", list(cases.OMIC_CASES))
def test_snn_matches_reference(goldens, name):
    cfg, gold = cases.OMIC_CASES[name], goldens["omic"][name]
    model = build_omic_model(cfg)
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    layers = [(blk[0].weight.detach(), blk[0].bias.detach()) for blk in model.fc_omic]
    x = cases.omic_batch(cfg)
    feats = O.snn_forward(x, layers)
    risk = (feats @ model.classifier.weight.detach().t() + model.classifier.bias.detach()).squeeze()
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    assert rel_err(feats, gold["features"]) < TOL and rel_err(risk, gold["risk"]) < TOL
    assert abs(O.cox_loss(risk, times, c).item() - gold["loss"].item()) < 1e-5


@pytest.mark.parametrize("name", list(cases.HEAD_CASES))
def test_kronecker_heads_match_reference(goldens, name):
    cfg, gold = cases.HEAD_CASES[name], goldens["heads"][name]
    model = build_head_model(cfg)
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    hr, hp, ho = cases.embeddings(cfg)
    from multimodalfusion_b200.models.coxranking_models_pretrained import _pick
    vs = _pick(cfg["mode"], hr, hp, ho)
    red, e1, e2 = xfusion_params(model.xfusion)
    MM = O.xfusion_forward(vs, red, e1, e2, skip=True)
    out = MM @ model.classifier.weight.detach().t() + model.classifier.bias.detach()
    if cfg["kind"] == "cox":
        assert rel_err(out, gold["risk"]) < 5e-5
    else:
        hz = torch.sigmoid(out)
        assert rel_err(hz, gold["hazards"]) < 5e-5
        assert rel_err(-torch.cumprod(1 - hz, 1).sum(1), gold["risk"]) < 5e-5


@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
def test_losses_match_reference(goldens, name):
    cfg, gold = cases.LOSS_CASES[name], goldens["losses"][name]
    if cfg["loss"] == "nll":
        hz, S, Y, c = cases.nll_inputs(cfg)
        hz.requires_grad_(True); S.requires_grad_(True)
        loss = O.nll_surv_loss(hz, S, Y, c, alpha=cfg["alpha"])
        loss.backward()
        assert abs(loss.item() - gold["loss"].item()) < 1e-5 * max(1, abs(gold["loss"].item()))
        assert rel_err(hz.grad, gold["d_hazards"]) < 1e-5 and rel_err(S.grad, gold["d_S"]) < 1e-5
        return
    r, times, c = cases.risk_inputs(cfg)
    r.requires_grad_(True)
    if cfg["loss"] == "cox":
        loss = O.cox_loss(r, times, c)
        assert abs(O.cox_loss_sorted(r.detach(), times, c).item() - gold["loss"].item()) < 1e-5
    else:
        loss = O.ranking_loss(r, times, c, cfg["phi"], cfg["reduction"]).reshape(())
    assert abs(loss.item() - gold["loss"].item()) < 1e-5 * max(1, abs(gold["loss"].item()))
    if loss.requires_grad:
        loss.backward()
        assert (r.grad - gold["d_risk"]).abs().max().item() < 1e-6 + 1e-5 * gold["d_risk"].abs().max().item()


def test_dropout_mask_statistics_and_determinism():
    m1 = O.dropout_scale_mask(0x1234_5678_9ABC, 0, 257, 512)
    m2 = O.dropout_scale_mask(0x1234_5678_9ABC, 0, 257, 512)
    m3 = O.dropout_scale_mask(0x1234_5678_9ABC, 1, 257, 512)
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    keep = (m1 > 0).float().mean().item()
    assert abs(keep - 0.75) < 0.01
    u = m1.unique().tolist()
    assert len(u) == 2 and u[0] == 0.0 and abs(u[1] - 1 / 0.75) < 1e-6


def test_dropout_mask_any_rate_statistics_and_determinism():
    """The 16-bit-field mask of the Kronecker encoder (csrc/small_kernels.cuh KronElem::keep_scale, drop > 1)."""
    for p in (0.7, 0.1, 0.5):
        m1 = O.dropout_scale_mask_p(0xABCDEF0123, 3, 129, 4913, p)
        assert torch.equal(m1, O.dropout_scale_mask_p(0xABCDEF0123, 3, 129, 4913, p))
        assert not torch.equal(m1, O.dropout_scale_mask_p(0xABCDEF0124, 3, 129, 4913, p))
        keep = (m1 > 0).float().mean().item()
        assert abs(keep - (1 - p)) < 0.01
        t = round(p * 65536)
        u = m1.unique().tolist()
        assert len(u) == 2 and u[0] == 0.0 and abs(u[1] - 65536 / (65536 - t)) < 1e-5
        assert abs(m1.mean().item() - 1.0) < 0.02          # inverted dropout keeps the expectation


def test_kron_dropout_code():
    from multimodalfusion_b200.ops import kron_dropout_code
    assert kron_dropout_code(False) == 0 and kron_dropout_code(0.0) == 0 and kron_dropout_code(True) == 1
    assert kron_dropout_code(0.25) == 1 and kron_dropout_code(0.7) == 45875 and kron_dropout_code(1e-9) == 2
    with pytest.raises(ValueError):
        kron_dropout_code(1.0)


def test_train_mode_backward_restatement_matches_autograd():
    """Analytic backward with dropout masks == autograd through the masked forward."""
    torch.manual_seed(3)
    N, L, D = 70, 256, 256
    x = cases.features(N, 77)
    W1 = (torch.randn(L, 1024) * 0.03).requires_grad_(True)
    b1 = (torch.randn(L) * 0.05).requires_grad_(True)
    Wa = (torch.randn(D, L) * 0.06).requires_grad_(True); ba = (torch.randn(D) * 0.05).requires_grad_(True)
    Wb = (torch.randn(D, L) * 0.06).requires_grad_(True); bb = (torch.randn(D) * 0.05).requires_grad_(True)
    wc = (torch.randn(1, D) * 0.1).requires_grad_(True); bc = torch.zeros(1, requires_grad=True)
    hs, as_, gs = (O.dropout_scale_mask(99, i, N, n) for i, n in ((0, L), (1, D), (2, D)))
    s, h, a, g = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc, h_scale=hs, a_scale=as_, g_scale=gs)
    M, m, l = O.softmax_pool(s, h)
    dM = torch.randn(L)
    dA = torch.randn(N) * 0.01
    ((M * dM).sum() + (s * dA).sum()).backward()
    with torch.no_grad():
        gr = O.amil_backward(x, W1, Wa, Wb, wc, s, h, a, g, M, m, l, dM, dA, drop_h=True, a_scale=as_, g_scale=gs)
    assert rel_err(gr["dW1"], W1.grad) < 1e-4 and rel_err(gr["db1"], b1.grad) < 1e-4
    assert rel_err(gr["dWab"], torch.cat([Wa.grad, Wb.grad])) < 1e-4
    assert rel_err(gr["dbab"], torch.cat([ba.grad, bb.grad])) < 1e-4
    assert rel_err(gr["dwc"], wc.grad.reshape(-1)) < 1e-4 and rel_err(gr["dbc"], bc.grad) < 1e-4


def test_concordance_index_simple():
    risk, t, e = [3.0, 2.0, 1.0], [1.0, 2.0, 3.0], [1, 1, 0]
    assert O.concordance_index(risk, t, e) == 1.0
    assert O.concordance_index(risk[::-1], t, e) == 0.0


def test_concordance_index_tied_times_follow_sksurv():
    """sksurv.concordance_index_censored: an event at time t is comparable to a sample CENSORED at the same t (and to
    nothing else at t). Hand-checked vector + the literal restatement of sksurv's _get_comparable on random tied data."""
    import numpy as np
    # times: two samples at t = 2 (one event, one censored), one later event, one earlier censored
    t = [2.0, 2.0, 3.0, 1.0]
    e = [1, 0, 1, 0]
    risk = [0.9, 0.1, 0.5, 0.7]
    # comparable: (0,1) tie-in-time event vs censored, (0,2) t0 < t2. Sample 3 is censored: starts no pair. Sample 2: nobody later.
    assert O.sksurv_comparable_pairs(t, e) == {(0, 1), (0, 2)}
    assert O.concordance_index(risk, t, e) == 1.0                 # 0.9 > 0.1 and 0.9 > 0.5
    assert O.concordance_index([0.1, 0.9, 0.5, 0.7], t, e) == 0.0
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = 60
        times = rng.integers(1, 12, n).astype(float)              # discretised months: many ties
        event = rng.integers(0, 2, n)
        r = rng.standard_normal(n)
        pairs = O.sksurv_comparable_pairs(times, event)
        conc = sum(r[i] > r[j] for i, j in pairs)
        want = conc / len(pairs)
        assert abs(O.concordance_index(r, times, event) - want) < 1e-12


@pytest.mark.parametrize("name", list(cases.HEAD2_CASES))
def test_fcnn_highway_heads_match_reference(goldens_heads2, name):
    """Oracle restatement of the fcnn / Highway fusion heads + ce_loss against the reference's own outputs
    (tests/golden/reference_goldens_heads2.pt, generated by oracle/make_goldens_heads2.py), weights rebuilt from
    the seed through the drop-in constructors (fingerprint-checked)."""
    from helpers import build_head2_model
    cfg, gold = cases.HEAD2_CASES[name], goldens_heads2["heads2"][name]
    model = build_head2_model(cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, fp in gold["weights_fp"].items():
        cases.check_fingerprint(sd[k], fp, 0.0, k)
    hr, hp, ho = cases.embeddings(cfg)
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    out = O.fusion_head2_eval(sd, cfg["kind"], cfg["train_type"], cfg["mode"], cfg["n_layers"], hr, hp, ho)
    if cfg["kind"] == "cox":
        assert out.shape == gold["risk"].shape
        assert torch.allclose(out, gold["risk"], rtol=1e-5, atol=1e-6)
        loss = O.cox_loss(out.reshape(-1), times, c) if cfg["loss"] == "cox" else O.ranking_loss(out.reshape(-1), times, c)
    else:
        hz = torch.sigmoid(out)
        S = torch.cumprod(1 - hz, dim=1)
        assert torch.allclose(hz, gold["hazards"], rtol=1e-5, atol=1e-6) and torch.allclose(-S.sum(1), gold["risk"], rtol=1e-5, atol=1e-6)
        Y = torch.arange(cfg["B"]) % 4
        loss = (O.nll_surv_loss(hz, S, Y, c, alpha=0.15) if cfg["loss"] == "nll" else O.ce_surv_loss(hz, S, Y, c, alpha=0.15))
    assert abs(float(loss) - gold["loss"].item()) < 1e-5


def test_batchnorm_train_restatement_matches_torch():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 128, generator=g) * 2 + 0.5
    w, b = torch.randn(128, generator=g), torch.randn(128, generator=g)
    bn = torch.nn.BatchNorm1d(128).train()
    with torch.no_grad():
        bn.weight.copy_(w); bn.bias.copy_(b)
    y = bn(x)
    yo, mean, var_u = O.batchnorm1d_train(x, w, b)
    assert torch.allclose(y, yo, rtol=1e-5, atol=1e-5)
    assert torch.allclose(bn.running_mean, 0.1 * mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, 0.9 + 0.1 * var_u, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(cases.XFUSION4_CASES))
def test_xfusion_four_modalities_matches_reference(goldens_xfusion4, name):
    """Drop-in XlinearFusion() (reference default: 4 modalities, 17^4-wide product) draws bit-identical initial weights,
    and the oracle restatement reproduces the reference's fused features."""
    from helpers import build_xfusion4
    cfg, gold = cases.XFUSION4_CASES[name], goldens_xfusion4["xfusion4"][name]
    model = build_xfusion4(cfg)
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    assert set(model.state_dict()) == set(gold["weights_fp"])
    vs, proj = cases.embeddings4(cfg)
    red, e1, e2 = xfusion_params(model)
    feats = O.xfusion_forward(vs, red, e1, e2, skip=True)
    assert rel_err(feats, gold["features"]) < 5e-5
    assert abs((feats * proj).sum().item() - gold["loss"].item()) < 1e-4


@pytest.mark.parametrize("name", list(cases.RADIO_TENSOR_CASES))
def test_radio_tensor_fusion_matches_repaired_reference(goldens_xfusion4, name):
    """radio_fusion='tensor' (one-name repair of the reference): identical initial weights and the
    oracle composition xfusion(slice 0 of each modality) -> one-row AMIL -> hazard head reproduces the golden."""
    from helpers import build_radio_tensor_model
    cfg, gold = cases.RADIO_TENSOR_CASES[name], goldens_xfusion4["radio_tensor"][name]
    model = build_radio_tensor_model(cfg)
    assert set(model.state_dict()) == set(gold["weights_fp"])
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    bags = cases.radio_bags(cfg)
    red, e1, e2 = xfusion_params(model.radio_xfusion)
    x1 = O.xfusion_forward([bags[m][0:1] for m in model.modalities], red, e1, e2, skip=False)
    s, h, _, _ = O.fc_attention(x1, *amil_weights(model.attention_net_radio))
    M, _, _ = O.softmax_pool(s, h)
    M = M.reshape(1, -1)
    hz, S, _ = O.hazard_head(M, model.classifier.weight.detach(), model.classifier.bias.detach())
    assert rel_err(s.reshape(1, -1), gold["A_raw"]) < 1e-4 and rel_err(M, gold["M"]) < 1e-4
    assert rel_err(hz, gold["hazards"]) < 1e-4 and rel_err(S, gold["S"]) < 1e-4


@pytest.mark.parametrize("gated", [True, False])
def test_head_projected_backward_equals_the_general_one(gated):
    """Spec of the round-2 'head-projected phase A' (DESIGN.md §8): with a linear classifier on the pooled embedding,
    t_i = dM.h_i can be formed from z_i = Wk h_i (K floats per instance, emitted by the forward) and dlogits — every
    gradient equals the general backward's to fp32 re-association, train-mode dropout scalings included."""
    g = torch.Generator().manual_seed(17)
    N, L, D, K = 300, 256, 256, 4
    x = cases.features(N, 91)
    W1, b1 = torch.randn(L, 1024, generator=g) * 0.03, torch.randn(L, generator=g) * 0.05
    Wa, ba = torch.randn(D, L, generator=g) * 0.06, torch.randn(D, generator=g) * 0.05
    Wb, bb = (torch.randn(D, L, generator=g) * 0.06, torch.randn(D, generator=g) * 0.05) if gated else (None, None)
    wc, bc = torch.randn(1, D, generator=g) * 0.1, torch.zeros(1)
    Wk, bk = torch.randn(K, L, generator=g) * 0.05, torch.randn(K, generator=g) * 0.05
    hs, as_, gs = O.dropout_scale_mask(7, 0, N, L), O.dropout_scale_mask(7, 1, N, D), O.dropout_scale_mask(7, 2, N, D)
    s, h, a, gg = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc, h_scale=hs, a_scale=as_, g_scale=gs)
    M, m, l = O.softmax_pool(s, h)
    Mr = M.reshape(1, -1).clone().requires_grad_(True)
    logits = Mr @ Wk.t() + bk
    logits.retain_grad()
    hazards = torch.sigmoid(logits)
    loss = O.nll_surv_loss(hazards, torch.cumprod(1 - hazards, dim=1), torch.tensor([2]), torch.tensor([0.0]), alpha=0.15)
    loss.backward()
    dA = torch.randn(N, generator=g) * 1e-3
    kw = dict(drop_h=True, a_scale=as_, g_scale=gs, need_dx=True)
    ref = O.amil_backward(x, W1, Wa, Wb, wc, s, h, a, gg, M, m, l, Mr.grad.reshape(-1), dA, **kw)
    z, mask = O.head_projection(h, Wk)
    assert z.shape == (N, K) and mask.dtype == torch.bool
    got = O.amil_backward_head_projected(x, W1, Wa, Wb, wc, s, h, a, gg, M, m, l, logits.grad.reshape(-1), Wk, z, mask,
                                         dA, **kw)
    assert set(got) == set(ref)
    for k in ref:
        assert rel_err(got[k], ref[k]) < 2e-5, k
