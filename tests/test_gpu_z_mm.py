"""GPU parity of MM_MIL_Attention_fc_surv end to end and of its captum* entry points against the
reference's own fp32 outputs (tests/golden/reference_goldens_mm.pt; oracle/make_goldens_mm.py runs the reference class
after a run-time repair, without editing it). The module's Python glue is pinned on the CPU in tests/test_mm_glue_cpu.py;
here the same goldens go through the real kernels: segmented bf16 reduce_dim GEMM -> fused AMIL (radio, path) -> SNN ->
Kronecker / concat fusion -> hazard head -> nll_surv, and the backward of all of it.

Tolerances: hazards / S / pathology attention scores 1e-2 (radiology scores 2e-2: two bf16 stages), gradients 2e-2 + 3/N (bf16 operands; tiny bags, see
test_gpu_parity._grad_tol), risk order n/a (one patient); captum* run in fp32 on the functor SGEMM kernels: 1e-4.
(This file sorts last on purpose: it was added after the round's GPU budget was spent and has not run on a device yet.)"""
import pytest
import torch

from helpers import build_mm_model, build_unimodal_model, rel_err, unimodal_input
from oracle import cases

pytestmark = pytest.mark.gpu
TOL_FWD_REF = 1e-2


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda")


def _robust_close(t, fp, tol, min_frac=0.98, norm_tol=None):
    """Fingerprint comparison that tolerates a flipped ReLU: >= min_frac of the sampled entries within tol * max|ref| and
    the Frobenius norm within norm_tol (default tol). Operand rounding (bf16 embeddings feeding the fp32 fusion layers; fp32 summation order)
    can flip a pre-activation that sits at zero, which switches one whole row of a weight gradient on or off — a couple of
    the 1024 sampled entries — without saying anything about the kernels; a systematic error moves every entry."""
    flat = t.detach().reshape(-1).float().cpu()
    ref = fp["vals"]
    scale = max(ref.abs().max().item(), 1e-30)
    err = (flat[cases._sample_idx(flat.numel())] - ref).abs() / scale
    frac = (err <= tol).float().mean().item()
    norm_err = abs(flat.double().norm().item() - fp["norm"]) / max(fp["norm"], 1e-30)
    ok = frac >= min_frac and norm_err <= max(tol if norm_tol is None else norm_tol, 1e-4)
    return ok, f"{100 * frac:.1f}% within {tol}, max {err.max().item():.2e}, norm {norm_err:.2e}"


def _check_param_grads(model, gold_grads, tol, relu_gated=()):
    """relu_gated: name prefixes of the fp32 layers that sit behind ReLUs fed by bf16-rounded embeddings (the fusion block
    and the classifier of the multimodal model: ~1300 ReLU units per patient, a pre-activation within ~2e-3 of zero flips
    with probability ~1.6e-3 each). One flipped unit switches a whole row of its layer's weight gradient — 1/16 of the
    entries of a 16-unit `reduce` layer — so those layers are held to >= 90 % of the sampled entries within tol and the
    norm within 5 %; everything else to 98 % and tol."""
    bad = {}
    for k, p in model.named_parameters():
        fp = gold_grads[k]
        if fp is None or fp["norm"] == 0.0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-4, k
            continue
        assert p.grad is not None, k
        if fp["vals"].abs().max().item() <= 1e-6:
            # e.g. the attention bias bc: the soft-max is shift-invariant, its gradient is rounding noise on both sides
            assert p.grad.abs().max().item() <= 1e-4, k
            continue
        loose = any(k.startswith(pre) for pre in relu_gated)
        ok, msg = _robust_close(p.grad, fp, tol, 0.90 if loose else 0.98, max(tol, 5e-2) if loose else None)
        if not ok:
            bad[k] = msg
    assert not bad, f"gradients beyond {tol}: {bad}"


def _check_param_grads_strict(model, gold_grads, tol):
    """North-star bar: every parameter gradient within tol (max |err| / max |ref| over the sampled entries)."""
    worst = {}
    for k, p in model.named_parameters():
        fp = gold_grads[k]
        if fp is None or fp["norm"] == 0.0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-4, k
            continue
        assert p.grad is not None, k
        ref = fp["vals"]
        if ref.abs().max().item() <= 1e-6:
            assert p.grad.abs().max().item() <= 1e-4, k
            continue
        got = p.grad.detach().reshape(-1).float().cpu()[cases._sample_idx(p.numel())]
        worst[k] = (got - ref).abs().max().item() / ref.abs().max().item()
    bad = {k: f"{v:.3e}" for k, v in worst.items() if v >= tol}
    assert not bad, f"gradients beyond {tol}: {bad} (worst of the rest: {max(worst.values()):.3e})"


@pytest.mark.parametrize("name", list(cases.MM_CASES))
def test_mm_model_vs_reference_goldens(dev, goldens_mm, name):
    from multimodalfusion_b200.utils import NLLSurvLoss
    cfg, gold = cases.MM_CASES[name], goldens_mm["mm"][name]
    model = build_mm_model(cfg).to(dev)
    kw = {k: v.to(dev) for k, v in cases.mm_inputs(cfg).items()}
    Y, c = cases.labels(cfg)
    hazards, S, Y_hat, A_raw = model(**kw)
    assert set(A_raw) == set(gold["A_raw"])
    for k in A_raw:
        # radiology scores pass through TWO bf16 stages (reduce_dim writes a bf16 bag, then the fused AMIL kernel): the
        # bf16-operand restatement itself reaches 0.9e-2 on these 17-40-slice bags (tests/test_mm_glue_cpu.py), so they
        # get 2e-2; pathology scores and the hazards keep the north star's 1e-2
        tol = TOL_FWD_REF
        assert A_raw[k].shape == gold["A_raw"][k].shape and rel_err(A_raw[k], gold["A_raw"][k]) < tol, k
    assert rel_err(hazards, gold["hazards"]) < TOL_FWD_REF and rel_err(S, gold["S"]) < TOL_FWD_REF
    assert Y_hat.shape == (1, 1) and Y_hat.dtype == torch.int64
    loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y.to(dev), c=c.to(dev))
    assert abs(loss.item() - gold["loss"].item()) < 1e-2 * max(1.0, abs(gold["loss"].item()))
    model.zero_grad()
    loss.backward()
    n_min = min(n for n in (cfg["Nr"], cfg["Np"]) if n)
    _check_param_grads_strict(model, gold["grads"], 2e-2)
    feats = model(**kw, return_features=True)
    assert feats.shape == (1, 512 if cfg["fusion"] == "tensor" else 256 * len(cfg["mode"].split("_")))


@pytest.mark.parametrize("name", list(cases.CAPTUM_CASES))
def test_captum_entry_points_vs_reference_goldens(dev, goldens_mm, name):
    cfg, gold = cases.CAPTUM_CASES[name], goldens_mm["captum"][name]
    model = build_mm_model(cfg).to(dev)
    args, w = cases.captum_inputs(cfg)
    args = [a.to(dev).requires_grad_(True) for a in args]
    risk = getattr(model, cfg["fn"])(*args)
    assert risk.shape == gold["risk"].shape and rel_err(risk, gold["risk"]) < 1e-4
    model.zero_grad()
    (risk * w.to(dev)).sum().backward()
    for a, fp in zip(args, gold["d_inputs"]):
        assert a.grad is not None
        ok, msg = _robust_close(a.grad, fp, 1e-4)
        assert ok, f"input attribution: {msg}"
    _check_param_grads(model, gold["grads"], 1e-4)


@pytest.mark.parametrize("name", list(cases.UNI_CASES))
def test_unimodal_heads_vs_reference_goldens(dev, goldens_unimodal, name):
    """unimonal_pretrained (fcnn / highway / residual on one modality's embedding) of both head files on the library's
    fp32 kernels (Dense, BatchNorm1d, highway mix, hazard head) against the reference: risk / hazards, loss, risk order,
    input gradient, every parameter gradient."""
    from multimodalfusion_b200.utils import CoxSurvLoss, CrossEntropySurvLoss, NLLSurvLoss, RankingSurvLoss
    cfg, gold = cases.UNI_CASES[name], goldens_unimodal["unimodal"][name]
    model = build_unimodal_model(cfg).to(dev)
    h = unimodal_input(cfg).to(dev).requires_grad_(True)
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    res = model(**{"h_" + cfg["mode"]: h})
    if cfg["kind"] == "cox":
        risk = res[0]
        assert res[1] is None and res[2] is None
        loss = (CoxSurvLoss()(risks=risk, times=times.to(dev), c=c.to(dev)) if cfg["loss"] == "cox"
                else RankingSurvLoss()(risks=risk.reshape(-1), times=times.to(dev), c=c.to(dev)))
    else:
        risk, hazards, S = res
        assert rel_err(hazards, gold["hazards"]) < 1e-5 and rel_err(S, gold["S"]) < 1e-5
        lf = NLLSurvLoss(alpha=0.15) if cfg["loss"] == "nll" else CrossEntropySurvLoss(alpha=0.15)
        loss = lf(hazards=hazards, S=S, Y=(torch.arange(cfg["B"]) % 4).to(dev), c=c.to(dev))
    assert risk.shape == gold["risk"].shape and rel_err(risk, gold["risk"]) < 1e-5
    assert torch.equal(torch.argsort(risk.reshape(-1).cpu()), torch.argsort(gold["risk"].reshape(-1)))
    assert abs(loss.item() - gold["loss"].item()) < 1e-5
    model.zero_grad()
    loss.backward()
    assert rel_err(h.grad, gold["d_input"]) < 1e-4
    _check_param_grads(model, gold["grads"], 1e-4)
