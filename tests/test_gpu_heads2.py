"""GPU parity of the fcnn / Highway fusion heads, BatchNorm1d and ce_loss: the drop-in modules on
the library's kernels against the reference's own fp32 outputs (tests/golden/reference_goldens_heads2.pt) — risk /
hazards, per-cohort risk ORDER, loss, gradients of every parameter and of the input embeddings — and BatchNorm1d in
training mode (batch statistics, running-stat update, backward) against torch on the CPU."""
import pytest
import torch

from helpers import build_head2_model, rel_err
from oracle import amil_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda")


@pytest.mark.parametrize("name", list(cases.HEAD2_CASES))
def test_fcnn_highway_heads_vs_reference_goldens(dev, goldens_heads2, name):
    from multimodalfusion_b200.utils import CoxSurvLoss, CrossEntropySurvLoss, NLLSurvLoss, RankingSurvLoss
    cfg, gold = cases.HEAD2_CASES[name], goldens_heads2["heads2"][name]
    model = build_head2_model(cfg).to(dev)
    hr, hp, ho = [t.to(dev).requires_grad_(True) for t in cases.embeddings(cfg)]
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    res = model(hr, hp, ho)
    if cfg["kind"] == "cox":
        risk = res[0]
        assert res[1] is None and res[2] is None and risk.shape == gold["risk"].shape
        loss = (CoxSurvLoss()(risks=risk, times=times.to(dev), c=c.to(dev)) if cfg["loss"] == "cox"
                else RankingSurvLoss()(risks=risk.reshape(-1), times=times.to(dev), c=c.to(dev)))
    else:
        risk, hazards, S = res
        assert rel_err(hazards, gold["hazards"]) < 1e-5 and rel_err(S, gold["S"]) < 1e-5
        lf = NLLSurvLoss(alpha=0.15) if cfg["loss"] == "nll" else CrossEntropySurvLoss(alpha=0.15)
        loss = lf(hazards=hazards, S=S, Y=(torch.arange(cfg["B"]) % 4).to(dev), c=c.to(dev))
    assert rel_err(risk, gold["risk"]) < 1e-5
    assert torch.equal(torch.argsort(risk.reshape(-1).cpu()), torch.argsort(gold["risk"].reshape(-1)))
    assert abs(loss.item() - gold["loss"].item()) < 1e-5
    model.zero_grad()
    loss.backward()
    for t, gd in zip((hr, hp, ho), gold["d_inputs"]):
        if gd is not None:
            assert t.grad is not None and rel_err(t.grad, gd) < 1e-4
    for k, p in model.named_parameters():
        fp = gold["grads"][k]
        if fp is None:                      # branch of a modality the mode does not use: no gradient in the reference
            assert p.grad is None or p.grad.abs().max().item() == 0, k
            continue
        ref = fp["vals"]
        got = p.grad.detach().reshape(-1).float().cpu()[cases._sample_idx(p.numel())]
        scale = max(ref.abs().max().item(), 1e-30)
        assert (got - ref).abs().max().item() <= 1e-4 * scale + 1e-7, k


@pytest.mark.parametrize("B,F", [(2, 128), (37, 128), (512, 768), (33, 100)])
def test_batchnorm1d_train_and_eval_vs_torch(dev, B, F):
    from multimodalfusion_b200.models.model_modules import batchnorm1d_forward
    g = torch.Generator().manual_seed(B + F)
    x = torch.randn(B, F, generator=g) * 1.7 + 0.3
    dy = torch.randn(B, F, generator=g)
    ref = torch.nn.BatchNorm1d(F)
    ours = torch.nn.BatchNorm1d(F)
    for m in (ref, ours):
        cases.perturb_bn(m, 5)
    ours = ours.to(dev)
    for train in (True, False):
        ref.train(train); ours.train(train)
        xr = x.clone().requires_grad_(True)
        xo = x.clone().to(dev).requires_grad_(True)
        yr = ref(xr)
        yo = batchnorm1d_forward(ours, xo)
        assert rel_err(yo, yr) < 1e-5
        ref.zero_grad(); ours.zero_grad()
        yr.backward(dy); yo.backward(dy.to(dev))
        assert rel_err(xo.grad, xr.grad) < 2e-5
        assert rel_err(ours.weight.grad, ref.weight.grad) < 2e-5 and rel_err(ours.bias.grad, ref.bias.grad) < 2e-5
        assert rel_err(ours.running_mean, ref.running_mean) < 1e-5 and rel_err(ours.running_var, ref.running_var) < 1e-5
        assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked)


def test_batchnorm1d_single_sample_training_raises_like_torch(dev):
    from multimodalfusion_b200._lib import MmfError
    from multimodalfusion_b200.models.model_modules import batchnorm1d_forward
    bn = torch.nn.BatchNorm1d(16).to(dev).train()
    with pytest.raises(MmfError):
        batchnorm1d_forward(bn, torch.randn(1, 16, device=dev))


@pytest.mark.parametrize("B,K,alpha", [(7, 4, 0.0), (64, 8, 0.15), (5, 4, 0.4)])
def test_ce_loss_vs_oracle(dev, B, K, alpha):
    from multimodalfusion_b200.utils import ce_loss
    hz, S, Y, c = cases.nll_inputs(dict(seed=90 + B, B=B, K=K))   # (ce_loss is inf on saturated inputs in the reference too)
    hr, Sr = hz.clone().requires_grad_(True), S.clone().requires_grad_(True)
    want = O.ce_surv_loss(hr, Sr, Y, c, alpha=alpha)
    want.backward()
    ho, So = hz.clone().to(dev).requires_grad_(True), S.clone().to(dev).requires_grad_(True)
    got = ce_loss(ho, So, Y.to(dev), c.to(dev), alpha=alpha)
    got.backward()
    assert abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel_err(ho.grad, hr.grad) < 1e-5 and rel_err(So.grad, Sr.grad) < 1e-4


@pytest.mark.parametrize("name", list(cases.XFUSION4_CASES))
def test_xfusion_four_modalities_vs_reference_goldens(dev, goldens_xfusion4, name):
    """XlinearFusion() with the reference's default num_modalities=4: the 83 521-wide Kronecker product is formed inside
    the encoder1 kernel (mmf_kron_enc_fwd, m = 4); features, input gradients and every parameter gradient against the
    reference's fp32 outputs (models/model_modules.py:156-178)."""
    from helpers import build_xfusion4
    cfg, gold = cases.XFUSION4_CASES[name], goldens_xfusion4["xfusion4"][name]
    model = build_xfusion4(cfg).to(dev)
    vs, proj = cases.embeddings4(cfg)
    vs = [v.to(dev).requires_grad_(True) for v in vs]
    feats = model(v_list=vs)
    assert feats.shape == gold["features"].shape and rel_err(feats, gold["features"]) < 1e-5
    loss = (feats * proj.to(dev)).sum()
    assert abs(loss.item() - gold["loss"].item()) < 1e-4
    model.zero_grad()
    loss.backward()
    for v, gd in zip(vs, gold["d_inputs"]):
        assert rel_err(v.grad, gd) < 1e-4
    for k, p in model.named_parameters():
        cases.check_fingerprint(p.grad, gold["grads"][k], 1e-4, f"grad {k}", atol=1e-7)


@pytest.mark.parametrize("name", list(cases.RADIO_TENSOR_CASES))
def test_radio_tensor_fusion_vs_repaired_reference(dev, goldens_xfusion4, name):
    """MIL_Attention_fc_surv_radio(radio_fusion='tensor'): 4-way Kronecker fusion of slice 0 of each modality (fp32
    kernels) feeding a ONE-row bag through the fused bf16 AMIL kernel, with dX flowing back into the fusion. Against the
    reference run with its one-name repair (oracle/make_goldens_xfusion4.py): forward 1e-2, gradients 3e-2
    (bf16 AMIL + bf16 dX in the chain). The attention net gets an exactly-zero gradient from a one-row softmax."""
    from helpers import build_radio_tensor_model
    from multimodalfusion_b200.utils import NLLSurvLoss
    cfg, gold = cases.RADIO_TENSOR_CASES[name], goldens_xfusion4["radio_tensor"][name]
    model = build_radio_tensor_model(cfg).to(dev)
    bags = {k: v.to(dev) for k, v in cases.radio_bags(cfg).items()}
    Y, c = cases.labels(cfg)
    hazards, S, Y_hat, A_raw = model(**bags)
    assert A_raw.shape == gold["A_raw"].shape == (1, 1)
    assert rel_err(A_raw, gold["A_raw"]) < 1e-2
    assert rel_err(hazards, gold["hazards"]) < 1e-2 and rel_err(S, gold["S"]) < 1e-2
    M = model(**bags, return_features=True)
    assert rel_err(M, gold["M"]) < 1e-2
    assert torch.equal(M.cpu() > 0, gold["M"] > 0), "a ReLU of the single row flipped under bf16 rounding"
    assert torch.equal(model(**bags, attention_only=True), A_raw)
    loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y.to(dev), c=c.to(dev))
    assert abs(loss.item() - gold["loss"].item()) < 1e-2 * abs(gold["loss"].item())
    model.zero_grad()
    loss.backward()
    bad = {}
    for k, p in model.named_parameters():
        fp = gold["grads"][k]
        assert p.grad is not None, k
        ref = fp["vals"]
        if ref.abs().max().item() <= 1e-6:
            assert p.grad.abs().max().item() <= 1e-4, k
            continue
        got = p.grad.detach().reshape(-1).float().cpu()[cases._sample_idx(p.numel())]
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        if err >= 3e-2:
            bad[k] = err
    assert not bad, f"gradients beyond 3e-2: {bad}"


@pytest.mark.parametrize("modname", ["coxranking_models_pretrained", "nll_models_pretrained"])
@pytest.mark.parametrize("train_type,mode,method", [
    ("kronecker", "radio_path_omic", "captum"), ("kronecker", "path_omic", "captum_path_omic"),
    ("late-fcnn", "radio_path", "captum_radio_path"), ("early-highway", "radio_omic", "captum_radio_omic"),
    ("late-highway", "radio_path_omic", "captum"), ("early-fcnn", "radio_path_omic", "captum")])
def test_pretrained_head_captum_on_gpu_equals_cpu_restatement(dev, modname, train_type, mode, method):
    """The captum* entry points of the pretrained heads on the kernels (risk and its gradient w.r.t. the embeddings, what
    IntegratedGradients consumes) == the same module run through the oracle's fp32 restatements on the CPU (which
    tests/test_dropin.py pins to the live reference)."""
    import copy
    import inspect
    from cpu_standins import oracle_kernels
    from multimodalfusion_b200 import models as M
    kw = dict(mode=mode, train_type=train_type, n_classes=4, bag_loss="nll_surv" if modname.startswith("nll") else "cox_surv")
    torch.manual_seed(3)
    cpu_model = getattr(getattr(M, modname), "multimodal_pretrained")(**kw).eval()
    gpu_model = copy.deepcopy(cpu_model).to(dev)
    g = torch.Generator().manual_seed(4)
    emb = {k: torch.randn(5, 256, generator=g) for k in ("h_radio", "h_path", "h_omic")}
    names = list(inspect.signature(getattr(cpu_model, method)).parameters)
    a_cpu = [emb[n].clone().requires_grad_() for n in names]
    a_gpu = [emb[n].to(dev).requires_grad_() for n in names]
    with oracle_kernels():
        want = getattr(cpu_model, method)(*a_cpu)
        want.sum().backward()
    got = getattr(gpu_model, method)(*a_gpu)
    got.sum().backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-5)
    for x, y in zip(a_gpu, a_cpu):
        scale = y.grad.abs().max().item() + 1e-8
        assert (x.grad.cpu() - y.grad).abs().max().item() <= 2e-4 * scale + 1e-7
