"""CPU: the Python glue of EVERY drop-in module family against the reference goldens, with the oracle's fp32 restatements
standing in for the CUDA kernels (tests/cpu_standins.py). What this pins without a GPU: constructor RNG consumption
(bit-identical initial weights), kwargs / flags (return_features, attention_only), which parameters feed which op, the
fusion order per mode, output shapes and dtypes, and — through torch autograd over the stand-ins — that every
parameter the reference trains receives the reference's gradient. Kernel numerics are the GPU tests' job."""
import pytest
import torch

from cpu_standins import oracle_kernels
from helpers import (build_head2_model, build_head_model, build_omic_model, build_path_model, build_radio_model,
                     build_radio_tensor_model, build_unimodal_model, build_xfusion4, rel_err, unimodal_input)
from oracle import amil_oracle as O
from oracle import cases

TOL = 5e-5


def _check_grads(model, gold_grads, tol=2e-4):
    for k, p in model.named_parameters():
        fp = gold_grads[k]
        if fp is None or fp["norm"] == 0.0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-7, k
            continue
        assert p.grad is not None, k
        cases.check_fingerprint(p.grad, fp, tol, f"grad {k}", atol=1e-7)


@pytest.mark.parametrize("name", list(cases.PATH_CASES))
def test_path_model_glue(goldens, name):
    cfg, gold = cases.PATH_CASES[name], goldens["path"][name]
    model = build_path_model(cfg)
    x = cases.path_bag(cfg)
    Y, c = cases.labels(cfg)
    with oracle_kernels():
        hazards, S, Y_hat, A_raw = model(path_features=x)
        M = model(path_features=x, return_features=True)
        assert torch.equal(model(path_features=x, attention_only=True), A_raw)
        loss = O.nll_surv_loss(hazards, S, Y, c, alpha=cfg["alpha"])
        model.zero_grad()
        loss.backward()
    assert A_raw.shape == gold["A_raw"].shape and Y_hat.shape == gold["Y_hat"].shape and Y_hat.dtype == torch.int64
    assert rel_err(A_raw, gold["A_raw"]) < TOL and rel_err(M, gold["M"]) < TOL
    assert rel_err(hazards, gold["hazards"]) < TOL and rel_err(S, gold["S"]) < TOL
    assert abs(loss.item() - gold["loss"].item()) < 1e-5 * max(1.0, abs(gold["loss"].item()))
    if cfg.get("peaky", 0) < 100:
        _check_grads(model, gold["grads"])


@pytest.mark.parametrize("name", list(cases.RADIO_CASES))
def test_radio_model_glue(goldens, name):
    cfg, gold = cases.RADIO_CASES[name], goldens["radio"][name]
    model = build_radio_model(cfg)
    bags = cases.radio_bags(cfg)
    Y, c = cases.labels(cfg)
    with oracle_kernels():
        hazards, S, Y_hat, A_raw = model(**bags)
        M = model(**bags, return_features=True)
        assert torch.equal(model(**bags, attention_only=True), A_raw)
        assert torch.equal(model(**bags, return_attention=True), A_raw)
        loss = O.nll_surv_loss(hazards, S, Y, c, alpha=cfg["alpha"])
        model.zero_grad()
        loss.backward()
    assert rel_err(A_raw, gold["A_raw"]) < TOL and rel_err(M, gold["M"]) < TOL
    assert rel_err(hazards, gold["hazards"]) < TOL and rel_err(S, gold["S"]) < TOL
    _check_grads(model, gold["grads"])


@pytest.mark.parametrize("name", list(cases.RADIO_TENSOR_CASES))
def test_radio_tensor_model_glue(goldens_xfusion4, name):
    cfg, gold = cases.RADIO_TENSOR_CASES[name], goldens_xfusion4["radio_tensor"][name]
    model = build_radio_tensor_model(cfg)
    bags = cases.radio_bags(cfg)
    Y, c = cases.labels(cfg)
    with oracle_kernels():
        hazards, S, Y_hat, A_raw = model(**bags)
        loss = O.nll_surv_loss(hazards, S, Y, c, alpha=cfg["alpha"])
        model.zero_grad()
        loss.backward()
    assert A_raw.shape == (1, 1) and rel_err(A_raw, gold["A_raw"]) < TOL and rel_err(hazards, gold["hazards"]) < TOL
    _check_grads(model, gold["grads"])


@pytest.mark.parametrize("name", list(cases.OMIC_CASES))
def test_snn_model_glue(goldens, name):
    cfg, gold = cases.OMIC_CASES[name], goldens["omic"][name]
    model = build_omic_model(cfg)
    x = cases.omic_batch(cfg).requires_grad_(True)
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    with oracle_kernels():
        risk = model(genomic_features=x)[0]
        feats = model(genomic_features=x, return_features=True)
        loss = O.cox_loss(risk, times, c)
        model.zero_grad()
        loss.backward()
    assert rel_err(risk, gold["risk"]) < TOL and rel_err(feats, gold["features"]) < TOL
    assert rel_err(x.grad, gold["dx"]) < 2e-4
    _check_grads(model, gold["grads"])


@pytest.mark.parametrize("bag_loss", ["cox_surv", "nll_surv"])
def test_snn_model_takes_one_patient_as_a_1d_row(bag_loss):
    """The reference's loaders hand MaxNet ONE patient's omics as a 1-D [d] tensor (the collate concatenates 1-D rows):
    nn.Linear accepts it, `logits.unsqueeze(0)` makes the [1,K] batch, the Cox risk is squeezed to a scalar
    (models/model_genomic.py:53-72). The mirror must return the same shapes and the values of the [1,d] batch; checked
    against the live reference class when its tree is mounted."""
    import os
    import sys
    from multimodalfusion_b200.models import MaxNet
    torch.manual_seed(3)
    model = MaxNet(36, bag_loss=bag_loss, n_classes=4).eval()
    x = torch.randn(36)
    with oracle_kernels():
        out1 = model(genomic_features=x)
        out2 = model(genomic_features=x.unsqueeze(0))
        f1 = model(genomic_features=x, return_features=True)
    assert f1.shape == (256,)
    if bag_loss == "nll_surv":
        assert out1[0].shape == (1, 4) and out1[1].shape == (1, 4) and out1[2].shape == (1, 1)
        assert torch.equal(out1[0], out2[0]) and torch.equal(out1[1], out2[1])
    else:
        assert out1[0].dim() == 0 and torch.equal(out1[0], out2[0])
    ref_root = os.environ.get("MMF_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "models")):
        import importlib
        sys.path.insert(0, ref_root)
        try:
            torch.cuda.FloatTensor = torch.FloatTensor
            saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
            ref_cls = importlib.import_module("models.model_genomic").MaxNet
            torch.manual_seed(3)
            ref = ref_cls(36, bag_loss=bag_loss, n_classes=4).eval()
            ref.load_state_dict(model.state_dict())
            r = ref(genomic_features=x)
            for a, b in zip(out1[:3], r[:3]):
                if b is not None:
                    assert a.shape == b.shape and rel_err(a.float(), b.float()) < TOL
            assert rel_err(f1, ref(genomic_features=x, return_features=True)) < TOL
        finally:
            sys.path.remove(ref_root)
            for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
                sys.modules.pop(k)
            sys.modules.update(saved)


def _head_case(model, cfg, gold):
    hr, hp, ho = [t.requires_grad_(True) for t in cases.embeddings(cfg)]
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    with oracle_kernels():
        res = model(hr, hp, ho)
        if cfg["kind"] == "cox":
            risk = res[0]
            assert res[1] is None and res[2] is None
            loss = (O.cox_loss(risk.reshape(-1), times, c) if cfg["loss"] == "cox"
                    else O.ranking_loss(risk.reshape(-1), times, c))
        else:
            risk, hazards, S = res
            assert rel_err(hazards, gold["hazards"]) < TOL and rel_err(S, gold["S"]) < TOL
            Y = torch.arange(cfg["B"]) % 4
            loss = (O.nll_surv_loss(hazards, S, Y, c, alpha=0.15) if cfg["loss"] == "nll"
                    else O.ce_surv_loss(hazards, S, Y, c, alpha=0.15))
        model.zero_grad()
        loss.backward()
    assert risk.shape == gold["risk"].shape and rel_err(risk, gold["risk"]) < TOL
    assert abs(loss.item() - gold["loss"].item()) < 2e-5
    for t, gd in zip((hr, hp, ho), gold["d_inputs"]):
        if gd is not None:
            assert rel_err(t.grad, gd) < 2e-4
    _check_grads(model, gold["grads"])


@pytest.mark.parametrize("name", list(cases.HEAD_CASES))
def test_kronecker_head_glue(goldens, name):
    _head_case(build_head_model(cases.HEAD_CASES[name]), cases.HEAD_CASES[name], goldens["heads"][name])


@pytest.mark.parametrize("name", list(cases.HEAD2_CASES))
def test_fcnn_highway_head_glue(goldens_heads2, name):
    _head_case(build_head2_model(cases.HEAD2_CASES[name]), cases.HEAD2_CASES[name], goldens_heads2["heads2"][name])


@pytest.mark.parametrize("name", list(cases.XFUSION4_CASES))
def test_xfusion4_glue(goldens_xfusion4, name):
    cfg, gold = cases.XFUSION4_CASES[name], goldens_xfusion4["xfusion4"][name]
    model = build_xfusion4(cfg)
    vs, proj = cases.embeddings4(cfg)
    vs = [v.requires_grad_(True) for v in vs]
    with oracle_kernels():
        feats = model(v_list=vs)
        model.zero_grad()
        (feats * proj).sum().backward()
    assert rel_err(feats, gold["features"]) < TOL
    for v, gd in zip(vs, gold["d_inputs"]):
        assert rel_err(v.grad, gd) < 2e-4
    _check_grads(model, gold["grads"])


@pytest.mark.parametrize("name", list(cases.UNI_CASES))
def test_unimodal_head_glue(goldens_unimodal, name):
    """unimonal_pretrained (fcnn / highway / residual) of both head files: identical initial weights, outputs, loss, input
    gradient and every parameter gradient against the reference."""
    cfg, gold = cases.UNI_CASES[name], goldens_unimodal["unimodal"][name]
    model = build_unimodal_model(cfg)
    assert list(model.state_dict()) == list(gold["weights_fp"])
    for k, v in model.state_dict().items():
        cases.check_fingerprint(v, gold["weights_fp"][k], 0.0, f"weight {k}")
    h = unimodal_input(cfg).requires_grad_(True)
    times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
    with oracle_kernels():
        res = model(**{"h_" + cfg["mode"]: h})
        if cfg["kind"] == "cox":
            risk = res[0]
            assert res[1] is None and res[2] is None
            loss = O.cox_loss(risk, times, c) if cfg["loss"] == "cox" else O.ranking_loss(risk.reshape(-1), times, c)
        else:
            risk, hazards, S = res
            assert rel_err(hazards, gold["hazards"]) < TOL and rel_err(S, gold["S"]) < TOL
            Y = torch.arange(cfg["B"]) % 4
            loss = (O.nll_surv_loss(hazards, S, Y, c, alpha=0.15) if cfg["loss"] == "nll"
                    else O.ce_surv_loss(hazards, S, Y, c, alpha=0.15))
        model.zero_grad()
        loss.backward()
    assert risk.shape == gold["risk"].shape and rel_err(risk, gold["risk"]) < TOL
    assert abs(loss.item() - gold["loss"].item()) < 2e-5
    assert rel_err(h.grad, gold["d_input"]) < 2e-4
    _check_grads(model, gold["grads"])
