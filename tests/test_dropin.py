"""CPU: drop-in surface — constructors, state_dict keys/shapes/values under a seed, forward
signature, relocate(), loud failure on CPU inputs. Compared against the live reference when
/root/reference is mounted, and against the committed golden fingerprints always."""
import inspect
import os
import sys

import pytest
import torch

from helpers import build_head_model, build_omic_model, build_path_model, build_radio_model
from multimodalfusion_b200 import models as M
from oracle import cases

REF = os.environ.get("MMF_REFERENCE", "/root/reference")
HAVE_REF = os.path.isdir(os.path.join(REF, "models"))


@pytest.fixture(scope="module")
def ref():
    if not HAVE_REF:
        pytest.skip("reference tree not mounted")
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor
    import importlib
    mods = {n: importlib.import_module(f"models.{n}") for n in
            ("model_modules", "model_attention_mil_path", "model_attention_mil_radio", "model_genomic",
             "coxranking_models_pretrained", "nll_models_pretrained")}
    yield mods
    sys.path.remove(REF)
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]


CTOR_MATRIX = [
    ("model_attention_mil_path", "MIL_Attention_fc_surv_path", dict()),
    ("model_attention_mil_path", "MIL_Attention_fc_surv_path", dict(gate_path=False, dropout=True, n_classes=8)),
    ("model_attention_mil_path", "MIL_Attention_fc_surv_path", dict(gate_path=False, dropout=False, model_size_wsi="big")),
    ("model_attention_mil_path", "MIL_Attention_fc_surv_path", dict(model_size_wsi="big", dropout=True)),
    ("model_attention_mil_radio", "MIL_Attention_fc_surv_radio", dict()),
    ("model_attention_mil_radio", "MIL_Attention_fc_surv_radio", dict(gate_radio=False, dropout=False, modalities=["T1", "T2"])),
    ("model_genomic", "MaxNet", dict(input_dim=36, bag_loss="cox_surv")),
    ("model_genomic", "MaxNet", dict(input_dim=186, model_size_omic="big", bag_loss="nll_surv", n_classes=8)),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="kronecker", mode="radio_path_omic")),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="kronecker", mode="path_omic")),
    ("nll_models_pretrained", "multimodal_pretrained", dict(train_type="kronecker", mode="radio_path", n_classes=8)),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="late-fcnn", mode="radio_path_omic")),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="early-fcnn", mode="radio_path")),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="early-highway", mode="path_omic", n_layers=2)),
    ("coxranking_models_pretrained", "multimodal_pretrained", dict(train_type="late-highway", mode="radio_path_omic")),
    ("nll_models_pretrained", "multimodal_pretrained", dict(train_type="late-fcnn", mode="radio_path_omic", n_classes=4)),
    ("nll_models_pretrained", "multimodal_pretrained", dict(train_type="early-fcnn", mode="radio_omic", n_classes=8)),
    ("nll_models_pretrained", "multimodal_pretrained", dict(train_type="early-highway", mode="radio_path_omic")),
    ("nll_models_pretrained", "multimodal_pretrained", dict(train_type="late-highway", mode="path_omic", n_layers=2)),
    ("model_modules", "Highway", dict(size=256, num_layers=2, f=torch.relu)),
    ("model_modules", "XlinearFusion", dict(num_modalities=3, dim=256, scale_dim=16, mmhid1=512, mmhid2=512)),
    ("model_modules", "Attn_Net_Gated", dict(L=512, D=384, dropout=True, n_classes=1)),
    ("model_modules", "Attn_Net", dict(L=256, D=256, dropout=True, n_classes=1)),
]


def _ours(modname, cls):
    import importlib
    return getattr(importlib.import_module(f"multimodalfusion_b200.models.{modname}"), cls)


@pytest.mark.parametrize("modname,cls,kwargs", CTOR_MATRIX)
def test_state_dict_identical_to_reference_under_seed(ref, modname, cls, kwargs):
    torch.manual_seed(123)
    theirs = getattr(ref[modname], cls)(**kwargs)
    torch.manual_seed(123)
    ours = _ours(modname, cls)(**kwargs)
    sd_t, sd_o = theirs.state_dict(), ours.state_dict()
    assert list(sd_t.keys()) == list(sd_o.keys())
    for k in sd_t:
        assert sd_t[k].shape == sd_o[k].shape, k
        assert torch.equal(sd_t[k], sd_o[k]), f"{k}: seeded init differs"
    ours.load_state_dict(sd_t, strict=True)
    theirs.load_state_dict(sd_o, strict=True)
    # same constructor signature
    assert (str(inspect.signature(getattr(ref[modname], cls).__init__))
            == str(inspect.signature(_ours(modname, cls).__init__)))


def test_golden_weight_fingerprints(goldens):
    """Works without the reference: seeded construction reproduces the reference weights recorded in
    tests/golden (keys, shapes, sampled values)."""
    for name, cfg in cases.PATH_CASES.items():
        sd = build_path_model(cfg).state_dict()
        fp = goldens["path"][name]["weights_fp"]
        assert list(sd.keys()) == list(fp.keys())
        for k, v in sd.items():
            cases.check_fingerprint(v, fp[k], 0.0, f"{name}:{k}")
    for name, cfg in cases.RADIO_CASES.items():
        sd = build_radio_model(cfg).state_dict()
        assert list(sd.keys()) == list(goldens["radio"][name]["weights_fp"].keys())
    for name, cfg in cases.OMIC_CASES.items():
        sd = build_omic_model(cfg).state_dict()
        assert list(sd.keys()) == list(goldens["omic"][name]["weights_fp"].keys())
    for name, cfg in cases.HEAD_CASES.items():
        sd = build_head_model(cfg).state_dict()
        fp = goldens["heads"][name]["weights_fp"]
        assert list(sd.keys()) == list(fp.keys())
        for k, v in sd.items():
            cases.check_fingerprint(v, fp[k], 0.0, f"{name}:{k}")


def test_param_counts_match_survey():
    assert sum(p.numel() for p in M.MIL_Attention_fc_surv_path().parameters()) == 395_269
    assert sum(p.numel() for p in M.MIL_Attention_fc_surv_path(model_size_wsi="big", n_classes=8).parameters()) == 923_273


def test_mm_model_repairs_the_reference_constructor():
    """The reference class raises at construction; the mirror accepts the same
    call and exposes the state_dict keys the reference code defines."""
    m = M.MM_MIL_Attention_fc_surv(input_dim=36, gate_omic=True, n_classes=4)
    keys = list(m.state_dict().keys())
    for prefix in ("fc_omic.0.0.", "attention_net_radio.0.", "attention_net_radio.3.attention_a.0.", "reduce_dim.",
                   "attention_net_WSI.3.attention_c.", "mm.reduce.0.0.0.", "mm.reduce.2.2.0.", "mm.encoder1.0.",
                   "mm.encoder2.0.", "classifier.0.", "classifier.3."):
        assert any(k.startswith(prefix) for k in keys), prefix
    assert m.mm.encoder1[0].weight.shape == (512, 17 ** 3)
    m2 = M.MM_MIL_Attention_fc_surv(fusion="concat", mode="path_omic")
    assert m2.classifier.weight.shape == (4, 512)


def test_forward_rejects_cpu_inputs_loudly():
    model = M.MIL_Attention_fc_surv_path()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(path_features=torch.randn(4, 1024))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.MaxNet(36, bag_loss="cox_surv")(genomic_features=torch.randn(2, 36))
    from multimodalfusion_b200.utils import CoxSurvLoss, NLLSurvLoss
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CoxSurvLoss()(risks=torch.randn(4), times=torch.rand(4), c=torch.zeros(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NLLSurvLoss()(hazards=torch.rand(1, 4), S=torch.rand(1, 4), Y=torch.tensor([1]), c=torch.tensor([0.0]))


def test_unsupported_modes_raise_like_the_reference():
    from multimodalfusion_b200.utils import ranking_loss
    with pytest.raises(NotImplementedError):
        ranking_loss(torch.zeros(1), torch.zeros(1), torch.zeros(1), "sigmoid", "mean")
    with pytest.raises(NotImplementedError):
        from multimodalfusion_b200.models.coxranking_models_pretrained import multimodal_pretrained
        multimodal_pretrained(train_type="early-residual")   # commented out in the reference as well
    from multimodalfusion_b200.models import coxranking_models_pretrained as cox_heads
    keys = list(cox_heads.unimonal_pretrained(train_type="residual", n_layers=1).state_dict())
    assert "residual.blocks.0.fc1.weight" in keys and "residual.blocks.0.bn2.running_var" in keys
    assert hasattr(M.MIL_Attention_fc_surv_path(), "relocate")


def test_product_package_never_imports_the_oracle():
    import pathlib
    pkg = pathlib.Path(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) / "multimodalfusion_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_l1_penalty_and_freeze_helpers_match_the_reference_definitions():
    """l1_reg_all / l1_reg_modules / dfs_freeze / dfs_unfreeze (utils/utils.py:235-269): value and gradient of the
    fused multi-tensor norm equal the reference's per-parameter abs().sum() chain."""
    from multimodalfusion_b200.utils import dfs_freeze, dfs_unfreeze, l1_reg_all, l1_reg_modules
    torch.manual_seed(0)
    m = M.MaxNet(36, bag_loss="cox_surv")
    for p in m.parameters():
        p.data.add_(0.01 * torch.randn_like(p))
    ours = l1_reg_all(m)
    ours.backward()
    g_ours = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    ref = None
    for W in m.parameters():                       # the reference's loop, verbatim semantics
        ref = torch.abs(W).sum() if ref is None else ref + torch.abs(W).sum()
    ref.backward()
    assert abs(ours.item() - ref.item()) < 1e-5 * ref.item()
    for a, p in zip(g_ours, m.parameters()):
        assert torch.equal(a, p.grad)
    assert abs(l1_reg_modules(m).item() - l1_reg_all(m.fc_omic).item()) < 1e-3     # MaxNet has no `mm` block
    dfs_freeze(m)
    assert not any(p.requires_grad for p in m.parameters())
    dfs_unfreeze(m)
    assert all(p.requires_grad for p in m.parameters())


CAPTUM_MATRIX = [(f, tt, mode, meth) for f in ("coxranking_models_pretrained", "nll_models_pretrained")
                 for tt in ("kronecker", "early-fcnn", "late-fcnn", "early-highway", "late-highway")
                 for mode, meth in (("radio_path_omic", "captum"), ("radio_path", "captum_radio_path"),
                                    ("path_omic", "captum_path_omic"), ("radio_omic", "captum_radio_omic"))]


@pytest.mark.parametrize("modname,train_type,mode,method", CAPTUM_MATRIX)
def test_pretrained_head_captum_entry_points_match_live_reference(ref, modname, train_type, mode, method):
    """multimodal_pretrained.captum / captum_radio_path / captum_path_omic / captum_radio_omic (the entry points
    create_attributions.py calls, models/*_models_pretrained.py:200-318) against the live reference in eval mode: the same
    seeded construction, the oracle's fp32 restatements standing in for the kernels (host glue: orders, reshapes)."""
    from cpu_standins import oracle_kernels
    kw = dict(mode=mode, train_type=train_type, n_classes=4, bag_loss="nll_surv" if modname.startswith("nll") else "cox_surv")
    torch.manual_seed(5)
    theirs = getattr(ref[modname], "multimodal_pretrained")(**kw).eval()
    torch.manual_seed(5)
    ours = getattr(getattr(M, modname), "multimodal_pretrained")(**kw).eval()
    ours.load_state_dict(theirs.state_dict(), strict=True)
    g = torch.Generator().manual_seed(9)
    emb = {k: torch.randn(6, 256, generator=g) for k in ("h_radio", "h_path", "h_omic")}
    names = list(inspect.signature(getattr(theirs, method)).parameters)
    assert list(inspect.signature(getattr(ours, method)).parameters) == names
    args = [emb[n] for n in names]
    want = getattr(theirs, method)(*args)
    with oracle_kernels():
        got = getattr(ours, method)(*args)
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-6)
