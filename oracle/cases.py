"""Seeded parity cases shared by oracle/make_goldens.py (reference side, build container) and the
tests (oracle / CUDA side, anywhere). TEST INFRASTRUCTURE ONLY.

Inputs follow SURVEY.md §8(d): pathology / radiology features are non-negative (ResNet post-avgpool)
``0.5*|N(0,1)|`` rounded ONCE to bf16 and handed to the fp32 reference as the same rounded values;
omics are z-scored; times ~ Exp(30 months) clipped to [0, 250] with deliberate ties;
censoring ~ Bernoulli(0.46).
"""
from __future__ import annotations

import torch

# name -> config.  N covers 1, tile tails (127/128/129), multi-tile; both presets; both attention nets.
PATH_CASES = {
    "small_gated_n200": dict(seed=11, size="small", gated=True, dropout=False, K=4, N=200, Y=2, c=0.0, alpha=0.0),
    "small_gated_n1": dict(seed=12, size="small", gated=True, dropout=False, K=4, N=1, Y=0, c=1.0, alpha=0.15),
    "small_gated_n7_drop": dict(seed=13, size="small", gated=True, dropout=True, K=4, N=7, Y=3, c=0.0, alpha=0.4),
    "small_ungated_n129": dict(seed=14, size="small", gated=False, dropout=False, K=4, N=129, Y=1, c=0.0, alpha=0.0),
    "small_ungated_n128_drop": dict(seed=15, size="small", gated=False, dropout=True, K=4, N=128, Y=2, c=1.0, alpha=0.15),
    "big_gated_n300_k8": dict(seed=16, size="big", gated=True, dropout=False, K=8, N=300, Y=5, c=0.0, alpha=0.0),
    "big_ungated_n127": dict(seed=17, size="big", gated=False, dropout=False, K=4, N=127, Y=0, c=0.0, alpha=0.0),
    "small_gated_n1000_peaky": dict(seed=18, size="small", gated=True, dropout=False, K=4, N=1000, Y=2, c=0.0,
                                    alpha=0.0, peaky=20.0),
    "big_gated_n1000_extreme": dict(seed=19, size="big", gated=True, dropout=False, K=4, N=1000, Y=1, c=0.0,
                                    alpha=0.0, peaky=150.0),
}
RADIO_CASES = {
    "radio_gated_n37": dict(seed=21, gated=True, dropout=True, K=4, N=37, Y=1, c=0.0, alpha=0.0),
    "radio_gated_n155": dict(seed=22, gated=True, dropout=True, K=4, N=155, Y=3, c=1.0, alpha=0.15),
    "radio_ungated_n80": dict(seed=23, gated=False, dropout=False, K=4, N=80, Y=0, c=0.0, alpha=0.0),
}
OMIC_CASES = {
    "omic_brain36_b5": dict(seed=31, d_in=36, B=5, bag_loss="cox_surv"),
    "omic_lung186_b64": dict(seed=32, d_in=186, B=64, bag_loss="cox_surv"),
}
HEAD_CASES = {
    "kron3_cox_b6": dict(seed=41, kind="cox", loss="cox", mode="radio_path_omic", B=6),
    "kron3_rank_b32": dict(seed=42, kind="cox", loss="ranking", mode="radio_path_omic", B=32),
    "kron2_cox_b9": dict(seed=43, kind="cox", loss="cox", mode="radio_path", B=9),
    "kron3_nll_b8": dict(seed=44, kind="nll", loss="nll", mode="radio_path_omic", B=8),
    "kron2_nll_b70": dict(seed=45, kind="nll", loss="nll", mode="path_omic", B=70),
}
# fcnn / Highway fusion heads: eval mode, BatchNorm running statistics and affine parameters
# perturbed (perturb_bn) so that the normalisation is not the identity
HEAD2_CASES = {
    "cox_late_fcnn_b12": dict(seed=71, kind="cox", loss="cox", mode="radio_path_omic", train_type="late-fcnn", B=12, n_layers=1),
    "cox_early_fcnn_b33": dict(seed=72, kind="cox", loss="ranking", mode="radio_path", train_type="early-fcnn", B=33, n_layers=1),
    "cox_early_highway_b9": dict(seed=73, kind="cox", loss="cox", mode="path_omic", train_type="early-highway", B=9, n_layers=2),
    "cox_late_highway_b20": dict(seed=74, kind="cox", loss="cox", mode="radio_path_omic", train_type="late-highway", B=20, n_layers=1),
    "nll_late_fcnn_b16": dict(seed=75, kind="nll", loss="nll", mode="radio_path_omic", train_type="late-fcnn", B=16, n_layers=1),
    "nll_early_fcnn_b8": dict(seed=76, kind="nll", loss="ce", mode="radio_omic", train_type="early-fcnn", B=8, n_layers=1),
    "nll_early_highway_b40": dict(seed=77, kind="nll", loss="nll", mode="radio_path_omic", train_type="early-highway", B=40, n_layers=1),
    "nll_late_highway_b6": dict(seed=78, kind="nll", loss="ce", mode="path_omic", train_type="late-highway", B=6, n_layers=2),
}
# XlinearFusion with the reference's DEFAULT num_modalities=4 (17^4 = 83 521-wide outer product)
XFUSION4_CASES = {
    "xfusion4_b5": dict(seed=81, B=5),
    "xfusion4_b33": dict(seed=82, B=33),
}
# radio_fusion='tensor' with the one-name repair of SURVEY.md App. B-3 (`model.xfusion = model.radio_xfusion`):
# slice 0 of each modality -> 4-way Kronecker fusion (dim 1024, scale 64 -> 17^4) -> a ONE-row bag -> AMIL -> head.
# The seed is chosen so that no fc pre-activation of that single row sits within bf16 rounding of zero (the golden
# generator prints the margin): one flipped ReLU would move a whole row of dW1.
RADIO_TENSOR_CASES = {
    "radio_tensor_n12": dict(seed=100, gated=True, dropout=True, K=4, N=12, Y=1, c=0.0, alpha=0.0),
}
# MM_MIL_Attention_fc_surv end to end against the reference made runnable without editing it
# (oracle/make_goldens_mm.py). Nr / Np = radiology / pathology bag sizes, d = omics width.
MM_CASES = {
    "mm_tensor_rpo": dict(seed=111, mode="radio_path_omic", fusion="tensor", Nr=23, Np=300, d=36, Y=2, c=0.0, alpha=0.0),
    "mm_concat_rpo": dict(seed=112, mode="radio_path_omic", fusion="concat", Nr=40, Np=157, d=36, Y=0, c=1.0, alpha=0.15),
    "mm_tensor_po": dict(seed=113, mode="path_omic", fusion="tensor", Nr=0, Np=260, d=80, Y=3, c=0.0, alpha=0.0),
    "mm_tensor_ro": dict(seed=114, mode="radio_omic", fusion="tensor", Nr=31, Np=0, d=36, Y=1, c=0.0, alpha=0.0),
    "mm_tensor_rp": dict(seed=115, mode="radio_path", fusion="tensor", Nr=17, Np=129, d=36, Y=1, c=1.0, alpha=0.4),
}
# the captum* entry points: batched 3-D bags [B, N, 1024] (sum pooling as written in the reference)
CAPTUM_CASES = {
    "captum_rpo": dict(seed=121, fn="captum", mode="radio_path_omic", fusion="tensor", B=3, Nr=9, Np=40, d=36),
    "captum_ro": dict(seed=122, fn="captum_radio_omic", mode="radio_omic", fusion="tensor", B=4, Nr=12, Np=0, d=36),
    "captum_rp": dict(seed=123, fn="captum_radio_path", mode="radio_path", fusion="tensor", B=2, Nr=7, Np=33, d=36),
    "captum_po_concat": dict(seed=124, fn="captum_path_omic", mode="path_omic", fusion="concat", B=5, Nr=0, Np=21, d=80),
}
# unimonal_pretrained heads on ONE modality's 256-d embedding (cox: fcnn / highway / residual; nll: fcnn / highway)
UNI_CASES = {
    "uni_cox_fcnn_path_b12": dict(seed=131, kind="cox", loss="cox", mode="path", train_type="fcnn", B=12, n_layers=1),
    "uni_cox_highway_radio_b9": dict(seed=132, kind="cox", loss="ranking", mode="radio", train_type="highway", B=9, n_layers=2),
    "uni_cox_residual_omic_b10": dict(seed=133, kind="cox", loss="cox", mode="omic", train_type="residual", B=10, n_layers=2),
    "uni_nll_fcnn_omic_b8": dict(seed=134, kind="nll", loss="nll", mode="omic", train_type="fcnn", B=8, n_layers=1),
    "uni_nll_highway_path_b16": dict(seed=135, kind="nll", loss="ce", mode="path", train_type="highway", B=16, n_layers=1),
}
LOSS_CASES = {
    "nll_b7_a0": dict(seed=51, loss="nll", B=7, K=4, alpha=0.0),
    "nll_b64_k8": dict(seed=52, loss="nll", B=64, K=8, alpha=0.15),
    "nll_b5_clamp": dict(seed=53, loss="nll", B=5, K=4, alpha=0.4, saturate=True),
    "cox_b64_ties": dict(seed=54, loss="cox", B=64),
    "cox_b2": dict(seed=55, loss="cox", B=2),
    "cox_b33_allcens": dict(seed=56, loss="cox", B=33, all_censored=True),
    "cox_b200": dict(seed=57, loss="cox", B=200),
    "rank_b33_sig_mean": dict(seed=58, loss="ranking", B=33, phi="sigmoid", reduction="mean"),
    "rank_b40_relu_sum": dict(seed=59, loss="ranking", B=40, phi="relu", reduction="sum"),
    "rank_b8_nopairs": dict(seed=60, loss="ranking", B=8, phi="sigmoid", reduction="mean", all_censored=True),
    "rank_b2": dict(seed=61, loss="ranking", B=2, phi="sigmoid", reduction="mean"),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(1_000_003 * seed + 17)
    return g


def bf16_values(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def features(n: int, seed: int, width: int = 1024) -> torch.Tensor:
    return bf16_values(0.5 * torch.randn(n, width, generator=_gen(seed)).abs())


def path_bag(cfg) -> torch.Tensor:
    return features(cfg["N"], cfg["seed"])


def radio_bags(cfg):
    return {m: features(cfg["N"], cfg["seed"] * 10 + i) for i, m in enumerate(["T1", "T2", "T1Gd", "FLAIR"])}


def labels(cfg):
    return torch.tensor([cfg["Y"]], dtype=torch.long), torch.tensor([cfg["c"]], dtype=torch.float32)


def omic_batch(cfg) -> torch.Tensor:
    return torch.randn(cfg["B"], cfg["d_in"], generator=_gen(cfg["seed"]))


def cohort_labels(B: int, seed: int, all_censored: bool = False):
    g = _gen(seed + 7)
    times = torch.empty(B).exponential_(1.0 / 30.0, generator=g).clamp_(0, 250)
    times = (times * 2).round() / 2            # half-month grid -> natural ties
    if B >= 4:
        times[B // 2] = times[0]               # and forced ones
        times[B - 1] = times[1]
    c = (torch.rand(B, generator=g) < 0.46).float()
    if all_censored:
        c = torch.ones(B)
    return times, c


def embeddings(cfg):
    g = _gen(cfg["seed"] + 3)
    B = cfg["B"]
    # path/radio embeddings are post-ReLU pooled features (non-negative); omics are SELU outputs
    hr = torch.randn(B, 256, generator=g).abs() * 0.5
    hp = torch.randn(B, 256, generator=g).abs() * 0.5
    ho = torch.randn(B, 256, generator=g)
    return hr, hp, ho


def embeddings4(cfg):
    """Four 256-d embeddings (radio, path, omic + a second non-negative pooled one) and the fixed projection that turns
    the fused features into a scalar objective for the gradient goldens."""
    hr, hp, ho = embeddings(cfg)
    g = _gen(cfg["seed"] + 4)
    h4 = torch.randn(cfg["B"], 256, generator=g).abs() * 0.5
    proj = torch.randn(256, generator=g) / 16.0
    return [hr, hp, ho, h4], proj


def mm_inputs(cfg):
    """kwargs of MM_MIL_Attention_fc_surv.forward for one patient."""
    kw = {}
    if "radio" in cfg["mode"]:
        kw.update(radio_bags(dict(N=cfg["Nr"], seed=cfg["seed"])))
    if "path" in cfg["mode"]:
        kw["path_features"] = features(cfg["Np"], cfg["seed"] + 7)
    if "omic" in cfg["mode"]:
        kw["genomic_features"] = torch.randn(cfg["d"], generator=_gen(cfg["seed"] + 8))
    return kw


def captum_inputs(cfg):
    """Positional inputs of the captum entry point named cfg['fn'] (batched bags are scaled down: the reference pools
    by SUMMING the instances) and the weights of the scalar objective sum(risk * w)."""
    g = _gen(cfg["seed"] + 9)
    B = cfg["B"]
    radio = [0.06 * torch.randn(B, cfg["Nr"], 1024, generator=g).abs() for _ in range(4)] if cfg["Nr"] else []
    path = 0.03 * torch.randn(B, cfg["Np"], 1024, generator=g).abs() if cfg["Np"] else None
    omic = torch.randn(B, cfg["d"], generator=g)
    w = torch.arange(1, B + 1, dtype=torch.float32)
    args = {"captum": radio + [path, omic], "captum_radio_omic": radio + [omic], "captum_radio_path": radio + [path],
            "captum_path_omic": [omic, path]}[cfg["fn"]]
    return args, w


def nll_inputs(cfg):
    g = _gen(cfg["seed"])
    B, K = cfg["B"], cfg["K"]
    hz = torch.sigmoid(torch.randn(B, K, generator=g))
    if cfg.get("saturate"):
        hz[0, :] = 1.0 - 1e-9      # S underflows below eps -> clamp branch
        hz[1, :] = 1e-9            # hazards below eps -> clamp branch
    S = torch.cumprod(1 - hz, dim=1)
    Y = torch.randint(0, K, (B,), generator=g)
    c = (torch.rand(B, generator=g) < 0.46).float()
    return hz, S, Y, c


def risk_inputs(cfg):
    g = _gen(cfg["seed"])
    r = torch.randn(cfg["B"], generator=g)
    times, c = cohort_labels(cfg["B"], cfg["seed"], cfg.get("all_censored", False))
    return r, times, c


def perturb_biases(model: torch.nn.Module, seed: int) -> None:
    """The reference initialises every bias to zero; give them seeded non-zero values so that bias
    handling is exercised. Applied identically on the reference and on the drop-in side."""
    g = _gen(seed + 101)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


def perturb_bn(model: torch.nn.Module, seed: int) -> None:
    """Seeded non-trivial BatchNorm1d state (running mean / var, weight, bias); identical on both sides."""
    g = _gen(seed + 211)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))


def make_peaky(model: torch.nn.Module, factor: float) -> None:
    """Scales the final attention projection so the softmax over the bag becomes (near) one-hot and
    raw scores reach tens to hundreds — exercises the online-softmax rescale."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("attention_c.weight") or ".module.2.weight" in name or ".module.3.weight" in name:
                p.mul_(factor)


def _sample_idx(numel: int, k: int = 1024) -> torch.Tensor:
    g = torch.Generator()
    g.manual_seed(numel * 7919 + 13)
    if numel <= k:
        return torch.arange(numel)
    return torch.randperm(numel, generator=g)[:k].clone()


def fingerprint(t: torch.Tensor) -> dict:
    flat = t.detach().reshape(-1).to(torch.float32).cpu()
    idx = _sample_idx(flat.numel())
    return {"shape": tuple(t.shape), "vals": flat[idx].clone(), "sum": flat.double().sum().item(),
            "norm": flat.double().norm().item()}


def fingerprint_state(sd) -> dict:
    return {k: fingerprint(v) for k, v in sd.items()}


def check_fingerprint(t: torch.Tensor, fp: dict, rtol: float, what: str = "", atol: float = 0.0) -> None:
    """|t - ref| <= rtol * max|ref| + atol on the sampled entries, plus norm agreement."""
    flat = t.detach().reshape(-1).to(torch.float32).cpu()
    assert tuple(t.shape) == tuple(fp["shape"]), f"{what}: shape {tuple(t.shape)} vs {fp['shape']}"
    ref = fp["vals"]
    scale = max(ref.abs().max().item(), 1e-30)
    err = (flat[_sample_idx(flat.numel())] - ref).abs().max().item()
    assert err <= rtol * scale + atol + 1e-12, f"{what}: sampled max err {err:.3e} vs tol {rtol * scale + atol:.3e}"
    n_err = abs(flat.double().norm().item() - fp["norm"])
    assert n_err <= rtol * max(fp["norm"], 1e-30) + atol * flat.numel() ** 0.5 + 1e-12, f"{what}: norm {n_err:.3e}"
