"""Generates tests/golden/reference_goldens.pt by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; it is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens.py

The reference is imported from where it lies (nothing is copied); the one shim applied is
``torch.cuda.FloatTensor = torch.FloatTensor`` so that XlinearFusion.forward
(models/model_modules.py:164) runs on CPU. Inputs are produced by ``oracle/cases.py`` from fixed
seeds, so tests regenerate them bit-identically instead of storing them; weights are regenerated
by constructing the model under the recorded seed (tests check a weight fingerprint). Large
tensors (weight gradients) are stored as a fingerprint: 1024 sampled entries (indices regenerated from the element count) + sum + L2 norm.
"""
from __future__ import annotations

import os
import sys

import torch

REF = os.environ.get("MMF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

from oracle import cases  # noqa: E402


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor  # CPU shim for model_modules.py:164
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    from models.model_attention_mil_path import MIL_Attention_fc_surv_path
    from models.model_attention_mil_radio import MIL_Attention_fc_surv_radio
    from models.model_genomic import MaxNet
    from models import coxranking_models_pretrained as cox_heads
    from models import nll_models_pretrained as nll_heads
    from utils.loss_utils import CoxSurvLoss, NLLSurvLoss, RankingSurvLoss

    out = {"torch": torch.__version__, "path": {}, "radio": {}, "omic": {}, "heads": {}, "losses": {}}

    # ---- pathology AMIL -------------------------------------------------------------------
    for name, cfg in cases.PATH_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = MIL_Attention_fc_surv_path(gate_path=cfg["gated"], model_size_wsi=cfg["size"],
                                           dropout=cfg["dropout"], n_classes=cfg["K"]).eval()
        cases.perturb_biases(model, cfg["seed"])
        if cfg.get("peaky"):
            cases.make_peaky(model, cfg["peaky"])
        x = cases.path_bag(cfg)
        Y, c = cases.labels(cfg)
        hazards, S, Y_hat, A_raw = model(path_features=x)
        M = model(path_features=x, return_features=True)
        loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y, c=c)
        model.zero_grad()
        loss.backward()
        out["path"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()),
            "A_raw": A_raw.detach().clone(), "M": M.detach().clone(), "hazards": hazards.detach().clone(),
            "S": S.detach().clone(), "Y_hat": Y_hat.clone(), "loss": loss.detach().clone(),
            "grads": {k: cases.fingerprint(p.grad) for k, p in model.named_parameters()},
        }
        print("path", name, float(loss))

    # ---- radiology AMIL -------------------------------------------------------------------
    for name, cfg in cases.RADIO_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = MIL_Attention_fc_surv_radio(gate_radio=cfg["gated"], dropout=cfg["dropout"],
                                            n_classes=cfg["K"]).eval()
        cases.perturb_biases(model, cfg["seed"])
        bags = cases.radio_bags(cfg)
        Y, c = cases.labels(cfg)
        hazards, S, Y_hat, A_raw = model(**bags)
        M = model(**bags, return_features=True)
        loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y, c=c)
        model.zero_grad()
        loss.backward()
        out["radio"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()),
            "A_raw": A_raw.detach().clone(), "M": M.detach().clone(), "hazards": hazards.detach().clone(),
            "S": S.detach().clone(), "loss": loss.detach().clone(),
            "grads": {k: cases.fingerprint(p.grad) for k, p in model.named_parameters()},
        }
        print("radio", name, float(loss))

    # ---- genomic SNN ----------------------------------------------------------------------
    for name, cfg in cases.OMIC_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = MaxNet(cfg["d_in"], bag_loss=cfg["bag_loss"], n_classes=4).eval()
        cases.perturb_biases(model, cfg["seed"])
        x = cases.omic_batch(cfg).requires_grad_(True)
        times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
        risk = model(genomic_features=x)[0]
        feats = model(genomic_features=x, return_features=True)
        loss = CoxSurvLoss()(risks=risk, times=times, c=c)
        model.zero_grad()
        loss.backward()
        out["omic"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()),
            "risk": risk.detach().clone(), "features": feats.detach().clone(), "loss": loss.detach().clone(),
            "dx": x.grad.clone(),
            "grads": {k: cases.fingerprint(p.grad) for k, p in model.named_parameters()},
        }
        print("omic", name, float(loss))

    # ---- kronecker heads on 256-d embeddings ----------------------------------------------
    for name, cfg in cases.HEAD_CASES.items():
        torch.manual_seed(cfg["seed"])
        mod = cox_heads if cfg["kind"] == "cox" else nll_heads
        model = mod.multimodal_pretrained(mode=cfg["mode"], train_type="kronecker", n_classes=4).eval()
        cases.perturb_biases(model, cfg["seed"])
        hr, hp, ho = [t.requires_grad_(True) for t in cases.embeddings(cfg)]
        times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
        res = model(hr, hp, ho)
        if cfg["kind"] == "cox":
            risk = res[0]
            loss = (CoxSurvLoss()(risks=risk, times=times, c=c) if cfg["loss"] == "cox"
                    else RankingSurvLoss()(risks=risk.reshape(-1), times=times, c=c))
            rec = {"risk": risk.detach().clone()}
        else:
            risk, hazards, S = res
            Y = torch.arange(cfg["B"]) % 4
            loss = NLLSurvLoss(alpha=0.15)(hazards=hazards, S=S, Y=Y, c=c)
            rec = {"risk": risk.detach().clone(), "hazards": hazards.detach().clone(), "S": S.detach().clone()}
        model.zero_grad()
        loss.backward()
        rec.update({
            "weights_fp": cases.fingerprint_state(model.state_dict()), "loss": loss.detach().reshape(()).clone(),
            "d_inputs": [None if t.grad is None else t.grad.clone() for t in (hr, hp, ho)],
            "grads": {k: cases.fingerprint(p.grad) for k, p in model.named_parameters()},
        })
        out["heads"][name] = rec
        print("head", name, float(loss))

    # ---- losses -----------------------------------------------------------------------------
    for name, cfg in cases.LOSS_CASES.items():
        rec = {}
        if cfg["loss"] == "nll":
            hz, S, Y, c = cases.nll_inputs(cfg)
            hz.requires_grad_(True); S.requires_grad_(True)
            loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hz, S=S, Y=Y, c=c)
            loss.backward()
            rec = {"loss": loss.detach().clone(), "d_hazards": hz.grad.clone(), "d_S": S.grad.clone()}
        else:
            r, times, c = cases.risk_inputs(cfg)
            r.requires_grad_(True)
            if cfg["loss"] == "cox":
                loss = CoxSurvLoss()(risks=r, times=times, c=c)
            else:
                loss = RankingSurvLoss(phi=cfg["phi"], reduction=cfg["reduction"])(risks=r, times=times, c=c)
            loss = loss.reshape(())
            if loss.requires_grad and loss.grad_fn is not None:
                loss.backward()
            rec = {"loss": loss.detach().clone(),
                   "d_risk": torch.zeros_like(r) if r.grad is None else r.grad.clone()}
        out["losses"][name] = rec
        print("loss", name, float(rec["loss"]))

    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_goldens.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
