"""Generates tests/golden/reference_goldens_unimodal.pt: `unimonal_pretrained` of the UNMODIFIED reference
(models/coxranking_models_pretrained.py:14-58 — fcnn / highway / residual; models/nll_models_pretrained.py:14-62 — fcnn /
highway) and the Residual blocks (models/model_modules.py:28-59) on the seeded UNI_CASES of oracle/cases.py.
Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens_unimodal.py
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
    torch.cuda.FloatTensor = torch.FloatTensor
    from models import coxranking_models_pretrained as cox_heads
    from models import nll_models_pretrained as nll_heads
    from utils.loss_utils import CoxSurvLoss, CrossEntropySurvLoss, NLLSurvLoss, RankingSurvLoss

    out = {"torch": torch.__version__, "unimodal": {}}
    for name, cfg in cases.UNI_CASES.items():
        torch.manual_seed(cfg["seed"])
        mod = cox_heads if cfg["kind"] == "cox" else nll_heads
        model = mod.unimonal_pretrained(mode=cfg["mode"], train_type=cfg["train_type"], n_classes=4,
                                        n_layers=cfg["n_layers"]).eval()
        cases.perturb_biases(model, cfg["seed"])
        cases.perturb_bn(model, cfg["seed"])
        hr, hp, ho = cases.embeddings(cfg)
        h = {"radio": hr, "path": hp, "omic": ho}[cfg["mode"]].requires_grad_(True)
        times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
        res = model(**{"h_" + cfg["mode"]: h})
        if cfg["kind"] == "cox":
            risk = res[0]
            loss = (CoxSurvLoss()(risks=risk, times=times, c=c) if cfg["loss"] == "cox"
                    else RankingSurvLoss()(risks=risk.reshape(-1), times=times, c=c))
            rec = {"risk": risk.detach().clone()}
        else:
            risk, hazards, S = res
            Y = torch.arange(cfg["B"]) % 4
            lf = NLLSurvLoss(alpha=0.15) if cfg["loss"] == "nll" else CrossEntropySurvLoss(alpha=0.15)
            loss = lf(hazards=hazards, S=S, Y=Y, c=c)
            rec = {"risk": risk.detach().clone(), "hazards": hazards.detach().clone(), "S": S.detach().clone()}
        model.zero_grad()
        loss.backward()
        rec.update({
            "weights_fp": cases.fingerprint_state(model.state_dict()), "loss": loss.detach().reshape(()).clone(),
            "d_input": h.grad.clone(),
            "grads": {k: (None if p.grad is None else cases.fingerprint(p.grad)) for k, p in model.named_parameters()},
        })
        out["unimodal"][name] = rec
        print("unimodal", name, float(loss.detach()), tuple(risk.shape))
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_goldens_unimodal.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
