"""Generates tests/golden/reference_goldens_mm.pt: the reference's own MM_MIL_Attention_fc_surv
(models/model_mm_attention_mil.py) — forward + nll_surv backward for five mode / fusion combinations and its four
captum* entry points — on the seeded MM_CASES / CAPTUM_CASES of oracle/cases.py.

The class cannot be constructed as shipped (SURVEY.md App. B-1,2): the subclass passes `gate_omic` to a base class that
does not take it (:124) and the base constructor reads an undefined name `size_path` (:83). Both are worked around at RUN
TIME, without editing or copying the reference:
  * `size_path` is resolved as a module global, so it is injected into the imported module (= size_dict_WSI['small']);
  * the instance is made with `__new__` and the BASE constructor (whose body is the whole construction) is called directly.
forward() and the captum* methods then run unmodified (radio_fusion='concat'). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens_mm.py
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


def build_reference_mm(mm_mod, cfg, d_in):
    cls = mm_mod.MM_MIL_Attention_fc_surv
    model = cls.__new__(cls)
    mm_mod.MM_MIL_Attention_fc.__init__(model, input_dim=d_in, radio_fusion="concat", fusion=cfg["fusion"], gate=True,
                                        gate_path=True, gate_radio=True, dropout=False, model_size_radio="small",
                                        model_size_wsi="small", model_size_omic="small", n_classes=4, mode=cfg["mode"])
    return model.eval()


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor          # CPU shim for model_modules.py:164
    import models.model_mm_attention_mil as mm_mod
    from utils.loss_utils import NLLSurvLoss
    mm_mod.size_path = [1024, 256, 256]                 # the name :83 looks up (size_dict_WSI['small'])

    out = {"torch": torch.__version__, "mm": {}, "captum": {}}
    for name, cfg in cases.MM_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = build_reference_mm(mm_mod, cfg, cfg["d"])
        cases.perturb_biases(model, cfg["seed"])
        kw = cases.mm_inputs(cfg)
        Y, c = cases.labels(cfg)
        hazards, S, Y_hat, A_raw = model(**kw)
        loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y, c=c)
        model.zero_grad()
        loss.backward()
        out["mm"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()), "hazards": hazards.detach().clone(),
            "S": S.detach().clone(), "Y_hat": Y_hat.clone(), "A_raw": {k: v.detach().clone() for k, v in A_raw.items()},
            "loss": loss.detach().clone(),
            "grads": {k: (None if p.grad is None else cases.fingerprint(p.grad)) for k, p in model.named_parameters()},
        }
        print("mm", name, float(loss.detach()), [tuple(v.shape) for v in A_raw.values()])
    for name, cfg in cases.CAPTUM_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = build_reference_mm(mm_mod, cfg, cfg["d"])
        cases.perturb_biases(model, cfg["seed"])
        args, w = cases.captum_inputs(cfg)
        args = [a.requires_grad_(True) for a in args]
        risk = getattr(model, cfg["fn"])(*args)
        model.zero_grad()
        (risk * w).sum().backward()
        out["captum"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()), "risk": risk.detach().clone(),
            "d_inputs": [cases.fingerprint(a.grad) for a in args],
            "grads": {k: (None if p.grad is None else cases.fingerprint(p.grad)) for k, p in model.named_parameters()},
        }
        print("captum", name, risk.detach().tolist())
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_goldens_mm.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
