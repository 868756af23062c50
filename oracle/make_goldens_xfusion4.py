"""Generates tests/golden/reference_goldens_xfusion4.pt: the UNMODIFIED reference XlinearFusion
(models/model_modules.py:113-178) with its default num_modalities=4 — the 17^4 = 83 521-wide Kronecker product that
`radio_fusion='tensor'` / the four-modality fusion of SURVEY.md §8f n4 needs — on the seeded XFUSION4_CASES of
oracle/cases.py: fused features, gradients of the four input embeddings and (as fingerprints) of every parameter for
the scalar objective sum(out * proj). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens_xfusion4.py
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
    torch.cuda.FloatTensor = torch.FloatTensor      # CPU shim for model_modules.py:164
    from models.model_modules import XlinearFusion

    out = {"torch": torch.__version__, "xfusion4": {}}
    for name, cfg in cases.XFUSION4_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = XlinearFusion().eval()                # reference defaults: 4 modalities, dim 256, scale 16, skip
        cases.perturb_biases(model, cfg["seed"])
        vs, proj = cases.embeddings4(cfg)
        vs = [v.requires_grad_(True) for v in vs]
        feats = model(v_list=vs)
        loss = (feats * proj).sum()
        model.zero_grad()
        loss.backward()
        out["xfusion4"][name] = {
            "weights_fp": cases.fingerprint_state(model.state_dict()), "features": feats.detach().clone(),
            "loss": loss.detach().clone(), "d_inputs": [v.grad.clone() for v in vs],
            "grads": {k: cases.fingerprint(p.grad) for k, p in model.named_parameters()},
        }
        print("xfusion4", name, float(loss), tuple(feats.shape))
    # ---- radiology AMIL with radio_fusion='tensor' (attribute-name repair only) ---------------------------
    from models.model_attention_mil_radio import MIL_Attention_fc_surv_radio
    from utils.loss_utils import NLLSurvLoss
    out["radio_tensor"] = {}
    for name, cfg in cases.RADIO_TENSOR_CASES.items():
        torch.manual_seed(cfg["seed"])
        model = MIL_Attention_fc_surv_radio(radio_fusion="tensor", gate_radio=cfg["gated"], dropout=cfg["dropout"],
                                            n_classes=cfg["K"]).eval()
        model.xfusion = model.radio_xfusion           # model_attention_mil_radio.py:84 calls the missing name
        cases.perturb_biases(model, cfg["seed"])
        bags = cases.radio_bags(cfg)
        Y, c = cases.labels(cfg)
        hazards, S, Y_hat, A_raw = model(**bags)
        M = model(**bags, return_features=True)
        loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hazards, S=S, Y=Y, c=c)
        model.zero_grad()
        loss.backward()
        with torch.no_grad():
            x1 = model.radio_xfusion(v_list=[bags[m][0].unsqueeze(0) for m in model.modalities])
            pre = model.attention_net_radio[0](x1)
        grads = {k: cases.fingerprint(p.grad) for k, p in model.named_parameters() if not k.startswith("xfusion.")}
        sd = {k: v for k, v in model.state_dict().items() if not k.startswith("xfusion.")}
        out["radio_tensor"][name] = {
            "weights_fp": cases.fingerprint_state(sd), "A_raw": A_raw.detach().clone(), "M": M.detach().clone(),
            "hazards": hazards.detach().clone(), "S": S.detach().clone(), "loss": loss.detach().clone(), "grads": grads,
            "relu_margin": (pre.abs().min() / pre.abs().max()).item(),
        }
        print("radio_tensor", name, float(loss), "fused row max", float(x1.abs().max()),
              "min|pre|/max|pre|", out["radio_tensor"][name]["relu_margin"])
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_goldens_xfusion4.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
