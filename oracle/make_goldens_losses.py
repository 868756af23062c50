"""Generates tests/golden/reference_goldens_losses.pt: RankingNLLSurvLoss of the UNMODIFIED reference
(utils/loss_utils.py:151-164 = ranking_loss over the label BINS Y as times + nll_ratio * nll_loss) with the gradients
w.r.t. the logits and the risks, on seeded cohorts (B = 2 .. 64, ties, all-censored, a zero-pair case). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens_losses.py
"""
from __future__ import annotations

import os
import sys

import torch

REF = os.environ.get("MMF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

CASES = {
    # name: (B, K, seed, phi, reduction, alpha, nll_ratio, censor mode)
    "rnll_b8_sigmoid_mean": (8, 4, 11, "sigmoid", "mean", 0.15, 0.5, "mixed"),
    "rnll_b2_relu_sum": (2, 4, 12, "relu", "sum", 0.0, 1.0, "mixed"),
    "rnll_b33_sigmoid_sum": (33, 4, 13, "sigmoid", "sum", 0.4, 0.25, "mixed"),
    "rnll_b64_relu_mean": (64, 8, 14, "relu", "mean", 0.15, 0.5, "mixed"),
    "rnll_b16_all_censored": (16, 4, 15, "sigmoid", "mean", 0.15, 0.5, "all"),     # no comparable pair: ranking term 0
    "rnll_b16_none_censored": (16, 4, 16, "sigmoid", "mean", 0.15, 0.5, "none"),
}


def case_inputs(name):
    B, K, seed, phi, red, alpha, ratio, mode = CASES[name]
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, K, generator=g)
    risks = torch.randn(B, generator=g)
    Y = torch.randint(0, K, (B,), generator=g)
    c = {"mixed": (torch.rand(B, generator=g) < 0.4).float(), "all": torch.ones(B), "none": torch.zeros(B)}[mode]
    return logits, risks, Y, c, dict(phi=phi, reduction=red, alpha=alpha, nll_ratio=ratio)


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor
    from utils.loss_utils import RankingNLLSurvLoss
    out = {"torch": torch.__version__, "losses": {}}
    for name in CASES:
        logits, risks, Y, c, kw = case_inputs(name)
        logits.requires_grad_(True); risks.requires_grad_(True)
        hazards = torch.sigmoid(logits)                       # models/model_attention_mil_path.py:59-60
        S = torch.cumprod(1 - hazards, dim=1)
        loss = RankingNLLSurvLoss(**kw)(hazards=hazards, risks=risks, S=S, Y=Y, c=c)
        loss = loss.reshape(()) if torch.is_tensor(loss) else torch.tensor(float(loss))
        if loss.requires_grad:
            loss.backward()
        out["losses"][name] = {"loss": loss.detach().clone(),
                               "dlogits": torch.zeros_like(logits) if logits.grad is None else logits.grad.clone(),
                               "drisks": torch.zeros_like(risks) if risks.grad is None else risks.grad.clone()}
        print(name, float(loss))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_goldens_losses.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
