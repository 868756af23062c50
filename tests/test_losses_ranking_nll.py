"""RankingNLLSurvLoss (utils/loss_utils.py:151-164): ranking_loss over the label bins Y (int64) as times +
nll_ratio * nll_loss. Goldens from the unmodified reference (oracle/make_goldens_losses.py): loss, d/dlogits, d/drisks.
CPU: the oracle restatement; GPU: the drop-in mirror (ranking pair-grid kernel + nll_surv kernel through the C-ABI)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import amil_oracle as O  # noqa: E402
from oracle.make_goldens_losses import CASES, case_inputs  # noqa: E402


def _check(loss, logits, risks, gold):
    assert abs(loss.item() - gold["loss"].item()) < 1e-5 * max(1.0, abs(gold["loss"].item()))
    if loss.requires_grad:
        loss.backward()
    dl = torch.zeros_like(logits) if logits.grad is None else logits.grad.cpu()
    dr = torch.zeros_like(risks) if risks.grad is None else risks.grad.cpu()
    assert (dl - gold["dlogits"]).abs().max().item() < 1e-6 + 1e-5 * gold["dlogits"].abs().max().item()
    assert (dr - gold["drisks"]).abs().max().item() < 1e-6 + 1e-5 * gold["drisks"].abs().max().item()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_ranking_nll_vs_reference_goldens(name, goldens_losses):
    logits, risks, Y, c, kw = case_inputs(name)
    logits.requires_grad_(True); risks.requires_grad_(True)
    hz = torch.sigmoid(logits)
    S = torch.cumprod(1 - hz, dim=1)
    loss = (O.ranking_loss(risks, Y, c, kw["phi"], kw["reduction"]).reshape(())
            + kw["nll_ratio"] * O.nll_surv_loss(hz, S, Y, c, alpha=kw["alpha"]).reshape(()))
    _check(loss, logits, risks, goldens_losses["losses"][name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_ranking_nll_vs_reference_goldens(name, goldens_losses):
    from multimodalfusion_b200.utils import RankingNLLSurvLoss
    dev = torch.device("cuda")
    logits, risks, Y, c, kw = case_inputs(name)
    logits = logits.to(dev).requires_grad_(True); risks = risks.to(dev).requires_grad_(True)
    hz = torch.sigmoid(logits)
    S = torch.cumprod(1 - hz, dim=1)
    loss = RankingNLLSurvLoss(**kw)(hazards=hz, risks=risks, S=S, Y=Y.to(dev), c=c.to(dev)).reshape(())
    _check(loss.cpu() if not loss.requires_grad else loss, logits, risks, goldens_losses["losses"][name])
