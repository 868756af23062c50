"""Survival losses with the reference's call signatures (reference: utils/loss_utils.py), each
evaluated (loss and gradient) by one libmmf_b200 kernel instead of ~12 ATen launches (NLL) or a
host-side O(B^2) Python loop (Cox, ranking)."""
import torch

from ..autograd import CeSurv, CoxLoss, NllSurv, RankingLoss


def nll_loss(hazards, S, Y, c, alpha=0.4, eps=1e-7):
    """utils/loss_utils.py:22-39."""
    if S is None:
        S = torch.cumprod(1 - hazards, dim=1)
    return NllSurv.apply(hazards, S, Y, c, float(alpha), float(eps))


def ce_loss(hazards, S, Y, c, alpha=0.4, eps=1e-7):
    """utils/loss_utils.py:41-56."""
    if S is None:
        S = torch.cumprod(1 - hazards, dim=1)
    return CeSurv.apply(hazards, S, Y, c, float(alpha), float(eps))


class CrossEntropySurvLoss(object):
    """utils/loss_utils.py:104-112."""

    def __init__(self, alpha=0.15):
        self.alpha = alpha

    def __call__(self, hazards, S, Y, c, alpha=None):
        return ce_loss(hazards, S, Y, c, alpha=self.alpha if alpha is None else alpha)


def ranking_loss(risks, times, c, phi, reduction):
    """utils/loss_utils.py:58-101. Zero comparable pairs -> 0 loss with zero gradient."""
    if len(times) == 1:
        raise NotImplementedError("Batch size must be at least 2")
    if phi not in ("sigmoid", "relu") or reduction not in ("mean", "sum"):
        raise NotImplementedError(f"phi={phi!r} reduction={reduction!r}")
    return RankingLoss.apply(risks.reshape(-1), times, c, phi, reduction)


class NLLSurvLoss(object):
    """utils/loss_utils.py:114-122."""

    def __init__(self, alpha=0.15):
        self.alpha = alpha

    def __call__(self, hazards, S, Y, c, alpha=None):
        return nll_loss(hazards, S, Y, c, alpha=self.alpha if alpha is None else alpha)


class CoxSurvLoss(object):
    """utils/loss_utils.py:124-139."""

    def __call__(self, risks, times, c, **kwargs):
        return CoxLoss.apply(risks.reshape(-1), times, c)


class RankingSurvLoss(object):
    """utils/loss_utils.py:142-149."""

    def __init__(self, phi='sigmoid', reduction='mean'):
        self.phi = phi
        self.reduction = reduction

    def __call__(self, risks, times, c):
        return ranking_loss(risks, times, c, self.phi, self.reduction)


class RankingNLLSurvLoss(object):
    """utils/loss_utils.py:151-164 (the ranking term uses the label bins Y as times, as the
    reference does)."""

    def __init__(self, phi='sigmoid', reduction='mean', alpha=0.15, nll_ratio=0.5):
        self.alpha, self.phi, self.reduction, self.nll_ratio = alpha, phi, reduction, nll_ratio

    def __call__(self, hazards, risks, S, Y, c, alpha=None):
        rank = ranking_loss(risks, Y, c, self.phi, self.reduction)
        nll = nll_loss(hazards, S, Y, c, alpha=self.alpha if alpha is None else alpha)
        return rank + nll * self.nll_ratio
