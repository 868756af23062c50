"""Weight initialisers consumed by the drop-in models (reference: utils/utils.py:217-233; same RNG consumption order
as the reference so that a seeded construction yields an identical state_dict) and the small model utilities the
training loops call next to the path (freeze / unfreeze, the L1 penalty: utils/utils.py:235-269)."""
import math

import torch
import torch.nn as nn


def initialize_weights(module):
    """Xavier-normal Linear weights, zero biases, unit BatchNorm (utils/utils.py:217-226)."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def init_max_weights(module):
    """N(0, 1/sqrt(fan_in)) Linear weights, zero biases (utils/utils.py:228-233)."""
    for m in module.modules():
        if type(m) == nn.Linear:
            m.weight.data.normal_(0, 1.0 / math.sqrt(m.weight.size(1)))
            m.bias.data.zero_()


def dfs_freeze(model):
    """utils/utils.py:235-239: every parameter below `model` stops requiring a gradient."""
    for _, child in model.named_children():
        for param in child.parameters():
            param.requires_grad = False
        dfs_freeze(child)


def dfs_unfreeze(model):
    """utils/utils.py:242-246."""
    for _, child in model.named_children():
        for param in child.parameters():
            param.requires_grad = True
        dfs_unfreeze(child)


def l1_reg_all(model, reg_type=None):
    """sum |W| over all parameters as a differentiable scalar (utils/utils.py:249-257) — the reference adds
    lambda_reg * this to every step's loss (utils/core_utils.py:218-221). Same call, same value, one fused
    multi-tensor norm instead of one abs + sum launch pair per parameter. The step-fused form (penalty gradient
    lambda * sign(W) applied inside the Adam kernel, nothing added to the autograd graph) is
    `FusedAdam(..., l1_lambda=lambda_reg)`."""
    params = list(model.parameters())
    if not params:
        return None
    norms = torch._foreach_norm(params, 1)
    return torch.stack([n.reshape(()) for n in norms]).sum()


def l1_reg_modules(model, reg_type=None):
    """utils/utils.py:259-269: L1 of the genomic SNN and, when the model has one, of the fusion block."""
    l1_reg = l1_reg_all(model.fc_omic)
    mm = getattr(model, "mm", None)
    if mm is not None:
        l1_reg = l1_reg + l1_reg_all(mm)
    return l1_reg
