"""Weight initialisers consumed by the drop-in models (reference: utils/utils.py:217-233).
Same RNG consumption order as the reference so that a seeded construction yields an identical
state_dict."""
import math

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
