"""Fusion heads over pre-extracted 256-d embeddings with a discrete-hazard output — drop-in for
the `kronecker` route of models/nll_models_pretrained.py (:101-103,179-197)."""
import torch
import torch.nn as nn

from ..autograd import HazardHead
from ..utils.utils import initialize_weights
from .coxranking_models_pretrained import _pick
from .model_modules import XlinearFusion


class multimodal_pretrained(nn.Module):
    def __init__(self, input_dim: int = 37, dropout=True, n_classes=4, mode='radio_path_omic', train_type=None,
                 bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.mode, self.train_type = n_classes, mode, train_type
        self.bag_loss, self.n_layers = bag_loss, n_layers
        num_modalities = sum(k in mode for k in ('radio', 'path', 'omic'))
        if train_type == 'kronecker':
            self.xfusion = XlinearFusion(num_modalities=num_modalities, dropout_rate=0.7)
            self.classifier = nn.Linear(256, n_classes)
        else:
            raise NotImplementedError(
                f"train_type={train_type!r}: only 'kronecker' is on the accelerated path this round")
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, h_radio, h_path, h_omic):
        MM = self.xfusion(v_list=_pick(self.mode, h_radio, h_path, h_omic))
        hazards, S, _ = HazardHead.apply(MM, self.classifier.weight, self.classifier.bias)
        risk = -torch.sum(S, dim=1)
        return risk, hazards, S
