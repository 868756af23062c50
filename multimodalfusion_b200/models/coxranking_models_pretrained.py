"""Fusion heads over pre-extracted 256-d embeddings with a scalar risk output — drop-in for the
`kronecker` route of models/coxranking_models_pretrained.py (:62-183, kronecker :95-97,171-180).
The fcnn / highway / residual heads (BatchNorm-based) are "next" rows of SURVEY.md §8(f)."""
import torch
import torch.nn as nn

from .._lib import ACT_NONE
from ..autograd import Dense
from ..utils.utils import initialize_weights
from .model_modules import XlinearFusion


def _pick(mode, h_radio, h_path, h_omic):
    """Modality order used by the reference's kronecker / early branches (:147-180)."""
    r, p, o = 'radio' in mode, 'path' in mode, 'omic' in mode
    if r and p and o:
        return [h_radio, h_path, h_omic]
    if r and p:
        return [h_radio, h_path]
    if r and o:
        return [h_radio, h_omic]
    if o and p:
        return [h_omic, h_path]
    raise NotImplementedError(f"mode {mode!r} needs at least two modalities")


class multimodal_pretrained(nn.Module):
    def __init__(self, dropout=True, n_classes=4, mode='radio_path_omic', train_type=None,
                 bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.mode, self.train_type = n_classes, mode, train_type
        self.bag_loss, self.n_layers = bag_loss, n_layers
        num_modalities = sum(k in mode for k in ('radio', 'path', 'omic'))
        if train_type == 'kronecker':
            self.xfusion = XlinearFusion(num_modalities=num_modalities, dropout_rate=0.7)
            self.classifier = nn.Linear(256, 1)
        else:
            raise NotImplementedError(
                f"train_type={train_type!r}: only 'kronecker' is on the accelerated path this round")
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, h_radio, h_path, h_omic):
        MM = self.xfusion(v_list=_pick(self.mode, h_radio, h_path, h_omic))
        risk = Dense.apply(MM, self.classifier.weight, self.classifier.bias, ACT_NONE)
        return risk, None, None
