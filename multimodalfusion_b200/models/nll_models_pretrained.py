"""Heads over pre-extracted 256-d embeddings with a discrete-hazard output — drop-in for
models/nll_models_pretrained.py: `unimonal_pretrained` (:14-62: fcnn / highway on ONE modality; its residual branch is
commented out in the reference) and `multimodal_pretrained` (:64-197): `kronecker` (:101-103,179-188), `early-fcnn` / `late-fcnn` (:82-91),
`early-highway` / `late-highway` (:92-99)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .._lib import ACT_NONE
from ..autograd import Dense, HazardHead
from ..utils.utils import initialize_weights
from .coxranking_models_pretrained import _pick, _pick_late, _unimodal_input
from .model_modules import Highway, XlinearFusion, fcnn_forward


class unimonal_pretrained(nn.Module):
    def __init__(self, dropout=True, n_classes=4, mode=None, train_type=None, bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.train_type, self.bag_loss, self.mode = n_classes, train_type, bag_loss, mode
        if train_type == 'fcnn':
            self.classifier = nn.Sequential(nn.Linear(256, n_classes), nn.Dropout(0.7))
        elif train_type == 'highway':
            self.highway = Highway(256, n_layers, F.relu)
            self.classifier = nn.Linear(256, n_classes)
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, **kwargs):
        h = _unimodal_input(self.mode, kwargs)
        if self.train_type == 'fcnn':
            lin = self.classifier[0]
            if self.training:
                # the reference drops LOGITS (Linear -> Dropout(0.7), :23): the head kernel fuses linear + sigmoid, so the
                # train-mode mask goes through the un-fused pair
                logits = self.classifier[1](Dense.apply(h, lin.weight, lin.bias, ACT_NONE))
                hazards = torch.sigmoid(logits)
                S = torch.cumprod(1 - hazards, dim=1)
                return -torch.sum(S, dim=1), hazards, S
        elif self.train_type == 'highway':
            h, lin = self.highway(h), self.classifier
        else:
            raise NotImplementedError(f"train_type={self.train_type!r}")
        hazards, S, _ = HazardHead.apply(h, lin.weight, lin.bias)
        return -torch.sum(S, dim=1), hazards, S


class multimodal_pretrained(nn.Module):
    def __init__(self, input_dim: int = 37, dropout=True, n_classes=4, mode='radio_path_omic', train_type=None,
                 bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.mode, self.train_type = n_classes, mode, train_type
        self.bag_loss, self.n_layers = bag_loss, n_layers
        num_modalities = sum(k in mode for k in ('radio', 'path', 'omic'))
        fcnn = lambda d_in, d_out=None: nn.Sequential(*([nn.Linear(d_in, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.7)]
                                                        + ([nn.Linear(128, d_out)] if d_out else [])))
        if train_type == 'early-fcnn':
            self.classifier = fcnn(num_modalities * 256, n_classes)
        elif train_type == 'late-fcnn':
            self.layer_WSI, self.layer_MRI, self.layer_omic = fcnn(256), fcnn(256), fcnn(256)
            self.classifier = nn.Sequential(nn.Linear(num_modalities * 128, n_classes))
        elif train_type == 'early-highway':
            self.highway = Highway(num_modalities * 256, n_layers, F.relu)
            self.classifier = nn.Linear(num_modalities * 256, n_classes)
        elif train_type == 'late-highway':
            self.highway_radio = Highway(256, n_layers, F.relu)
            self.highway_path = Highway(256, n_layers, F.relu)
            self.highway_omic = Highway(256, n_layers, F.relu)
            self.classifier = nn.Linear(num_modalities * 256, n_classes)
        elif train_type == 'kronecker':
            self.xfusion = XlinearFusion(num_modalities=num_modalities, dropout_rate=0.7)
            self.classifier = nn.Linear(256, n_classes)
        else:
            raise NotImplementedError(f"train_type={train_type!r}")
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, h_radio, h_path, h_omic):
        tt = self.train_type
        if tt == 'kronecker':
            MM, lin = self.xfusion(v_list=_pick(self.mode, h_radio, h_path, h_omic)), self.classifier
        elif tt.startswith('late'):
            branch = {'radio': (self.layer_MRI, h_radio), 'path': (self.layer_WSI, h_path), 'omic': (self.layer_omic, h_omic)} \
                if tt == 'late-fcnn' else \
                {'radio': (self.highway_radio, h_radio), 'path': (self.highway_path, h_path), 'omic': (self.highway_omic, h_omic)}
            outs = {k: (fcnn_forward(m, h.float()) if tt == 'late-fcnn' else m(h)) for k, (m, h) in branch.items()
                    if k in self.mode}
            MM = torch.cat(_pick_late(self.mode, outs), dim=1)
            lin = self.classifier[0] if tt == 'late-fcnn' else self.classifier
        else:
            MM = torch.cat([h.float() for h in _pick(self.mode, h_radio, h_path, h_omic)], dim=1)
            if tt == 'early-fcnn':
                MM = fcnn_forward(nn.Sequential(*list(self.classifier)[:4]), MM)    # up to the Dropout; last Linear = the head
                lin = self.classifier[4]
            else:
                MM, lin = self.highway(MM), self.classifier
        hazards, S, _ = HazardHead.apply(MM, lin.weight, lin.bias)
        risk = -torch.sum(S, dim=1)
        return risk, hazards, S

    # ---- captum entry points (models/nll_models_pretrained.py:200-318; called by create_attributions.py on
    # initiate_pretrained_model(...)): risk = -sum_k S_k as a function of the embeddings only, fixed modality orders ----
    def _captum_risk(self, pairs):
        """pairs: [(modality key, embedding)] in the reference method's concatenation order."""
        tt = self.train_type
        hs = [h for _, h in pairs]
        if tt == 'kronecker':
            MM, lin = self.xfusion(v_list=hs), self.classifier
        elif tt.startswith('late'):
            mods = {'radio': self.layer_MRI, 'path': self.layer_WSI, 'omic': self.layer_omic} if tt == 'late-fcnn' else \
                {'radio': self.highway_radio, 'path': self.highway_path, 'omic': self.highway_omic}
            outs = [fcnn_forward(mods[k], h.float()) if tt == 'late-fcnn' else mods[k](h) for k, h in pairs]
            MM = torch.cat(outs, dim=1)
            lin = self.classifier[0] if tt == 'late-fcnn' else self.classifier
        else:
            MM = torch.cat([h.float() for h in hs], dim=1)
            if tt == 'early-fcnn':
                MM = fcnn_forward(nn.Sequential(*list(self.classifier)[:4]), MM)
                lin = self.classifier[4]
            else:
                MM, lin = self.highway(MM), self.classifier
        _, S, _ = HazardHead.apply(MM, lin.weight, lin.bias)
        return -torch.sum(S, dim=1)

    def captum_radio_path(self, h_radio, h_path):
        return self._captum_risk([('radio', h_radio), ('path', h_path)])

    def captum_path_omic(self, h_omic, h_path):
        return self._captum_risk([('omic', h_omic), ('path', h_path)])

    def captum_radio_omic(self, h_radio, h_omic):
        return self._captum_risk([('radio', h_radio), ('omic', h_omic)])

    def captum(self, h_radio, h_path, h_omic):
        return self._captum_risk([('radio', h_radio), ('path', h_path), ('omic', h_omic)])
