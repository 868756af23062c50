"""Heads over pre-extracted 256-d embeddings with a scalar risk output — drop-in for
models/coxranking_models_pretrained.py: `unimonal_pretrained` (:14-58: fcnn / highway / residual on ONE modality) and
`multimodal_pretrained` (:62-183): `kronecker` (:95-97,171-180), `early-fcnn` / `late-fcnn`
(:80-86,137-140,166), `early-highway` / `late-highway` (:87-94,141-144,167-168)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .._lib import ACT_NONE
from ..autograd import Dense
from ..utils.utils import initialize_weights
from .model_modules import Highway, Residual, XlinearFusion, fcnn_forward


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


def _pick_late(mode, outs):
    """Concatenation order of the late branches (:146-153): (radio, path), (radio, omic), (omic, path),
    (radio, path, omic)."""
    r, p, o = 'radio' in mode, 'path' in mode, 'omic' in mode
    if r and p and o:
        return [outs['radio'], outs['path'], outs['omic']]
    if r and p:
        return [outs['radio'], outs['path']]
    if r and o:
        return [outs['radio'], outs['omic']]
    if o and p:
        return [outs['omic'], outs['path']]
    raise NotImplementedError(f"mode {mode!r} needs at least two modalities")


def _unimodal_input(mode, kwargs):
    """h_path / h_radio / h_omic by mode, as the reference selects it (:42-47)."""
    if mode not in ('path', 'radio', 'omic'):
        raise NotImplementedError(f"mode={mode!r}")       # the reference hits an UnboundLocalError here
    return kwargs['h_' + mode].float()


class unimonal_pretrained(nn.Module):
    def __init__(self, dropout=True, n_classes=4, mode='radio', train_type=None, bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.train_type, self.bag_loss = n_classes, train_type, bag_loss
        self.mode, self.n_layers = mode, n_layers
        if train_type == 'fcnn':
            self.classifier = nn.Sequential(nn.Linear(256, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.7),
                                            nn.Linear(128, 1))
        elif train_type == 'highway':
            self.highway = Highway(256, n_layers, F.relu)
            self.classifier = nn.Linear(256, 1)
        elif train_type == 'residual':
            self.residual = Residual(256, n_layers)
            self.classifier = nn.Linear(256, 1)
        initialize_weights(self)      # (like the reference, other train_type values construct nothing and fail in forward)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, **kwargs):
        h = _unimodal_input(self.mode, kwargs)
        if self.train_type == 'fcnn':
            return fcnn_forward(self.classifier, h).squeeze(), None, None
        if self.train_type == 'highway':
            h = self.highway(h)
        elif self.train_type == 'residual':
            h = self.residual(h)
        else:
            raise NotImplementedError(f"train_type={self.train_type!r}")
        return Dense.apply(h, self.classifier.weight, self.classifier.bias, ACT_NONE).squeeze(), None, None


class multimodal_pretrained(nn.Module):
    def __init__(self, dropout=True, n_classes=4, mode='radio_path_omic', train_type=None,
                 bag_loss=None, n_layers=1):
        super().__init__()
        self.n_classes, self.mode, self.train_type = n_classes, mode, train_type
        self.bag_loss, self.n_layers = bag_loss, n_layers
        num_modalities = sum(k in mode for k in ('radio', 'path', 'omic'))
        fcnn = lambda d_in: nn.Sequential(nn.Linear(d_in, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.7),
                                          nn.Linear(128, 1))
        if train_type == 'late-fcnn':
            self.layer_WSI, self.layer_MRI, self.layer_omic = fcnn(256), fcnn(256), fcnn(256)
            self.classifier = nn.Sequential(nn.Linear(num_modalities, 1))
        elif train_type == 'early-fcnn':
            self.classifier = fcnn(num_modalities * 256)
        elif train_type == 'early-highway':
            self.highway = Highway(num_modalities * 256, n_layers, F.relu)
            self.classifier = nn.Linear(num_modalities * 256, 1)
        elif train_type == 'late-highway':
            self.highway_radio = Highway(256, n_layers, F.relu)
            self.highway_path = Highway(256, n_layers, F.relu)
            self.highway_omic = Highway(256, n_layers, F.relu)
            self.classifier = nn.Linear(num_modalities * 256, 1)
        elif train_type == 'kronecker':
            self.xfusion = XlinearFusion(num_modalities=num_modalities, dropout_rate=0.7)
            self.classifier = nn.Linear(256, 1)
        else:
            # the reference constructs nothing for other values and fails later in forward (AttributeError)
            raise NotImplementedError(f"train_type={train_type!r}")
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, h_radio, h_path, h_omic):
        tt = self.train_type
        if tt == 'kronecker':
            MM = self.xfusion(v_list=_pick(self.mode, h_radio, h_path, h_omic))
            return Dense.apply(MM, self.classifier.weight, self.classifier.bias, ACT_NONE), None, None
        if tt.startswith('late'):
            # the reference evaluates all three branches whatever the mode (:137-144); we run the ones the mode uses
            branch = {'radio': (self.layer_MRI, h_radio), 'path': (self.layer_WSI, h_path), 'omic': (self.layer_omic, h_omic)} \
                if tt == 'late-fcnn' else \
                {'radio': (self.highway_radio, h_radio), 'path': (self.highway_path, h_path), 'omic': (self.highway_omic, h_omic)}
            outs = {k: (fcnn_forward(m, h.float()) if tt == 'late-fcnn' else m(h)) for k, (m, h) in branch.items()
                    if k in self.mode}
            MM = torch.cat(_pick_late(self.mode, outs), dim=1)                      # == cat(axis=2) of the unsqueezed layers
            lin = self.classifier[0] if tt == 'late-fcnn' else self.classifier
            risk = Dense.apply(MM, lin.weight, lin.bias, ACT_NONE)
            return risk.unsqueeze(0).squeeze(), None, None                          # reference: [1,B,1].squeeze()
        MM = torch.cat([h.float() for h in _pick(self.mode, h_radio, h_path, h_omic)], dim=1)
        if tt == 'early-fcnn':
            return fcnn_forward(self.classifier, MM), None, None
        MM = self.highway(MM)
        return Dense.apply(MM, self.classifier.weight, self.classifier.bias, ACT_NONE), None, None

    # ---- captum entry points (models/coxranking_models_pretrained.py:202-305; called by create_attributions.py on
    # initiate_pretrained_model(...)): the risk as a function of the embeddings only, fixed modality orders -------------
    def _captum_risk(self, pairs):
        """pairs: [(modality key, embedding)] in the reference method's concatenation order."""
        tt = self.train_type
        hs = [h for _, h in pairs]
        if tt == 'kronecker':
            MM = self.xfusion(v_list=hs)
            return Dense.apply(MM, self.classifier.weight, self.classifier.bias, ACT_NONE)
        if tt.startswith('late'):
            mods = {'radio': self.layer_MRI, 'path': self.layer_WSI, 'omic': self.layer_omic} if tt == 'late-fcnn' else \
                {'radio': self.highway_radio, 'path': self.highway_path, 'omic': self.highway_omic}
            outs = [fcnn_forward(mods[k], h.float()) if tt == 'late-fcnn' else mods[k](h) for k, h in pairs]
            MM = torch.cat(outs, dim=1)                     # == cat(axis=2) of the unsqueeze(0)-ed layers
            lin = self.classifier[0] if tt == 'late-fcnn' else self.classifier
            return Dense.apply(MM, lin.weight, lin.bias, ACT_NONE).unsqueeze(0).squeeze()
        MM = torch.cat([h.float() for h in hs], dim=1)
        if tt == 'early-fcnn':
            return fcnn_forward(self.classifier, MM)
        MM = self.highway(MM)
        return Dense.apply(MM, self.classifier.weight, self.classifier.bias, ACT_NONE)

    def captum_radio_path(self, h_radio, h_path):
        return self._captum_risk([('radio', h_radio), ('path', h_path)])

    def captum_path_omic(self, h_omic, h_path):
        return self._captum_risk([('omic', h_omic), ('path', h_path)])

    def captum_radio_omic(self, h_radio, h_omic):
        return self._captum_risk([('radio', h_radio), ('omic', h_omic)])

    def captum(self, h_radio, h_path, h_omic):
        return self._captum_risk([('radio', h_radio), ('path', h_path), ('omic', h_omic)])
