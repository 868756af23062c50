"""Genomic self-normalising network ("MaxNet") — drop-in for models/model_genomic.py
(constructor :13-44; forward :53-72; captum wrapper :77-91)."""
import torch
import torch.nn as nn

from .._lib import ACT_NONE
from ..autograd import Dense, HazardHead
from ..utils.utils import init_max_weights
from .model_modules import SNN_Block, snn_forward


class MaxNet_base(nn.Module):
    def __init__(self, input_dim: int, model_size_omic: str = 'small', bag_loss=None, n_classes: int = 4):
        super().__init__()
        self.n_classes = n_classes
        self.size_dict_omic = {'small': [256, 256], 'big': [1024, 256]}
        self.bag_loss = bag_loss
        hidden = self.size_dict_omic[model_size_omic]
        blocks = [SNN_Block(dim1=input_dim, dim2=hidden[0])]
        for i in range(len(hidden) - 1):
            blocks.append(SNN_Block(dim1=hidden[i], dim2=hidden[i + 1], dropout=0.25))
        self.fc_omic = nn.Sequential(*blocks)
        # `'nll' in None` raises TypeError in the reference as well (model_genomic.py:33)
        self.classifier = nn.Linear(hidden[-1], n_classes if 'nll' in self.bag_loss else 1)
        init_max_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.fc_omic = self.fc_omic.to(device)
        self.classifier = self.classifier.to(device)

    def features(self, x):
        return snn_forward(self.fc_omic, x)

    def forward(self, **kwargs):
        pass


class MaxNet(MaxNet_base):
    def forward(self, **kwargs):
        x = kwargs['genomic_features']
        # the reference's loaders hand over ONE patient's omics as a 1-D [d] tensor (the collate concatenates 1-D rows,
        # datasets/dataset_survival.py); nn.Linear takes it as is and `logits.unsqueeze(0)` makes the [1,K] batch
        one_d = x.dim() == 1
        feats = self.features(x.unsqueeze(0) if one_d else x)
        if kwargs.get('return_features'):
            return feats.squeeze(0) if one_d else feats
        if 'nll' in self.bag_loss:
            hazards, S, Y_hat = HazardHead.apply(feats, self.classifier.weight, self.classifier.bias)
            # reference: logits.unsqueeze(0) -> [1,B,K]; cumprod/topk along dim=1 (the batch axis
            # for B>1). Only B=1 (the training loop's batch) is meaningful there and is what we keep:
            return hazards, S, Y_hat, None
        risk = Dense.apply(feats, self.classifier.weight, self.classifier.bias, ACT_NONE).squeeze()
        return risk, None, None, None


class MaxNet_captum(MaxNet_base):
    def forward(self, x):
        feats = self.features(x.unsqueeze(0) if x.dim() == 1 else x)
        if 'nll' in self.bag_loss:
            _, S, _ = HazardHead.apply(feats, self.classifier.weight, self.classifier.bias)
            return -torch.sum(S, dim=1)
        return Dense.apply(feats, self.classifier.weight, self.classifier.bias, ACT_NONE)
