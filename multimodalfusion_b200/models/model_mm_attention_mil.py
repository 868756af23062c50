"""End-to-end three-modality model (radiology AMIL + pathology AMIL + genomic SNN -> Kronecker or
concat fusion -> discrete-hazard head) — drop-in for models/model_mm_attention_mil.py
(constructor :19-98,:118-126; forward :128-200).

The reference class cannot be instantiated or run as shipped (SURVEY.md App. B-1,2,3,6). This
mirror keeps the constructor signature, attribute names and state_dict keys the reference
*defines*, and implements the evidently intended semantics:
  * ``gate_omic`` is accepted and ignored (the base class has no such parameter, :19-23 vs :124);
  * ``size_path`` (:83) is read as ``size_WSI``;
  * ``self.xfusion`` (:141) is read as ``self.radio_xfusion``; the branch is otherwise run as written
    (slice 0 of each modality -> 4-way Kronecker fusion -> a one-row radiology bag);
  * ``return_features`` returns the fused embedding ``MM`` (the reference references undefined
    names there, :196-198);
  * the ``captum*`` entry points (:202-396) keep the reference's arithmetic as written, including the
    soft-max over a singleton dimension that turns their pooling into a plain sum.
Parity is pinned against the reference itself made runnable WITHOUT editing it (oracle/make_goldens_mm.py:
the missing module global ``size_path`` is injected and the base-class constructor is called directly).
"""
import torch
import torch.nn as nn

from .._lib import ACT_NONE, ACT_RELU
from ..autograd import Dense, HazardHead, reduce_dim_forward
from ..utils.utils import initialize_weights
from .model_modules import AmilBranch, Attn_Net, Attn_Net_Gated, SNN_Block, XlinearFusion, snn_forward


class MM_MIL_Attention_fc(nn.Module):
    def __init__(self, input_dim: int = 80, radio_fusion='concat', fusion='tensor',
                 gate=True, gate_path=True, gate_radio=True, dropout=True,
                 model_size_radio: str = 'small', model_size_wsi: str = 'small',
                 model_size_omic: str = 'small', n_classes=4,
                 modalities=['T1', 'T2', 'T1Gd', 'FLAIR'], mode='radio_path_omic'):
        super().__init__()
        self.radio_fusion, self.fusion, self.n_classes = radio_fusion, fusion, n_classes
        self.size_dict_radio = {"small": [1024, 256, 256], "big": [1024, 256, 384]}
        self.size_dict_WSI = {"small": [1024, 256, 256], "big": [1024, 256, 384]}
        self.size_dict_omic = {'small': [256, 256], 'big': [1024, 256]}
        self.modalities, self.mode = modalities, mode

        size_omic = self.size_dict_omic[model_size_omic]
        blocks = [SNN_Block(dim1=input_dim, dim2=size_omic[0])]
        for i in range(len(size_omic) - 1):
            blocks.append(SNN_Block(dim1=size_omic[i], dim2=size_omic[i + 1], dropout=0.25))
        self.fc_omic = nn.Sequential(*blocks)

        size_radio = self.size_dict_radio[model_size_radio]
        attn_r = Attn_Net_Gated if gate_radio else Attn_Net
        self.attention_net_radio = nn.Sequential(
            nn.Linear(size_radio[0], size_radio[1]), nn.ReLU(), nn.Dropout(0.25),
            attn_r(L=size_radio[1], D=size_radio[2], dropout=dropout, n_classes=1))
        if radio_fusion == 'tensor':
            self.radio_xfusion = XlinearFusion(dim=1024, scale_dim=64, mmhid1=1024, mmhid2=1024, skip=0)
        elif radio_fusion == 'concat':
            self.reduce_dim = nn.Linear(size_radio[0] * len(modalities), size_radio[0])

        size_WSI = self.size_dict_WSI[model_size_wsi]
        attn_p = Attn_Net_Gated if gate_path else Attn_Net
        self.attention_net_WSI = nn.Sequential(
            nn.Linear(size_WSI[0], size_WSI[1]), nn.ReLU(), nn.Dropout(0.25),
            attn_p(L=size_WSI[1], D=size_WSI[2], dropout=dropout, n_classes=1))

        widths = {'radio': size_radio[1], 'path': size_WSI[1], 'omic': size_omic[1]}
        present = [k for k in ('radio', 'path', 'omic') if k in mode]
        classifier_size = sum(widths[k] for k in present)
        if fusion == 'tensor':
            self.mm = XlinearFusion(dim=256, scale_dim=16, mmhid1=512, mmhid2=512,
                                    num_modalities=len(present), gate=gate, skip=1)
            self.classifier = nn.Sequential(nn.Linear(512, 256), nn.ReLU(), nn.Dropout(0.25),
                                            nn.Linear(256, n_classes))
        elif fusion == 'concat':
            self.classifier = nn.Linear(classifier_size, n_classes)
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    def forward(self, h, return_features=False, attention_only=False):
        pass


class MM_MIL_Attention_fc_surv(MM_MIL_Attention_fc):
    def __init__(self, input_dim: int = 80, radio_fusion: str = 'concat', fusion: str = 'tensor',
                 gate=True, gate_path=True, gate_omic=True, gate_radio=True,
                 model_size_radio="small", model_size_wsi: str = 'small', model_size_omic='small',
                 dropout=False, n_classes=4, mode='radio_path_omic'):
        super().__init__(input_dim=input_dim, radio_fusion=radio_fusion, fusion=fusion, gate=gate,
                         gate_path=gate_path, gate_radio=gate_radio, model_size_radio='small',
                         model_size_wsi=model_size_wsi, model_size_omic=model_size_omic,
                         dropout=dropout, n_classes=n_classes, mode=mode)

    def forward(self, **kwargs):
        A_raw = {}
        emb = {}
        if 'radio' in self.mode:
            bags = [kwargs[m] for m in self.modalities]
            if len(bags) > 1:
                if self.radio_fusion == 'concat':
                    x = reduce_dim_forward(self.reduce_dim.weight, self.reduce_dim.bias, bags)
                elif self.radio_fusion == 'tensor':     # as written at :141 with the attribute name repaired
                    x = self.radio_xfusion(v_list=[b[0].unsqueeze(0) for b in bags])
                else:
                    raise NotImplementedError(f"radio_fusion={self.radio_fusion!r}")
            else:
                x = bags[0]
            A_raw['radiology'], emb['radio'] = AmilBranch.pooled(self.attention_net_radio, x, self.training)
        if 'path' in self.mode:
            A_raw['pathology'], emb['path'] = AmilBranch.pooled(self.attention_net_WSI,
                                                                kwargs['path_features'], self.training)
        if 'omic' in self.mode:
            o = kwargs['genomic_features']
            o = o.unsqueeze(0) if o.dim() == 1 else o
            emb['omic'] = snn_forward(self.fc_omic, o)
        # fusion order of the reference (:167-186): (radio, path), (radio, omic), (omic, path),
        # (radio, path, omic)
        has = lambda k: k in emb
        if has('radio') and has('path') and has('omic'):
            order = ['radio', 'path', 'omic']
        elif has('radio') and has('path'):
            order = ['radio', 'path']
        elif has('radio') and has('omic'):
            order = ['radio', 'omic']
        elif has('omic') and has('path'):
            order = ['omic', 'path']
        else:
            raise NotImplementedError(f"mode {self.mode!r} needs at least two modalities")
        vs = [emb[k] for k in order]
        MM, hid, Wk, bk = self._fuse(vs)
        if kwargs.get('return_features'):
            return MM
        hazards, S, Y_hat = HazardHead.apply(hid, Wk, bk)
        return hazards, S, Y_hat, A_raw

    def _fuse(self, vs):
        """Fusion + the classifier's hidden layer: returns (MM, input of the hazard layer, its weight, its bias)."""
        if self.fusion == 'tensor':
            MM = self.mm(v_list=vs)
            hid = Dense.apply(MM, self.classifier[0].weight, self.classifier[0].bias, ACT_RELU)
            hid = self.classifier[2](hid)
            return MM, hid, self.classifier[3].weight, self.classifier[3].bias
        MM = torch.cat(vs, dim=1)
        return MM, MM, self.classifier.weight, self.classifier.bias

    # ---- Captum entry points (models/model_mm_attention_mil.py:202-396) ------------------------------------------
    # Positional, batched 3-D bag inputs [B, N, 1024] (integrated gradients stacks n_steps interpolated copies of one
    # patient) and 2-D omics [B, d]; they return risk [B] = -sum_k S_k. Reference semantics kept AS WRITTEN: the
    # attention scores are transposed to [B, 1, N] and soft-maxed over dim=1 — a singleton — so every weight is 1 and
    # the "attention pooling" is a plain SUM over the instances (:222-226, :265-269, :280-285, :364-369); the attention
    # net therefore does not influence the returned risk. Everything runs in fp32 on the functor SGEMM kernels
    # (Dense / KronEncoder / HazardHead), which provide dX: attributions differentiate w.r.t. the inputs.
    def _captum_sum_pool(self, seq, x):
        if x.dim() != 3:
            raise ValueError("captum entry points take batched bags [B, N, 1024]")
        B, N, width = x.shape
        h = Dense.apply(x.reshape(B * N, width).float(), seq[0].weight, seq[0].bias, ACT_RELU)
        h = seq[2](h)                                   # nn.Dropout(0.25): identity in eval
        return h.view(B, N, -1).sum(dim=1)

    def _captum_radio(self, bags):
        if 'radio' not in self.mode:
            raise NotImplementedError('use another captum function')
        if self.radio_fusion != 'concat':
            # the reference's tensor branch (:218-220) feeds 3-D slices into XlinearFusion's 2-D torch.cat / bmm chain
            raise NotImplementedError("captum with radio_fusion='tensor' cannot run in the reference either")
        x = torch.cat([b.float() for b in bags], dim=2)
        B, N, width = x.shape
        x = Dense.apply(x.reshape(B * N, width), self.reduce_dim.weight, self.reduce_dim.bias, ACT_NONE)
        return self._captum_sum_pool(self.attention_net_radio, x.view(B, N, -1))

    def _captum_path(self, h_path):
        if 'path' not in self.mode:
            raise NotImplementedError('use another captum function')
        return self._captum_sum_pool(self.attention_net_WSI, h_path)

    def _captum_omic(self, h_omic):
        if 'omic' not in self.mode:
            raise NotImplementedError('use another captum function')
        return snn_forward(self.fc_omic, h_omic.float())

    def _captum_risk(self, vs):
        _, hid, Wk, bk = self._fuse(vs)
        _, S, _ = HazardHead.apply(hid, Wk, bk)
        return -torch.sum(S, dim=1)

    def captum_radio_omic(self, T1, T2, T1Gd, FLAIR, h_omic):
        given = dict(T1=T1, T2=T2, T1Gd=T1Gd, FLAIR=FLAIR)
        return self._captum_risk([self._captum_radio([given[m] for m in self.modalities]), self._captum_omic(h_omic)])

    def captum(self, T1, T2, T1Gd, FLAIR, h_path, h_omic):
        given = dict(T1=T1, T2=T2, T1Gd=T1Gd, FLAIR=FLAIR)
        return self._captum_risk([self._captum_radio([given[m] for m in self.modalities]), self._captum_path(h_path),
                                  self._captum_omic(h_omic)])

    def captum_radio_path(self, T1, T2, T1Gd, FLAIR, h_path):
        given = dict(T1=T1, T2=T2, T1Gd=T1Gd, FLAIR=FLAIR)
        M_radio, M_path = self._captum_radio([given[m] for m in self.modalities]), self._captum_path(h_path)
        if 'omic' in self.mode:
            raise NotImplementedError('use another captum function')
        return self._captum_risk([M_radio, M_path])

    def captum_path_omic(self, h_omic, h_path):
        M_path, O = self._captum_path(h_path), self._captum_omic(h_omic)
        if 'radio' in self.mode:
            raise NotImplementedError('use another captum function')
        return self._captum_risk([O, M_path])
