"""Pathology attention-MIL survival model — drop-in for
models/model_attention_mil_path.py (constructors :13-34,:46; forward :50-72).

forward(path_features=[N,1024]) runs two kernels for the bag (fused tile kernel + combine) and one
for the discrete-hazard head; the [N,L] activations never reach HBM in the forward.
"""
import torch
import torch.nn as nn

from ..autograd import HazardHead
from ..utils.utils import initialize_weights
from .model_modules import AmilBranch, Attn_Net, Attn_Net_Gated


class MIL_Attention_fc_path(nn.Module):
    def __init__(self, gate_path=True, dropout=True, model_size_wsi: str = 'small', n_classes=4):
        super().__init__()
        self.size_dict_WSI = {"small": [1024, 256, 256], "big": [1024, 512, 384]}
        in_dim, L, D = self.size_dict_WSI[model_size_wsi]
        attn_cls = Attn_Net_Gated if gate_path else Attn_Net
        self.attention_net_WSI = nn.Sequential(
            nn.Linear(in_dim, L), nn.ReLU(), nn.Dropout(0.25),
            attn_cls(L=L, D=D, dropout=dropout, n_classes=1))
        self.classifier = nn.Linear(L, n_classes)
        initialize_weights(self)
        self.bag_group = None  # set to a torch.distributed group to shard one bag across ranks

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.attention_net_WSI = self.attention_net_WSI.to(device)
        self.classifier = self.classifier.to(device)

    def forward(self, h, return_features=False, attention_only=False):
        pass


class MIL_Attention_fc_surv_path(MIL_Attention_fc_path):
    def __init__(self, gate_path=True, model_size_wsi: str = 'small', dropout=False, n_classes=4):
        super().__init__(gate_path=gate_path, model_size_wsi=model_size_wsi, dropout=dropout,
                         n_classes=n_classes)

    def forward(self, **kwargs):
        x = kwargs['path_features']
        A_raw, M = AmilBranch.pooled(self.attention_net_WSI, x, self.training, self.bag_group)
        if kwargs.get('return_features'):
            return M
        if kwargs.get('attention_only'):
            return A_raw
        hazards, S, Y_hat = HazardHead.apply(M, self.classifier.weight, self.classifier.bias)
        return hazards, S, Y_hat, A_raw

    # ---- fused batch-1 training step (utils/core_utils.py:200-247 in three launches) ---------------------------
    def enable_fused_step(self):
        """Allocates ONE flat fp32 gradient buffer and makes every parameter's ``.grad`` a view of it (the layout the
        fused step's kernels accumulate into; ``optimizer.zero_grad(set_to_none=False)`` / ``FusedAdam`` keep the
        views). Call after ``.cuda()`` / ``relocate()``."""
        from . import _fused_step
        return _fused_step.enable(self, self.attention_net_WSI, self.classifier)

    def fused_step(self, path_features, Y, c, alpha=0.0, loss_scale=1.0, accumulate=False, eps=1e-7):
        """``hazards, S, Y_hat, A_raw = model(path_features=x); loss = nll_surv(...); (loss * loss_scale).backward()``
        as THREE kernel launches and no autograd graph: fused forward, fused gate + hidden backward with the head
        (combine, classifier, hazards, loss, their gradients) in its prologue, grouped wgrad. Gradients land in the
        parameters' ``.grad`` (views of the flat buffer of ``enable_fused_step``); ``accumulate=False`` clears them
        first inside the forward kernel (``optimizer.zero_grad()``), ``True`` adds (gradient accumulation over ``gc``
        bags, ``loss_scale = 1 / gc``). Train mode only; bags of up to 65536 instances; returns
        (hazards [1,K], S [1,K], Y_hat [1,1], A_raw [1,N], loss) — views of buffers reused by the next call."""
        from . import _fused_step
        if not hasattr(self, "_fused"):
            self.enable_fused_step()
        if self.bag_group is not None:
            raise NotImplementedError("fused_step runs whole bags; instance-sharded bags go through forward()")
        return _fused_step.run(self, self.attention_net_WSI, self.classifier, path_features, Y, c, alpha, loss_scale,
                               accumulate, eps)[:5]

    def fused_window_step(self, bags, Y, c, alpha=0.0, accumulate=False, eps=1e-7):
        """The `gc` bags of one gradient-accumulation window (utils/core_utils.py:242-247: ``loss / gc`` -> backward per bag,
        optimizer step every gc bags) as ONE launch set instead of gc steps: the bags are packed varlen (each on a 128-row
        boundary), the fused forward runs every tile of the window, one head launch evaluates every bag's classifier /
        hazards / nll_surv (scaled by 1 / gc), the backward accumulates the window's summed gradients into the parameters'
        ``.grad``. bags: list of [N_i, 1024] CUDA tensors; Y, c: one entry per bag. Train-mode dropout draws one mask per
        packed row. Returns (hazards [gc,K], S [gc,K], Y_hat [gc,1], [A_raw_i [1,N_i]], loss [gc] (unscaled))."""
        from .. import ops
        from . import _fused_step
        from .model_modules import _seed_from_torch
        if not hasattr(self, "_fused"):
            self.enable_fused_step()
        if self.bag_group is not None:
            raise NotImplementedError("fused_window_step runs whole bags")
        f = self._fused
        seq, attn = self.attention_net_WSI, self.attention_net_WSI[3]
        prep = AmilBranch.prepared(seq)
        flags = ops.amil_flags(prep.gated, dropout_h=self.training, dropout_attn=self.training and attn.use_dropout)
        packed = ops.pack_bags(bags)
        out = ops.amil_window_step(packed, prep, flags, _seed_from_torch() if self.training else 0,
                                   self.classifier.weight.detach(), self.classifier.bias.detach(), Y, c, alpha, f["grads"],
                                   dWk=f["dWk"], dbk=f["dbk"], eps=eps, loss_scale=1.0 / len(bags),
                                   zero=None if accumulate else f["flat"])
        A = [out["A_raw"][o:o + n].view(1, n) for o, n in zip(packed.row_offsets, packed.sizes)]
        return out["hazards"], out["S"], out["Y_hat"], A, out["loss"]

    def graphed_fused_step(self, optimizer, path_features, Y, c, alpha=0.0, loss_scale=1.0, eps=1e-7):
        """fused_step + ``optimizer.step()`` (FusedAdam) as one CUDA-graph launch per bag size; see _fused_step.graphed."""
        from . import _fused_step
        return _fused_step.graphed(self, optimizer, {"path_features": path_features}, Y, c, alpha, loss_scale, eps)

    @torch.no_grad()
    def infer_cohort(self, bags):
        """Eval-mode forward of MANY slides at once (new capability; the reference loops batch-1, e.g.
        pre_trained_feature.py:116-162 / create_heatmaps.py): the bags are packed into one varlen buffer and run as
        one fused-forward launch + one head launch. Returns (hazards [n,K], S [n,K], Y_hat [n,1], [A_raw_i [1,N_i]])
        — row i equals forward(path_features=bags[i]) in eval mode."""
        from .. import ops
        fc, attn = self.attention_net_WSI[0], self.attention_net_WSI[3]
        prep = ops.prepare_amil_weights(fc.weight, fc.bias, *attn.amil_weights())
        packed = ops.pack_bags(bags)
        out = ops.amil_infer_varlen(packed, prep, self.classifier.weight, self.classifier.bias)
        A = [out["A_raw"][o:o + n].view(1, n) for o, n in zip(packed.row_offsets, packed.sizes)]
        return out["hazards"], out["S"], out["Y_hat"], A
