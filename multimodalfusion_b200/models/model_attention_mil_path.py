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
        from .. import ops
        fc, attn = self.attention_net_WSI[0], self.attention_net_WSI[3]
        Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
        order = [fc.weight, fc.bias, Wa] + ([Wb] if Wb is not None else []) + [ba] + ([bb] if bb is not None else []) \
            + [wc, bc, self.classifier.weight, self.classifier.bias]
        dev = fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("enable_fused_step needs the model on a CUDA device (no CPU fallback)")
        n = sum(p.numel() for p in order)
        flat = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=dev)
        o = 0
        for p_ in order:
            p_.grad = flat[o:o + p_.numel()].view_as(p_)
            o += p_.numel()
        L, D = fc.weight.shape[0], Wa.shape[0]
        KD = D * (2 if Wb is not None else 1)
        self._fused = dict(flat=flat, L=L, D=D, KD=KD, gated=Wb is not None, bufs={})
        g = {}
        o = 0
        for name, cnt, shape in (("dW1", L * 1024, (L, 1024)), ("db1", L, (L,)), ("dWab", KD * L, (KD, L)),
                                 ("dbab", KD, (KD,)), ("dwc", D, (D,)), ("dbc", 1, (1,))):
            g[name] = flat[o:o + cnt].view(shape)
            o += cnt
        K = self.classifier.weight.shape[0]
        self._fused.update(grads=g, dWk=flat[o:o + K * L].view(K, L), dbk=flat[o + K * L:o + K * L + K])
        return flat

    def fused_step(self, path_features, Y, c, alpha=0.0, loss_scale=1.0, accumulate=False, eps=1e-7):
        """``hazards, S, Y_hat, A_raw = model(path_features=x); loss = nll_surv(...); (loss * loss_scale).backward()``
        as THREE kernel launches and no autograd graph: fused forward, fused gate + hidden backward with the head
        (combine, classifier, hazards, loss, their gradients) in its prologue, grouped wgrad. Gradients land in the
        parameters' ``.grad`` (views of the flat buffer of ``enable_fused_step``); ``accumulate=False`` clears them
        first inside the forward kernel (``optimizer.zero_grad()``), ``True`` adds (gradient accumulation over ``gc``
        bags, ``loss_scale = 1 / gc``). Train mode only; bags of up to 65536 instances; returns
        (hazards [1,K], S [1,K], Y_hat [1,1], A_raw [1,N], loss) — views of buffers reused by the next call."""
        from .. import ops
        from .model_modules import AmilBranch, _seed_from_torch
        if not hasattr(self, "_fused"):
            self.enable_fused_step()
        if self.bag_group is not None:
            raise NotImplementedError("fused_step runs whole bags; instance-sharded bags go through forward()")
        f = self._fused
        seq = self.attention_net_WSI
        prep = AmilBranch.prepared(seq)
        N = path_features.shape[0]
        if N > 65536:
            raise NotImplementedError("fused_step merges at most 512 per-tile head rows per CTA (N <= 65536)")
        attn = seq[3]
        flags = ops.amil_flags(prep.gated, dropout_h=self.training, dropout_attn=self.training and attn.use_dropout)
        if N <= ops.PRECISE_FC_MAX_ROWS:      # small bag: split-precision fc (see autograd.AmilPool)
            x = ops.split_bag(path_features)
            flags |= ops.MMF_PRECISE_FC
        else:
            x = ops.to_bf16(path_features)
        seed = _seed_from_torch() if self.training else 0
        K = self.classifier.weight.shape[0]
        buf = f["bufs"].get(N)
        if buf is None:
            if len(f["bufs"]) >= 4:     # bags come in many sizes: keep a few workspaces, not one per size
                f["bufs"].pop(next(iter(f["bufs"])))
            buf = f["bufs"][N] = ops.FusedStepBuffers(N, prep, flags, K, x.device)
        Yd = Y.detach().reshape(-1).to(device=x.device, dtype=torch.int64)
        cd = c.detach().reshape(-1).to(device=x.device, dtype=torch.float32)
        ops.amil_fused_step(x, prep, flags, seed, buf, self.classifier.weight.detach(), self.classifier.bias.detach(),
                            Yd, cd, alpha, f["grads"], dWk=f["dWk"], dbk=f["dbk"], eps=eps, loss_scale=loss_scale,
                            zero=None if accumulate else f["flat"])
        return buf.hazards, buf.S, buf.Y_hat, buf.A_raw.view(1, -1), buf.loss

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
