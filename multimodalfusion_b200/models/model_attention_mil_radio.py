"""Radiology attention-MIL survival model — drop-in for models/model_attention_mil_radio.py
(constructors :14-51,:67-71; forward :73-115).

The four modality bags are never concatenated: the reduce_dim GEMM reads them as four K-segments
through separate TMA descriptors and writes the bf16 [N,1024] bag the fused AMIL kernel consumes.
``radio_fusion='tensor'`` runs with the one-name repair of SURVEY.md App. B-3 (`xfusion` -> `radio_xfusion`).
"""
import torch
import torch.nn as nn

from ..autograd import HazardHead, reduce_dim_forward
from ..utils.utils import initialize_weights
from .model_modules import AmilBranch, Attn_Net, Attn_Net_Gated, XlinearFusion


class MIL_Attention_fc_radio(nn.Module):
    def __init__(self, radio_fusion='concat', gate_radio=True, dropout=True,
                 model_size_radio: str = 'small', n_classes=4, modalities=['T1', 'T2', 'T1Gd', 'FLAIR']):
        super().__init__()
        self.radio_fusion = radio_fusion
        self.n_classes = n_classes
        self.size_dict_radio = {"small": [1024, 256, 256], "big": [1024, 512, 384]}
        self.modalities = modalities
        in_dim, L, D = self.size_dict_radio[model_size_radio]
        if len(modalities) > 1:
            if radio_fusion == 'tensor':
                self.radio_xfusion = XlinearFusion(dim=1024, scale_dim=64, mmhid1=1024, mmhid2=1024, skip=0)
            elif radio_fusion == 'concat':
                self.reduce_dim = nn.Linear(in_dim * len(modalities), in_dim)
        attn_cls = Attn_Net_Gated if gate_radio else Attn_Net
        self.attention_net_radio = nn.Sequential(
            nn.Linear(in_dim, L), nn.ReLU(), nn.Dropout(0.25),
            attn_cls(L=L, D=D, dropout=dropout, n_classes=1))
        self.classifier = nn.Linear(L, n_classes)
        initialize_weights(self)
        self.bag_group = None

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if len(self.modalities) > 1:
            if self.radio_fusion == 'tensor':
                self.radio_xfusion = self.radio_xfusion.to(device)
            elif self.radio_fusion == 'concat':
                self.reduce_dim = self.reduce_dim.to(device)
        self.attention_net_radio = self.attention_net_radio.to(device)
        self.classifier = self.classifier.to(device)

    def forward(self, h, return_features=False, attention_only=False):
        pass


class MIL_Attention_fc_surv_radio(MIL_Attention_fc_radio):
    def __init__(self, radio_fusion='concat', gate_radio=True, dropout=True,
                 model_size_radio: str = 'small', n_classes=4, modalities=['T1', 'T2', 'T1Gd', 'FLAIR']):
        # the reference pins model_size_radio to 'small' regardless of the argument (:70)
        super().__init__(radio_fusion=radio_fusion, gate_radio=gate_radio, model_size_radio='small',
                         dropout=dropout, n_classes=n_classes, modalities=modalities)

    def forward(self, **kwargs):
        bags = [kwargs[m] for m in self.modalities]
        if len(bags) > 1:
            if self.radio_fusion == 'concat':
                x = reduce_dim_forward(self.reduce_dim.weight, self.reduce_dim.bias, bags)
            elif self.radio_fusion == 'tensor':
                # repaired semantics: the reference calls `self.xfusion` (:84) but defines
                # `radio_xfusion` (:29); read as written otherwise — slice 0 of every modality enters a
                # 4-way Kronecker fusion (17^4-wide, formed inside the encoder kernel) and the bag has ONE row
                x = self.radio_xfusion(v_list=[b[0].unsqueeze(0) for b in bags])
            else:
                raise NotImplementedError(f"radio_fusion={self.radio_fusion!r}")
        else:
            x = bags[0]
        A_raw, M = AmilBranch.pooled(self.attention_net_radio, x, self.training, self.bag_group)
        if kwargs.get('attention_only'):
            return A_raw
        if kwargs.get('return_features'):
            return M
        if kwargs.get('return_attention'):
            return A_raw
        hazards, S, Y_hat = HazardHead.apply(M, self.classifier.weight, self.classifier.bias)
        return hazards, S, Y_hat, A_raw

    # ---- fused batch-1 training step (utils/core_utils.py:200-247; radio loop :184-247) ------------------------------
    def enable_fused_step(self):
        """As MIL_Attention_fc_surv_path.enable_fused_step; reduce_dim keeps ordinary ``.grad`` tensors (16.8 MB: too large
        to be cleared by the two tile CTAs of a 155-slice bag)."""
        from . import _fused_step
        flat = _fused_step.enable(self, self.attention_net_radio, self.classifier)
        if hasattr(self, "reduce_dim"):
            for p_ in self.reduce_dim.parameters():
                p_.grad = torch.zeros_like(p_)
        return flat

    def graphed_fused_step(self, optimizer, Y, c, alpha=0.0, loss_scale=1.0, eps=1e-7, **bags):
        """One patient of the batch-1 radiology loop — fused_step + ``optimizer.step()`` (FusedAdam) — as ONE CUDA-graph
        launch per slice count (the ~29 kernels of the patient step cost ~16 us of host time each when launched from
        Python); see _fused_step.graphed."""
        from . import _fused_step
        return _fused_step.graphed(self, optimizer, {m: bags[m] for m in self.modalities}, Y, c, alpha, loss_scale, eps)

    def fused_window_step(self, patients, Y, c, alpha=0.0, accumulate=False, eps=1e-7):
        """The `gc` patients of one gradient-accumulation window (utils/core_utils.py:242-247) as ONE launch set: every
        modality's slices are packed varlen (patients on 128-row boundaries), reduce_dim runs once over the packed rows, the
        attention-MIL window step (ops.amil_window_step) returns the gradient of the reduced bags, reduce_dim's weight
        gradient is one tensor-core GEMM straight from the packed modality buffers. patients: list of {modality: bf16
        [N_i, 1024]} (stored-bf16 features; `radio_fusion='concat'`); Y, c: one entry per patient. Returns (hazards [gc,K],
        S [gc,K], Y_hat [gc,1], [A_raw_i [1,N_i]], loss [gc] (unscaled)); gradients (sum of loss_i / gc) in ``.grad``."""
        from .. import ops
        from .._lib import ACT_NONE
        from .model_modules import _seed_from_torch
        if not hasattr(self, "_fused"):
            self.enable_fused_step()
        if self.bag_group is not None or len(self.modalities) < 2 or self.radio_fusion != 'concat':
            raise NotImplementedError("fused_window_step: whole bags, several modalities, radio_fusion='concat'")
        if any(p_[m].dtype != torch.bfloat16 for p_ in patients for m in self.modalities):
            raise NotImplementedError("fused_window_step packs stored-bf16 features; fp32 bags go through fused_step")
        f = self._fused
        seq, attn = self.attention_net_radio, self.attention_net_radio[3]
        prep = AmilBranch.prepared(seq)
        flags = ops.amil_flags(prep.gated, dropout_h=self.training, dropout_attn=self.training and attn.use_dropout)
        packed = []
        for m in self.modalities:
            packed.append(ops.pack_bags([p_[m] for p_ in patients], index_from=packed[0] if packed else None))
        Wr, br = self.reduce_dim.weight, self.reduce_dim.bias
        # reduce_dim over every packed row on the tensor cores, straight from the packed modality buffers: the slices are
        # bf16 values already, the weight goes in as a bf16 hi + lo pair (W = hi + lo to 2^-17 relative: fp32-grade products,
        # fp32 accumulation) — two segmented GEMM passes instead of a 43-GFLOP fp32 SGEMM (1.7 ms per 32 patients)
        Wh = ops.to_bf16(Wr.detach())
        Wl = ops.to_bf16(Wr.detach() - Wh.float())
        segs = [pk.x for pk in packed]
        h0 = ops.linear_bf16(segs, Wh, br.detach(), out_dtype=torch.float32)
        h0 += ops.linear_bf16(segs, Wl, None, out_dtype=torch.float32)
        win = ops.PackedBags(h0, packed[0].tile_valid, packed[0].seg_tile_offsets, packed[0].row_offsets, packed[0].sizes,
                             packed[0].tile_bag, packed[0].tile_valid_even)
        out = ops.amil_window_step(win, prep, flags, _seed_from_torch() if self.training else 0,
                                   self.classifier.weight.detach(), self.classifier.bias.detach(), Y, c, alpha, f["grads"],
                                   dWk=f["dWk"], dbk=f["dbk"], eps=eps, loss_scale=1.0 / len(patients),
                                   zero=None if accumulate else f["flat"], need_dx=True)
        if not accumulate:
            Wr.grad.zero_(); br.grad.zero_()
        ops.linear_bf16_wgrad(out["dx"], [pk.x for pk in packed], Wr.grad, br.grad)   # padding rows: dx = 0
        A = [out["A_raw"][o:o + n].view(1, n) for o, n in zip(win.row_offsets, win.sizes)]
        return out["hazards"], out["S"], out["Y_hat"], A, out["loss"]

    def fused_step(self, Y, c, alpha=0.0, loss_scale=1.0, accumulate=False, eps=1e-7, **bags):
        """One patient of the reference's batch-1 radiology loop without an autograd graph: ``model(T1=.., T2=.., ...)`` ->
        nll_surv -> ``(loss * loss_scale).backward()`` as reduce_dim (one fp32 functor-SGEMM launch on the concatenated
        slices) + the library's fused step on the reduced bag (3 launches + the dx GEMM) + reduce_dim's weight gradient
        (one launch, accumulated in place): ~9 launches per patient instead of ~40 through autograd. ``radio_fusion=
        'concat'`` (or a single modality) and bags of up to 4096 slices; returns (hazards, S, Y_hat, A_raw, loss)."""
        from .. import ops
        from .._lib import ACT_NONE
        from . import _fused_step
        if not hasattr(self, "_fused"):
            self.enable_fused_step()
        if self.bag_group is not None:
            raise NotImplementedError("fused_step runs whole bags; instance-sharded bags go through forward()")
        xs = [bags[m] for m in self.modalities]
        from ..autograd import AmilPool
        if AmilPool.precise_small_bags and xs[0].shape[0] <= AmilBranch.TINY_BAG_FP32_ROWS:
            # tiny bags (real radiology patients: 17-40 slices) keep the exact-fp32 kernels of forward() — a pooled
            # embedding exact to fp32 keeps every ReLU unit on the reference's side (model_modules.AmilBranch) — through
            # the autograd.Functions; same gradient buffers, same return values
            from ..utils.loss_utils import nll_loss
            if not accumulate:
                self.zero_grad(set_to_none=False)
            hz, S, Y_hat, A_raw = self(**bags)
            loss = nll_loss(hz, S, Y, c, alpha=alpha, eps=eps)
            (loss * loss_scale).backward()
            return hz.detach(), S.detach(), Y_hat, A_raw.detach(), loss.detach()
        if len(xs) == 1:
            return _fused_step.run(self, self.attention_net_radio, self.classifier, xs[0], Y, c, alpha, loss_scale,
                                   accumulate, eps)[:5]
        if self.radio_fusion != 'concat' or xs[0].shape[0] > ops.PRECISE_FC_MAX_ROWS:
            raise NotImplementedError("fused_step: radio_fusion='concat' with at most 4096 slices; use forward() + autograd")
        Wr, br = self.reduce_dim.weight, self.reduce_dim.bias
        # [N, 1024 m] fp32: one concat in the bags' dtype + one conversion (a conversion per modality + the concat were five
        # launches of ~4 us each)
        xc = (torch.cat(xs, dim=1) if len({b.dtype for b in xs}) == 1 else torch.cat([b.float() for b in xs], dim=1)).float()
        h0 = ops.dense_fwd(xc, Wr.detach(), br.detach(), ACT_NONE)             # reduce_dim, fp32 (models/...radio.py:81-82)
        out = _fused_step.run(self, self.attention_net_radio, self.classifier, h0, Y, c, alpha, loss_scale, accumulate,
                              eps, need_dx=True)
        if not accumulate:
            Wr.grad.zero_(); br.grad.zero_()
        dh0 = out[5]                                                            # bf16 [N, 1024] (the AMIL dx GEMM's output)
        if all(b.dtype == torch.bfloat16 and b.is_contiguous() for b in xs):
            # dWr += dh0^T cat(bags), dbr += colsum(dh0) on the tensor cores: BOTH operands are bf16 values already (the
            # stored bags and the dx GEMM's output), so their products are exact in the fp32 accumulator — the same
            # gradient as the fp32 SGEMM up to summation order, read straight from the four modality bags (47 -> ~10 us)
            ops.linear_bf16_wgrad(dh0, xs, Wr.grad, br.grad)
        else:
            # fp32 bags: accumulated in place by the functor SGEMM
            ops.dense_bwd_into(xc, Wr.detach(), h0, dh0.float(), Wr.grad, br.grad)
        return out[:5]

