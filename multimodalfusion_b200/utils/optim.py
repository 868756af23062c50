"""Training-step glue on the device: fused multi-tensor Adam with the reference's
weight decay and l1 regulariser folded in, and the concordance index.

`get_optim(model, args)` mirrors utils/utils.py:144-151 of the reference (`--opt adam|sgd`, `--lr`, `--reg`)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from .._lib import check, lib, ptr_array


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics, one kernel launch per step for ALL parameters
    (`mmf_adam_step_multi`). Extras: ``l1_lambda`` adds the reference's l1_reg_all penalty gradient
    lambda * sign(W) inside the update (utils/utils.py:249-257, utils/core_utils.py:218-221) and accumulates the
    penalty value in ``self.l1_value`` (device scalar, sum |W| before the update); ``grad_scale`` folds the 1/gc of
    gradient accumulation; ``step(zero_grad=True)`` clears the gradients in the same launch."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, l1_lambda: float = 0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, l1_lambda=l1_lambda))
        self.l1_value: Optional[torch.Tensor] = None
        self._graph_steps = 0    # steps taken by CUDA-graph replays, not yet folded into the per-parameter "step" entries

    # ---- graph-captured steps (multimodalfusion_b200.graphs): the step count lives on the device -------------------
    def note_graph_step(self):
        self._graph_steps += 1

    def flush_graph_steps(self):
        if self._graph_steps:
            for st in self.state.values():
                if "step" in st:
                    st["step"] += self._graph_steps
            self._graph_steps = 0

    def host_step(self) -> int:
        self.flush_graph_steps()
        return max((st.get("step", 0) for st in self.state.values()), default=0)

    def state_dict(self):
        self.flush_graph_steps()
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False, grad_scale: float = 1.0):
        from ..graphs import current_state
        dev_state = current_state()      # not None: this call is being captured into a CUDA graph
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if dev_state is None:
            self.flush_graph_steps()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                if dev_state is None:
                    st["step"] += 1
            step = self.state[ps[0]]["step"]
            dev = ps[0].device
            if self.l1_value is None or self.l1_value.device != dev:
                self.l1_value = torch.zeros((), dtype=torch.float32, device=dev)
            self.l1_value.zero_()
            numel = (C.c_int64 * len(ps))(*[p.numel() for p in ps])
            fn = lib().mmf_adam_step_multi if dev_state is None else lib().mmf_adam_step_multi_dev
            check(fn(
                ptr_array([p.data_ptr() for p in ps]), ptr_array([p.grad.data_ptr() for p in ps]),
                ptr_array([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                ptr_array([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]), numel, len(ps),
                step if dev_state is None else dev_state.step_ptr, float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                float(group["weight_decay"]), float(grad_scale), float(group["l1_lambda"]), int(zero_grad),
                self.l1_value.data_ptr(), torch.cuda.current_stream().cuda_stream), "mmf_adam_step_multi")
            # parameters changed in place behind autograd's back: bump the version counters (no kernel) so that the
            # cached bf16 weight copies (AmilBranch) are rebuilt
            torch._C._autograd._unsafe_set_version_counter(tuple(ps), tuple(p._version + 1 for p in ps))
        return loss


def get_optim(model, args):
    """utils/utils.py:144-151."""
    params = [p for p in model.parameters() if p.requires_grad]
    if args.opt == "adam":
        return FusedAdam(params, lr=args.lr, weight_decay=args.reg)
    if args.opt == "sgd":
        return torch.optim.SGD(params, lr=args.lr, momentum=0.9, weight_decay=args.reg)
    raise NotImplementedError


def concordance_index(risk: torch.Tensor, times: torch.Tensor, event: torch.Tensor, tied_tol: float = 1e-8) -> float:
    """Harrell's C as sksurv.concordance_index_censored(event, time, risk)[0] (utils/core_utils.py:258), counted on the
    device over the O(B^2) pair grid. `event` = 1 - censorship."""
    if not risk.is_cuda:
        raise RuntimeError("multimodalfusion_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback.")
    r = risk.detach().reshape(-1).float().contiguous()
    t = torch.as_tensor(times).detach().reshape(-1).to(r.device, torch.float32).contiguous()
    e = torch.as_tensor(event).detach().reshape(-1).to(r.device, torch.float32).contiguous()
    counts = torch.zeros(3, dtype=torch.int64, device=r.device)
    check(lib().mmf_cindex_counts(r.data_ptr(), t.data_ptr(), e.data_ptr(), r.numel(), float(tied_tol),
                                  counts.data_ptr(), torch.cuda.current_stream().cuda_stream), "mmf_cindex_counts")
    c, d, tie = counts.tolist()
    tot = c + d + tie
    return (c + 0.5 * tie) / tot if tot else float("nan")


def to_percentiles(scores: torch.Tensor, ref_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils/wsi_utils.py:171-174 (`to_percentiles(scores)`) and utils/heatmap_utils.py:32-34 (`score2percentile`
    against reference scores): scipy's percentileofscore (kind='rank') of every score, on the device."""
    if not scores.is_cuda:
        raise RuntimeError("multimodalfusion_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback.")
    q = scores.detach().reshape(-1).float().contiguous()
    ref = q if ref_scores is None else ref_scores.detach().reshape(-1).to(q.device, torch.float32).contiguous()
    out = torch.empty_like(q)
    check(lib().mmf_percentile_of_score(ref.data_ptr(), ref.numel(), q.data_ptr(), q.numel(), out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "mmf_percentile_of_score")
    return out.view(scores.shape)
