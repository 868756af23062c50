"""CPU port of the reference's training step for the pathology AMIL model, used ONLY as the timed
CPU baseline (bench.py `cpu_baseline` and `--impl reference`) — test/bench infrastructure, never a
product path.

The reference tree is not present on the GPU box (it cannot travel), so this restates, with the same
stock ATen ops the reference dispatches to, what one iteration of its hot loop does:
  model(**features) -> NLLSurvLoss -> loss.backward()         (utils/core_utils.py:200-247)
  model = MIL_Attention_fc_surv_path                         (models/model_attention_mil_path.py:13-72)
fp32, autograd, all host threads, train mode (Dropout(0.25) on h active, as the reference trains).
tests/test_oracle.py pins the same math (oracle.amil_oracle) to the reference's golden outputs.
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn.functional as F

from . import amil_oracle as O


class CpuAmilStep:
    def __init__(self, L=512, D=384, K=4, gated=True, seed=0):
        g = torch.Generator().manual_seed(seed)

        def xavier(o, i):
            return (torch.randn(o, i, generator=g) * (2.0 / (o + i)) ** 0.5).requires_grad_(True)

        self.W1, self.b1 = xavier(L, 1024), torch.zeros(L, requires_grad=True)
        self.Wa, self.ba = xavier(D, L), torch.zeros(D, requires_grad=True)
        self.Wb, self.bb = (xavier(D, L), torch.zeros(D, requires_grad=True)) if gated else (None, None)
        self.wc, self.bc = xavier(1, D), torch.zeros(1, requires_grad=True)
        self.Wk, self.bk = xavier(K, L), torch.zeros(K, requires_grad=True)
        self.params = [p for p in (self.W1, self.b1, self.Wa, self.ba, self.Wb, self.bb, self.wc, self.bc,
                                   self.Wk, self.bk) if p is not None]

    def step(self, x: torch.Tensor, Y: torch.Tensor, c: torch.Tensor, alpha: float = 0.0, train: bool = True):
        for p in self.params:
            p.grad = None
        h = F.dropout(torch.relu(F.linear(x, self.W1, self.b1)), 0.25, train)
        a = torch.tanh(F.linear(h, self.Wa, self.ba))
        if self.Wb is not None:
            a = a * torch.sigmoid(F.linear(h, self.Wb, self.bb))
        A = F.linear(a, self.wc, self.bc).t()
        M = torch.mm(F.softmax(A, dim=1), h)
        hazards, S, _ = O.hazard_head(M, self.Wk, self.bk)
        loss = O.nll_surv_loss(hazards, S, Y, c, alpha=alpha)
        loss.backward()
        return loss.item(), float(-S.detach().sum())


def time_cpu_steps(N=16384, L=512, D=384, K=4, steps=5, warmup=1, threads=None, seed=0):
    """Returns (patches_per_s, seconds_per_step, threads). One step = fwd+bwd over one N x 1024 bag."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = CpuAmilStep(L, D, K, True, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = (0.5 * torch.randn(N, 1024, generator=g).abs()).to(torch.bfloat16).float()
    Y, c = torch.tensor([2]), torch.tensor([0.0])
    for _ in range(warmup):
        model.step(x, Y, c)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.step(x, Y, c)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return N / dt, dt, threads
