"""Multi-GPU plumbing for the two ways the path shards, one process per GPU over
torch.distributed (NCCL on the B200 box, gloo in the CPU tests):

1. instance sharding of ONE giant bag: every rank runs the fused tile kernel on its row range, the
   rank-local (m, l, acc[L]) partial (L+2 floats) is all-gathered and combined on every rank, so the
   pooled vector M is replicated; the backward needs no further exchange until the weight-gradient
   SUM all-reduce;
2. cohort data parallelism: independent bags are dealt to ranks by size; per-patient risks are
   all-gathered for cohort losses (Cox / ranking need every risk), weight gradients are all-reduced.

Nothing here computes on tensors itself — combine functions are passed in (the CUDA kernels in the
product, the oracle in the CPU tests) so the protocol is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

PAIR_ROWS = 256  # a CTA pair owns 256 instances; shard boundaries are aligned to it


def shard_rows(n_rows: int, rank: int, world: int, align: int = PAIR_ROWS) -> Tuple[int, int]:
    """Contiguous, aligned, near-equal row range [lo, hi) of rank `rank`; ranges tile [0, n_rows)."""
    units = (n_rows + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * align, n_rows), min(hi_u * align, n_rows)


def deal_cohort(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time deal of bags to ranks: indices per rank, deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += sizes[i]
    return out


def all_gather_combine(local_partial: torch.Tensor, combine: Callable[[torch.Tensor], tuple], group=None):
    """local_partial: [L+2] = (m, l, acc) of this rank's rows (m = -inf, l = 0 for an empty shard).
    Returns combine(stack of all ranks' partials) — identical on every rank."""
    world = dist.get_world_size(group)
    gathered = torch.empty(world, local_partial.numel(), dtype=local_partial.dtype, device=local_partial.device)
    dist.all_gather_into_tensor(gathered, local_partial.reshape(1, -1).contiguous(), group=group)
    return combine(gathered)


def empty_partial(L: int, device=None) -> torch.Tensor:
    p = torch.zeros(L + 2, dtype=torch.float32, device=device)
    p[0] = float("-inf")
    return p


def allreduce_sum_(tensors: Sequence[torch.Tensor], group=None) -> None:
    """One flat SUM all-reduce over a list of gradient tensors (in place)."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()


def sync_sharded_bag_grads(model: torch.nn.Module, group=None) -> None:
    """After backward of an instance-sharded bag: fc / attention-net gradients are per-shard
    contributions (SUM over ranks); everything downstream of the replicated M is already identical."""
    amil = [p.grad for n, p in model.named_parameters()
            if p.grad is not None and (n.startswith("attention_net_") or n.startswith("reduce_dim"))]
    allreduce_sum_(amil, group)


def sync_cohort_grads(model: torch.nn.Module, group=None, average: bool = True) -> None:
    """Cohort data parallelism: all parameters' gradients are summed (averaged) over ranks."""
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    allreduce_sum_(grads, group)
    if average:
        w = dist.get_world_size(group)
        for g in grads:
            g.div_(w)


def gather_risks(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather of per-patient risks with uneven counts per rank (padded to the max)."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros(mx, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local.reshape(-1)
    out = torch.empty(world, mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.reshape(1, -1), group=group)
    return torch.cat([out[r, :counts[r]] for r in range(world)])
