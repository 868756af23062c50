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
    contributions (SUM over ranks); everything downstream of the replicated M is already identical.
    The collective is unconditional and has the same size on every rank: a rank whose shard is empty (fewer
    256-row units than ranks) has no gradient for these parameters and contributes zeros — selecting on
    ``p.grad is not None`` would leave it out of the all-reduce and hang the others."""
    params = [p for n, p in model.named_parameters()
              if p.requires_grad and (n.startswith("attention_net_") or n.startswith("reduce_dim"))]
    if not params:
        return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    allreduce_sum_(grads, group)
    for p, g in zip(params, grads):
        if p.grad is None:
            p.grad = g


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


class PeerAllReduce:
    """SUM all-reduce of flat fp32 buffers over NVLink peer memory with the library's own kernel
    (``mmf_p2p_allreduce_sum_f32``: ready handshake, reduce-scatter + all-gather by peer loads / stores, done
    handshake — one launch on the caller's stream, CUDA-graph capturable). torch's symmetric memory is
    used only to allocate the buffers and to map them into every rank (plumbing).

        ar = PeerAllReduce(numel, n_buffers=2)      # collective: every rank of the group
        ar.buffer(0)                                 # flat fp32 tensor the backward accumulates into
        ar.all_reduce(0)                             # in place, on the current stream

    Raises if symmetric memory cannot be set up (no P2P); callers then use ``dist.all_reduce``.
    """

    def __init__(self, numel: int, n_buffers: int = 1, group=None, n_ctas: int = 0, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerAllReduce supports up to 8 ranks (one NVSwitch domain)")
        self.numel = (numel + 3) // 4 * 4
        self.n_ctas = n_ctas
        dev = torch.device("cuda", torch.cuda.current_device())
        flag_words = _lib.lib().mmf_p2p_flag_bytes() // 4
        self._bufs, self._flags, self._ptrs, self._flag_ptrs, self._mc = [], [], [], [], []
        for _ in range(n_buffers):
            buf = symm_mem.empty(self.numel, dtype=torch.float32, device=dev)
            flg = symm_mem.empty(flag_words, dtype=torch.int32, device=dev)
            buf.zero_(); flg.zero_()
            hb = symm_mem.rendezvous(buf, self.group)
            hf = symm_mem.rendezvous(flg, self.group)
            self._bufs.append(buf); self._flags.append(flg)
            self._ptrs.append(_lib.ptr_array([int(p) for p in hb.buffer_ptrs]))
            self._flag_ptrs.append(_lib.ptr_array([int(p) for p in hf.buffer_ptrs]))
            mc = int(getattr(hb, "multicast_ptr", 0) or 0) if use_multicast else 0
            self._mc.append(mc if mc else None)   # NVLS multicast mapping when the fabric supports it
            self._keep = getattr(self, "_keep", []) + [hb, hf]
        if self.n_ctas <= 0:
            # the exchange is latency-bound (~27 us at 3.7 MB whatever the CTA count): with NVLS 4 CTAs (2 TPCs) suffice
            # and leave the SMs to the step's own kernels when the all-reduce overlaps them; peer loads want more
            self.n_ctas = 4 if self.multicast else 16
        torch.cuda.synchronize()
        dist.barrier(self.group)   # every rank's flags are zeroed before the first handshake

    @property
    def multicast(self) -> bool:
        return all(m is not None for m in self._mc)

    def buffer(self, i: int = 0) -> torch.Tensor:
        return self._bufs[i]

    def all_reduce(self, i: int = 0) -> None:
        from . import _lib
        _lib.check(_lib.lib().mmf_p2p_allreduce_sum_f32(self._ptrs[i], self._flag_ptrs[i], self._mc[i], self.world, self.rank,
                                                        self.numel, self.n_ctas,
                                                        torch.cuda.current_stream().cuda_stream),
                   "mmf_p2p_allreduce_sum_f32")
