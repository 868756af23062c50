"""CPU, world_size 2 over gloo: the host-side protocol of both sharding schemes with
the oracle standing in for the kernels — shard ranges, all-gather + combine of (m, l, acc) partials,
per-shard backward + SUM all-reduce == whole-bag gradients, cohort dealing, risk all-gather + Cox."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalfusion_b200 import parallel as P
from oracle import amil_oracle as O
from oracle import cases


def test_shard_rows_tile_the_bag():
    for n in (1, 255, 256, 257, 1000, 16384, 262144):
        for world in (1, 2, 3, 4, 8):
            spans = [P.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for lo, hi in spans:
                if hi > lo:
                    assert lo % P.PAIR_ROWS == 0 and (hi % P.PAIR_ROWS == 0 or hi == n)
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 2 * P.PAIR_ROWS


def test_deal_cohort_balances_and_is_deterministic():
    g = torch.Generator().manual_seed(0)
    sizes = torch.exp(torch.randn(1000, generator=g) * 0.8 + 9).clamp(500, 64000).long().tolist()
    deal = P.deal_cohort(sizes, 8)
    assert sorted(i for r in deal for i in r) == list(range(1000))
    loads = [sum(sizes[i] for i in r) for r in deal]
    assert max(loads) / (sum(loads) / 8) < 1.01
    assert deal == P.deal_cohort(sizes, 8)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        N, L, D = 1000, 256, 256
        x = cases.features(N, 5)
        g = torch.Generator().manual_seed(1)
        W1, b1 = torch.randn(L, 1024, generator=g) * 0.03, torch.randn(L, generator=g) * 0.05
        Wa, ba = torch.randn(D, L, generator=g) * 0.06, torch.randn(D, generator=g) * 0.05
        Wb, bb = torch.randn(D, L, generator=g) * 0.06, torch.randn(D, generator=g) * 0.05
        wc, bc = torch.randn(1, D, generator=g) * 0.1, torch.zeros(1)
        dM = torch.randn(L, generator=g)
        # whole-bag reference
        s, h, a, gg = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc)
        M, m, l = O.softmax_pool(s, h)
        full = O.amil_backward(x, W1, Wa, Wb, wc, s, h, a, gg, M, m, l, dM)
        # this rank's shard
        lo, hi = P.shard_rows(N, rank, world)
        xs = x[lo:hi]
        ss, hs, as_, gs = O.fc_attention(xs, W1, b1, Wa, ba, Wb, bb, wc, bc)
        local = O.combine_partials(O.tile_partials(ss, hs), normalize=False) if hi > lo else P.empty_partial(L)
        Mg, mg, lg = P.all_gather_combine(local, lambda parts: O.combine_partials(parts, True))
        assert torch.allclose(Mg, M, rtol=1e-5, atol=1e-6) and abs(mg - m) < 1e-6 and abs(lg / l - 1) < 1e-5
        grads = O.amil_backward(xs, W1, Wa, Wb, wc, ss, hs, as_, gs, Mg, mg, lg, dM)
        keys = ["dW1", "db1", "dWab", "dbab", "dwc", "dbc"]
        P.allreduce_sum_([grads[k] for k in keys])
        for k in keys:
            err = (grads[k] - full[k]).abs().max()
            assert err < 1e-4 * full[k].abs().max() + 2e-5, (k, float(err))
        # cohort: deal 13 patients, all-gather risks, Cox on every rank == single-process Cox
        B = 13
        risks = torch.randn(B, generator=torch.Generator().manual_seed(7))
        times, c = cases.cohort_labels(B, 3)
        deal = P.deal_cohort([100 + 7 * i for i in range(B)], world)
        mine = torch.tensor([risks[i] for i in deal[rank]])
        allr = P.gather_risks(mine, [len(d) for d in deal])
        order = [i for d in deal for i in d]
        loss = O.cox_loss(allr, times[order], c[order])
        assert abs(loss.item() - O.cox_loss(risks, times, c).item()) < 1e-6
        # cohort gradient averaging
        lin = torch.nn.Linear(4, 2)
        torch.manual_seed(rank)
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, float(rank + 1))
        P.sync_cohort_grads(lin)
        assert torch.allclose(lin.weight.grad, torch.full_like(lin.weight, (1 + world) / 2))
        # an instance-sharded bag with an EMPTY shard (fewer 256-row units than ranks): the rank without rows has no
        # gradient for the fc / attention parameters and must still take part in the SUM all-reduce, with zeros
        class Tiny(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.attention_net_WSI = torch.nn.Linear(3, 2)
                self.classifier = torch.nn.Linear(2, 1)
        tiny = Tiny()
        lo2, hi2 = P.shard_rows(200, rank, world)
        assert (hi2 > lo2) == (rank == 0)
        if hi2 > lo2:
            for p_ in tiny.attention_net_WSI.parameters():
                p_.grad = torch.full_like(p_, 3.0)
        tiny.classifier.weight.grad = torch.ones_like(tiny.classifier.weight)   # replicated: not part of the exchange
        P.sync_sharded_bag_grads(tiny)      # would hang (rank 1 skipping the collective) before the fix
        for p_ in tiny.attention_net_WSI.parameters():
            assert p_.grad is not None and torch.allclose(p_.grad, torch.full_like(p_, 3.0))
        assert torch.allclose(tiny.classifier.weight.grad, torch.ones_like(tiny.classifier.weight))
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_bag_and_cohort_protocol_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
