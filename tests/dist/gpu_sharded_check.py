"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it): the two sharding schemes of
SURVEY.md §8(e) on real devices over NCCL.
  1. instance-sharded bag: every rank runs the fused kernels on its row range of ONE bag through the drop-in
     model (bag_group set); pooled M / hazards must equal the single-GPU whole-bag result, and the SUM
     all-reduce of the per-shard weight gradients must equal the whole-bag gradients;
  2. cohort data parallel: bags dealt by size, risks all-gathered, Cox loss identical on every rank.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from multimodalfusion_b200 import parallel as P
from multimodalfusion_b200.models import MIL_Attention_fc_surv_path
from multimodalfusion_b200.utils import CoxSurvLoss, NLLSurvLoss


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-30)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # (N = 200: one 256-row unit for `world` ranks — every rank but 0 owns an EMPTY shard: neutral partial in the
    #  all-gather, zeros in the gradient all-reduce; hung before round 2's sync_sharded_bag_grads fix)
    for size, N in (("small", 5000), ("big", 16384 + 77), ("small", 200)):
        torch.manual_seed(0)
        model = MIL_Attention_fc_surv_path(gate_path=True, model_size_wsi=size, n_classes=4).to(dev).eval()
        g = torch.Generator().manual_seed(3)
        x = (0.5 * torch.randn(N, 1024, generator=g).abs()).to(torch.bfloat16).to(dev)
        Y, c = torch.tensor([1], device=dev), torch.tensor([0.0], device=dev)
        loss_fn = NLLSurvLoss(alpha=0.0)
        # whole bag on every rank (reference for the sharded run), in the arithmetic the sharded path uses: plain bf16
        # operands at every size (a whole bag of <= 4096 instances would otherwise take the split-precision fc)
        from multimodalfusion_b200.autograd import AmilPool
        AmilPool.precise_small_bags = False
        hz, S, _, A = model(path_features=x)
        loss_fn(hazards=hz, S=S, Y=Y, c=c).backward()
        AmilPool.precise_small_bags = True
        ref = {n: p.grad.clone() for n, p in model.named_parameters()}
        model.zero_grad(set_to_none=True)
        # sharded
        lo, hi = P.shard_rows(N, rank, world)
        model.bag_group = dist.group.WORLD
        hz2, S2, _, A2 = model(path_features=x[lo:hi])
        loss_fn(hazards=hz2, S=S2, Y=Y, c=c).backward()
        P.sync_sharded_bag_grads(model)
        model.bag_group = None
        assert rel(hz2, hz) < 1e-5, ("hazards", rel(hz2, hz))
        assert torch.equal(A2.reshape(-1), A.reshape(-1)[lo:hi]), "A_raw of a shard must equal the slice of the whole bag"
        wc_scale = ref["attention_net_WSI.3.attention_c.weight"].abs().max().item()
        for n, p in model.named_parameters():
            if n.endswith("attention_c.bias"):
                # d loss / d bc = sum_i ds_i is exactly 0 in exact arithmetic (softmax shift invariance): both
                # values are rounding residue, compare on the scale of the neighbouring dwc
                assert (p.grad - ref[n]).abs().max().item() < 1e-3 * wc_scale + 1e-7, (size, n)
                continue
            e = rel(p.grad, ref[n])
            assert e < 2e-3, (size, n, e)   # split-K order differs; fp32 atomics
        if rank == 0:
            print(f"sharded bag ok: {size} N={N} world={world}", flush=True)
    # cohort: deal, per-rank risks, gather, Cox
    B = 37
    sizes = [300 + 97 * i for i in range(B)]
    deal = P.deal_cohort(sizes, world)
    torch.manual_seed(1)
    risks_all = torch.randn(B, device=dev)
    times = torch.rand(B, device=dev) * 100
    cens = (torch.rand(B, device=dev) < 0.4).float()
    dist.broadcast(risks_all, 0); dist.broadcast(times, 0); dist.broadcast(cens, 0)
    mine = risks_all[deal[rank]]
    allr = P.gather_risks(mine, [len(d) for d in deal])
    order = [i for d in deal for i in d]
    l1 = CoxSurvLoss()(allr, times[order], cens[order])
    l0 = CoxSurvLoss()(risks_all, times, cens)
    assert abs(l1.item() - l0.item()) < 1e-5
    if rank == 0:
        print("cohort gather + cox ok", flush=True)
    # the library's peer-memory all-reduce kernel == NCCL all-reduce (bitwise: fixed rank order, fp32)
    from multimodalfusion_b200.parallel import PeerAllReduce
    n = 921221
    for use_mc in (True, False):
      ar = PeerAllReduce(n, n_buffers=2, use_multicast=use_mc)
      for it in range(6):
        b = it % 2
        torch.manual_seed(100 * it + rank)
        src = torch.randn(ar.numel, device=dev)
        ar.buffer(b).copy_(src)
        ref = src.clone()
        dist.all_reduce(ref)
        ar.all_reduce(b)
        err = (ar.buffer(b) - ref).abs().max().item()
        assert err <= 1e-5 * ref.abs().max().item(), ("peer all-reduce", use_mc, ar.multicast, it, err)
    # and captured in a CUDA graph, replayed
    g = torch.cuda.CUDAGraph()
    src = torch.randn(ar.numel, device=dev)
    with torch.cuda.graph(g):
        ar.buffer(0).copy_(src)
        ar.all_reduce(0)
    ref = src.clone(); dist.all_reduce(ref)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert (ar.buffer(0) - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    if rank == 0:
        print("peer all-reduce ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
