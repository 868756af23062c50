"""torchrun probe: time the peer-memory all-reduce kernel, NCCL all-reduce, and a fill on symmetric vs ordinary memory."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from multimodalfusion_b200.parallel import PeerAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 921224


def timeit(fn, reps=20, inner=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    return statistics.median(ts)


for ctas, mc in ((8, True), (32, True), (16, False), (32, False)):
    ar = PeerAllReduce(n, n_buffers=1, n_ctas=ctas, use_multicast=mc)
    src = torch.randn(ar.numel, device=dev) + rank
    ar.buffer(0).copy_(src); ref = src.clone(); dist.all_reduce(ref); ar.all_reduce(0)
    err = (ar.buffer(0) - ref).abs().max().item()
    t = timeit(lambda: ar.all_reduce(0))
    if rank == 0:
        print(f"peer all-reduce {n * 4 / 1e6:.2f} MB, {ctas} CTAs, multicast={ar.multicast}: {t:.1f} us (eager), max err vs NCCL {err:.2e}")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            ar.all_reduce(0)
    t = timeit(g.replay, inner=1) / 10
    if rank == 0:
        print(f"   in a graph (10 per replay): {t:.1f} us")
x = torch.zeros(n, device=dev)
t = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    print(f"NCCL all-reduce: {t:.1f} us")
t = timeit(lambda: x.zero_())
t2 = timeit(lambda: ar.buffer(0).zero_())
if rank == 0:
    print(f"fill ordinary {t:.1f} us, symmetric {t2:.1f} us")
dist.barrier(); dist.destroy_process_group()
