"""torchrun probe: does the all-reduce get slower when the NVLink links sat idle before it? Graph of
[spin kernel of `gap` us (SMs busy, links idle), all-reduce] x 10; reports (replay time / 10 - gap)."""
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
ar = PeerAllReduce(n, n_buffers=1)
clk_mhz = 1965.0
for gap_us in (0, 10, 30, 120, 500):
    cycles = int(gap_us * clk_mhz)
    g = torch.cuda.CUDAGraph()
    for _ in range(2):
        ar.all_reduce(0)
    torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.graph(g):
        for _ in range(10):
            if cycles:
                torch.cuda._sleep(cycles)
            ar.all_reduce(0)
    # the spin alone, to subtract its true duration
    gs = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gs):
        for _ in range(10):
            if cycles:
                torch.cuda._sleep(cycles)
    def t(gr):
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize(); dist.barrier()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 10)
        return statistics.median(ts)
    both, spin = t(g), (t(gs) if cycles else 0.0)
    if rank == 0:
        print(f"gap {gap_us:4d} us (measured spin {spin:6.1f} us): all-reduce after the gap = {both - spin:6.1f} us  (multicast={ar.multicast})", flush=True)
dist.barrier(); dist.destroy_process_group()
