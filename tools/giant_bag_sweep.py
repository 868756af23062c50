"""BASELINE.json config 4: giant-bag sweep 1k..256k x 1024 (big preset), fwd+bwd, one bag instance-sharded over the
ranks (python tools/giant_bag_sweep.py, or under torchrun for N > 1). Per size: patches/s of the whole bag, time =
max over ranks, CUDA-graph timed. Sharded mode per step: tile kernel on the local rows -> all-gather of the (L+2)
partial (NCCL) -> combine + head on every rank -> local backward -> SUM all-reduce of the fc/attention grads (own
peer-memory kernel when available)."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from multimodalfusion_b200 import ops
from multimodalfusion_b200 import parallel as P

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L, D, K = 512, 384, 4
torch.manual_seed(0)
W = [torch.randn(L, 1024, device=dev) * 0.03, torch.randn(L, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05,
     torch.randn(D, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05, torch.randn(D, device=dev) * 0.05,
     torch.randn(1, D, device=dev) * 0.1, torch.zeros(1, device=dev)]
Wk, bk = torch.randn(K, L, device=dev) * 0.05, torch.zeros(K, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(*W)
flags = ops.amil_flags(True, dropout_h=True)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
tot = (sum(sizes) + 3) // 4 * 4
ar = None
if world > 1:
    try:
        ar = P.PeerAllReduce(tot, n_buffers=1)
        flat = ar.buffer(0)
    except Exception as e:
        flat = torch.zeros(tot, device=dev)
else:
    flat = torch.zeros(tot, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
out = []
for N in [1024 << i for i in range(9)]:
    lo, hi = P.shard_rows(N, rank, world)
    n_loc = hi - lo
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    xs = [(0.5 * torch.randn(max(n_loc, 1), 1024, device=dev, generator=g).abs()).to(torch.bfloat16) for _ in range(2 if N > 65536 else 4)]
    ws = ops.amil_bwd_workspace(max(n_loc, 1), prep, flags, dev)

    def step(x):
        if n_loc > 0:
            A_raw, parts, st = ops.amil_partials_train(x, prep, flags, 1, workspace=ws, zero=flat)
            part = ops.amil_combine(parts, L, False)
        else:
            flat.zero_(); part = P.empty_partial(L, dev)
        if world > 1:
            allp = torch.empty(world, L + 2, device=dev)
            dist.all_gather_into_tensor(allp, part.reshape(1, -1))
        else:
            allp = part.reshape(1, -1)
        t = ops.amil_head_nll_step(allp, Wk, bk, Y, c, 0.0, dWk=vs[6], dbk=vs[7])
        if n_loc > 0:
            ops.amil_backward(x, prep, flags, 1, A_raw, t["ml"], t["M"], t["dM"], grads=grads, stash=st)
        if world > 1:
            if ar is not None:
                ar.all_reduce(0)
            else:
                dist.all_reduce(flat)
        return t["loss"]

    for i in range(3):
        step(xs[i % len(xs)])
    torch.cuda.synchronize()
    reps = 8
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            step(xs[i % len(xs)])
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ms = statistics.median(ts)
    if world > 1:
        tt = torch.tensor([ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = tt.item()
    if rank == 0:
        rec = {"config": "giant-bag sweep (BASELINE config 4)", "N": N, "n_gpus": world, "ms_per_bag": ms,
               "patches_per_s": N / (ms * 1e-3), "timed": "eager launches, CUDA events, median of 5 x 8 steps, max over ranks",
               "allreduce": "peer-memory kernel" if ar is not None else ("nccl" if world > 1 else None)}
        print(json.dumps(rec), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
