"""torchrun diagnostic for the cohort-data-parallel step (bench.py --gpus N): where does the per-step gradient
exchange spend its time? Every rank runs the fused 3-launch step on its own bags; the exchange kernel
(p2p_allreduce_sum_kernel) stamps %globaltimer at kernel start / ready handshake done / data phase done / done
handshake done (mmf_debug_set_p2p_stamp_buffer). Modes:
  none     no exchange (lower bound; invalid as a result)
  overlap  exchange on a communication stream, overlapping the next step (bench.py's mode)
  inline   exchange on the step's stream right after the wgrad GEMM
  nccl     dist.all_reduce on the communication stream
Prints, per mode: us/step (max over ranks) and per-rank medians of the exchange's phases.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dp_diag.py
"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
from multimodalfusion_b200.models import MIL_Attention_fc_surv_path
from multimodalfusion_b200.parallel import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = mmf.lib()
N, L, D, K, NB = 16384, 512, 384, 4, 8
STEPS = int(os.environ.get("STEPS", 96))
torch.manual_seed(0)
model = MIL_Attention_fc_surv_path(gate_path=True, model_size_wsi="big", dropout=False, n_classes=K).to(dev).train()
fc, attn = model.attention_net_WSI[0], model.attention_net_WSI[3]
prep = ops.prepare_amil_weights(fc.weight, fc.bias, *attn.amil_weights())
Wk, bk = model.classifier.weight.detach(), model.classifier.bias.detach()
flags = ops.amil_flags(True, dropout_h=True)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
bags = [(0.5 * torch.randn(N, 1024, device=dev, generator=g).abs()).to(torch.bfloat16) for _ in range(NB)]
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
n_flat = (sum(sizes) + 3) // 4 * 4
n_ctas = int(os.environ.get("P2P_CTAS", 0))
ar = PeerAllReduce(n_flat, n_buffers=2, n_ctas=n_ctas)
flats, views, grads = [], [], []
for bi in range(2):
    fl = ar.buffer(bi)
    vs, o = [], 0
    for sz in sizes:
        vs.append(fl[o:o + sz]); o += sz
    flats.append(fl); views.append(vs)
    grads.append(dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5]))
fb = ops.FusedStepBuffers(N, prep, flags, K, dev)
fb.pack_head(Wk)


def step(x, b):
    return ops.amil_fused_step(x, prep, flags, 7 + rank, fb, Wk, bk, Y, c, 0.0, grads[b], dWk=views[b][6].view(K, L),
                               dbk=views[b][7], zero=flats[b], repack_head=False)


for i in range(2):
    step(bags[i], i % 2)
torch.cuda.synchronize()
graphs = []
for i in range(NB):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step(bags[i], i % 2)
    graphs.append(gr)
comm = torch.cuda.Stream()
stamps = torch.zeros(8 + 4 * 4000, dtype=torch.int64, device=dev)


def run(mode, n):
    reduced = [None, None]
    cur = torch.cuda.current_stream()
    for i in range(n):
        b = i % 2
        if reduced[b] is not None:
            cur.wait_event(reduced[b])
        graphs[i % NB].replay()
        if mode == "none":
            continue
        if mode == "inline":
            ar.all_reduce(b)
            continue
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            if mode == "nccl":
                dist.all_reduce(flats[b])
            else:
                ar.all_reduce(b)
            reduced[b] = torch.cuda.Event()
            reduced[b].record(comm)
    for ev in reduced:
        if ev is not None:
            cur.wait_event(ev)


# mode "graph": 8 steps AND their exchanges in ONE graph — the exchange of step i forks onto a second stream and is
# joined before step i + 2 clears the same gradient buffer (no host work per step)
def capture_graph_mode():
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(gr):
        cap = torch.cuda.current_stream()
        done = [None, None]
        for i in range(NB):
            b = i % 2
            if done[b] is not None:
                cap.wait_event(done[b])
            step(bags[i], b)
            ev = torch.cuda.Event(); ev.record(cap)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                ar.all_reduce(b)
                done[b] = torch.cuda.Event(); done[b].record(side)
        for ev in done:
            cap.wait_event(ev)
    return gr


modes = os.environ.get("MODES", "none,overlap,inline,nccl").split(",")
loop_graph = capture_graph_mode() if "graph" in modes else None
_run = run


def run(mode, n):
    if mode != "graph":
        return _run(mode, n)
    for _ in range(n // NB):
        loop_graph.replay()
for mode in modes:
    run(mode, 16)
    torch.cuda.synchronize(); dist.barrier()
    stamps.zero_()
    lib.mmf_debug_set_p2p_stamp_buffer(stamps.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    run(mode, STEPS)
    e1.record()
    torch.cuda.synchronize()
    lib.mmf_debug_set_p2p_stamp_buffer(None)
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / STEPS], device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    st = stamps.cpu()
    k = int(st[0])
    line = f"[{mode}] rank {rank}: {t.item():7.1f} us/step"
    if k > 4:
        rec = st[8:8 + 4 * k].view(k, 4).double()[2:]
        ready_w, data, done_w = (rec[:, 1] - rec[:, 0]) / 1e3, (rec[:, 2] - rec[:, 1]) / 1e3, (rec[:, 3] - rec[:, 2]) / 1e3
        period = (rec[1:, 0] - rec[:-1, 0]) / 1e3
        q = lambda v: f"{v.median().item():6.1f}/{v.quantile(0.9).item():6.1f}/{v.max().item():6.1f}"
        line += (f" | exchange us (median/p90/max): ready-wait {q(ready_w)}  data {q(data)}  done-wait {q(done_w)}"
                 f"  start-to-start {q(period)}  [{k} exchanges]")
    gathered = [None] * world
    dist.all_gather_object(gathered, line)
    if rank == 0:
        print(f"== mode {mode}: {tmax.item():.1f} us/step (max over {world} ranks), multicast={ar.multicast}, ctas={ar.n_ctas}")
        for l in gathered:
            print("   " + l)
        sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
