"""Where does a training step's time go BETWEEN kernels? %globaltimer of first-CTA-start / last-CTA-end per kernel,
for the CUDA-graph step of bench.py (8 steps in one graph); prints the median step timeline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200 import ops
L, D, N = 512, 384, 16384
dev = torch.device("cuda")
torch.manual_seed(0)
W = [torch.randn(L, 1024, device=dev) * 0.03, torch.randn(L, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05,
     torch.randn(D, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05, torch.randn(D, device=dev) * 0.05,
     torch.randn(1, D, device=dev) * 0.1, torch.zeros(1, device=dev)]
Wk, bk = torch.randn(4, L, device=dev) * 0.05, torch.zeros(4, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(*W)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(8)]
flags = ops.amil_flags(True, dropout_h=True)
ws = ops.amil_bwd_workspace(N, prep, flags, dev)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, 4 * L, 4]
flat = torch.zeros((sum(sizes) + 3) // 4 * 4, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
tl = torch.zeros(8, 16, dtype=torch.int64, device=dev)
lib = mmf.lib()


def step(i):
    lib.mmf_debug_set_timeline_buffer(tl[i].data_ptr()) if False else None
    A_raw, parts, st = ops.amil_partials_train(xs[i], prep, flags, 1, workspace=ws, zero=flat)
    t = ops.amil_head_nll_step(parts, Wk, bk, Y, c, 0.0, dWk=vs[6], dbk=vs[7])
    ops.amil_backward(xs[i], prep, flags, 1, A_raw, t["ml"], t["M"], t["dM"], grads=grads, stash=st)


for i in range(2):
    step(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(8):
        step(i)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
names = ["fwd", "head", "gate+hidden", "wgrad", "gate(recompute)", "gemm2"]
buf = torch.zeros(16 + 2 * 8192, dtype=torch.int64, device=dev)
lib.mmf_debug_set_timeline_buffer(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); g.replay(); e1.record()
torch.cuda.synchronize()
lib.mmf_debug_set_timeline_buffer(None)
b = buf.cpu().tolist()
ns, ne = b[0], b[1]
starts = [(b[16 + 2 * i], b[17 + 2 * i]) for i in range(ns)]
ends = [(b[16 + 8192 + 2 * i], b[17 + 8192 + 2 * i]) for i in range(ne)]
t0 = starts[0][1]
print(f"{ns} kernel starts, {ne} ends over two replays of the 8-step graph; event time {e0.elapsed_time(e1) * 1e3 / 16:.1f} us/step")
print("CTA 0 of each kernel, us since the first kernel's start (steps 4-6 of the first replay):")
ev = sorted([(t, "start", names[k]) for k, t in starts] + [(t, "end", names[k]) for k, t in ends])
prev = None
for t, what, nm in ev:
    us = (t - t0) / 1e3
    if 4 * 116 <= us <= 7 * 118:
        print(f"  {us:9.2f}  {what:5s} {nm}")
