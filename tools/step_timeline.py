"""Where does a training step's time go BETWEEN kernels? Needs a library built with -DMMF_DEBUG_TIMELINE=1 (passed through
MMF_LIB_PATH; tools/ab_variant.py build timeline -DMMF_DEBUG_TIMELINE=1): every CTA of the three kernels of the fused step
logs %globaltimer at CTA start, after griddepcontrol.wait and at CTA end. Prints, per kernel of the median step of the
8-step CUDA graph: first / median / last CTA start, first / last wait-return, first / median / last CTA end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200 import ops
L, D, N, K = 512, 384, int(os.environ.get("N", 16384)), 4
dev = torch.device("cuda")
torch.manual_seed(0)
W = [torch.randn(L, 1024, device=dev) * 0.03, torch.randn(L, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05,
     torch.randn(D, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05, torch.randn(D, device=dev) * 0.05,
     torch.randn(1, D, device=dev) * 0.1, torch.zeros(1, device=dev)]
Wk, bk = torch.randn(K, L, device=dev) * 0.05, torch.zeros(K, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(*W)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(8)]
flags = ops.amil_flags(True, dropout_h=True)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
flat = torch.zeros((sum(sizes) + 3) // 4 * 4, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
lib = mmf.lib()


def step(i):
    ops.amil_fused_step(xs[i % 8], prep, flags, 1, buf, Wk, bk, Y, c, 0.0, grads, dWk=vs[6].view(K, L), dbk=vs[7], zero=flat)


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
log = torch.zeros(16 + 5 * 32768, dtype=torch.int64, device=dev)
lib.mmf_debug_set_timeline_buffer(log.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); g.replay(); e1.record()
torch.cuda.synchronize()
lib.mmf_debug_set_timeline_buffer(None)
b = log.cpu()
n = int(b[0])
if n == 0:
    raise SystemExit("no records: the library was not built with -DMMF_DEBUG_TIMELINE=1")
rec = b[16:16 + 5 * n].view(n, 5)
print(f"{n} CTA records over two replays of the 8-step graph; event time {e0.elapsed_time(e1) * 1e3 / 16:.2f} us/step")
names = {0: "fwd tile", 2: "head+gate+hidden", 3: "wgrad"}
# split into launches: records of one launch share (id) and are contiguous in time; cluster by start gaps > 20 us per id
launches = []
for kid in names:
    r = rec[rec[:, 0] == kid]
    r = r[r[:, 2].argsort()]
    cur = [r[0]]
    for row in r[1:]:
        if row[2] - cur[0][2] > 60000:   # a new launch of this kernel starts > 60 us after the previous one's first CTA
            launches.append((kid, torch.stack(cur))); cur = []
        cur.append(row)
    launches.append((kid, torch.stack(cur)))
launches.sort(key=lambda kv: int(kv[1][:, 2].min()))
t_first = int(launches[0][1][:, 2].min())
us = lambda t: (float(t) - t_first) / 1e3
print("per launch (us since the first CTA of the first kernel): CTAs | start first/median/last | wait-return first/last | end first/median/last")
prev_end = None
for kid, r in launches[9:27]:   # steps 3-8 of the first replay
    st, wt, en = r[:, 2], r[:, 3], r[:, 4]
    cyc = (r[:, 1] >> 32).double()
    ghz = (cyc / (en - wt).double()).median().item()
    gap = "" if prev_end is None else f"  gap after previous kernel's last CTA end: first start {us(st.min()) - prev_end:+.2f}, last wait-return {us(wt.max()) - prev_end:+.2f}"
    print(f"  {names[kid]:17s} {r.shape[0]:3d} | {us(st.min()):8.2f} {us(st.median()):8.2f} {us(st.max()):8.2f} | {us(wt.min()):8.2f} {us(wt.max()):8.2f} |"
          f" {us(en.min()):8.2f} {us(en.median()):8.2f} {us(en.max()):8.2f} | {cyc.median().item() / 1e3:6.1f}k cycles wait->end, {ghz:.3f} GHz{gap}")
    prev_end = us(en.max())
