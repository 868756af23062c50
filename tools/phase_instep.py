"""clock64 phase stamps of ONE kernel of the fused training step while the whole step runs back to back in a CUDA graph
(8 steps per graph, rotating bags) — the in-step timeline differs from a lone launch (L2 state, write-backs, PDL overlap).
usage: MMF_STAMP_KERNEL=0|2|3 python tools/phase_instep.py   (0 forward tile, 2 head + gate + hidden, 3 grouped wgrad)
With a library built -DMMF_TILE2_CHUNK_STAMPS=1 / -DMMF_HEAD_STAMPS=1 (MMF_LIB_PATH) the forward / hidden stamps change meaning."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
kid = int(os.environ.get("MMF_STAMP_KERNEL", "0"))
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
grid = 148 * 2
tb = torch.zeros(grid, 16, dtype=torch.int64, device=dev)


def step(i):
    ops.amil_fused_step(xs[i % 8], prep, flags, 1, buf, Wk, bk, Y, c, 0.0, grads, dWk=vs[6].view(K, L), dbk=vs[7], zero=flat)


for i in range(2):
    step(i)
torch.cuda.synchronize()
lib.mmf_debug_set_timing_buffer(tb.data_ptr())   # baked into the captured launches of the selected kernel
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(8):
        step(i)
lib.mmf_debug_set_timing_buffer(None)
for _ in range(4):
    g.replay()
torch.cuda.synchronize()
t = tb.cpu().double()
t = t[t[:, 0] > 0]
rel = t - t[:, :1]
print(f"== kernel id {kid} inside the 8-step graph (stamps of the last step): cycles since CTA start, median over {t.shape[0]} CTAs (max) ==")
for k in range(16):
    col = rel[:, k][t[:, k] > 0]
    if col.numel():
        print(f"  {k:2d}  {col.median().item():9.0f}  ({col.max().item():.0f})")
