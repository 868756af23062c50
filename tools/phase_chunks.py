"""Chunk-level timeline of the forward tile kernel's GEMM2 phase in its four forms (needs a library built with
-DMMF_TILE2_CHUNK_STAMPS=1, passed through MMF_LIB_PATH): inference, + H stash, training (H + [a|g] stash), training + z / mask."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
L, D, N, K = 512, 384, int(os.environ.get("N", 16384)), 4
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
Wk = torch.randn(K, L, device=dev) * 0.05; bk = torch.zeros(K, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(4)]
flags = ops.amil_flags(True, dropout_h=True)
ws = ops.amil_bwd_workspace(N, prep, flags, dev)
hbuf = torch.empty(N, L, dtype=torch.bfloat16, device=dev)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
flat = torch.zeros((sum(sizes) + 3) // 4 * 4, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
flags0 = ops.amil_flags(True)
forms = {
    "inference, no dropout": lambda x: ops.amil_partials(x, prep, flags0, 1),
    "inference": lambda x: ops.amil_partials(x, prep, flags, 1),
    "+ H stash": lambda x: ops.amil_partials(x, prep, flags, 1, h_stash=hbuf),
    "training (H + a|g stash)": lambda x: ops.amil_partials_train(x, prep, flags, 1, workspace=ws),
}
names = {0: "start", 1: "after cluster sync", 6: "mma: GEMM1 issued", 10: "epi: acc1 seen", 11: "epi: EPI1 done", 7: "mma: H ready",
         2: "epi: chunk 0 accumulators seen", 5: "epi: chunk 0 done", 3: "epi: chunk 1 accumulators seen", 9: "epi: chunk 1 done",
         4: "epi: chunk 2 accumulators seen", 15: "epi: chunk 2 done", 8: "mma: GEMM2 (+z) issued", 12: "epi: EPI2 (+ z) done",
         13: "epi: tail done", 14: "after final cluster sync"}
order = [0, 1, 6, 10, 11, 7, 2, 5, 3, 9, 4, 15, 8, 12, 13, 14]
tiles = 2 * ((N + 255) // 256)
for title, fn in forms.items():
    for i in range(3):
        fn(xs[i])
    torch.cuda.synchronize()
    tb = torch.zeros(tiles, 16, dtype=torch.int64, device=dev)
    mmf.lib().mmf_debug_set_timing_buffer(tb.data_ptr())
    fn(xs[3])
    torch.cuda.synchronize()
    mmf.lib().mmf_debug_set_timing_buffer(None)
    t = tb.cpu().double()
    rel = t - t[:, :1]
    print(f"== {title}: cycles since CTA start, median over {tiles} CTAs (max) ==")
    for k in order:
        col = rel[:, k]
        col = col[t[:, k] > 0] if k else col
        if col.numel():
            print(f"  {k:2d} {names[k]:36s} {col.median().item():8.0f}  ({col.max().item():.0f})")
