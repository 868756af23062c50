"""In-kernel phase breakdown of the fused tile kernel (clock64 stamps, see MMF_STAMP in csrc/amil_tile2.cuh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
L, D, N = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384))
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(6)]
flags = ops.amil_flags(True, dropout_h=True)
dM = torch.randn(L, device=dev) * 0.1
names = {0: "start", 1: "after cluster sync", 2: "producer: GEMM1 loads issued", 3: "producer: acc1 seen",
         5: "mma: first stage landed", 6: "mma: GEMM1 issued", 7: "mma: H ready", 8: "mma: GEMM2 issued",
         9: "epi: vectors staged", 10: "epi: acc1 seen", 11: "epi: EPI1 done", 12: "epi: EPI2 done", 13: "epi: tail done",
         14: "after final cluster sync"}
for mode in ("fwd", "bwd"):
    for i in range(3):
        A_raw, M, ml = ops.amil_forward(xs[i], prep, flags, 1)
    g = ops.amil_backward(xs[0], prep, flags, 1, A_raw, ml, M, dM)
    torch.cuda.synchronize()
    tiles = 2 * ((N + 255) // 256)
    buf = torch.zeros(tiles, 16, dtype=torch.int64, device=dev)
    mmf.lib().mmf_debug_set_timing_buffer(buf.data_ptr())
    if mode == "fwd":
        ops.amil_partials(xs[3], prep, flags, 1)
    else:
        A_raw, M, ml = ops.amil_forward(xs[4], prep, flags, 1)
        buf.zero_()
        ops.amil_backward(xs[4], prep, flags, 1, A_raw, ml, M, dM)
    torch.cuda.synchronize()
    mmf.lib().mmf_debug_set_timing_buffer(None)
    t = buf.cpu().double()
    rel = t - t[:, :1]
    print(f"== {mode}: cycles since CTA start, median over {tiles} CTAs (leader rows / peer rows) ==")
    for k in sorted(names):
        col = rel[:, k]
        lead, peer = col[0::2], col[1::2]
        lv = lead[t[0::2, k] > 0]; pv = peer[t[1::2, k] > 0]
        print(f"  {k:2d} {names[k]:32s} leader {lv.median().item() if len(lv) else float('nan'):10.0f}   peer {pv.median().item() if len(pv) else float('nan'):10.0f}")
