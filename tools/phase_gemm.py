"""clock64 phase stamps of the pair GEMM kernels (hidden dU GEMM, grouped wgrad): median cycles since CTA start."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
from multimodalfusion_b200._lib import AmilGrads, check
L, D, N = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384))
dev = torch.device("cuda")
torch.manual_seed(0)
W = [torch.randn(L, 1024, device=dev) * 0.03, torch.randn(L, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05,
     torch.randn(D, device=dev) * 0.05, torch.randn(D, L, device=dev) * 0.05, torch.randn(D, device=dev) * 0.05,
     torch.randn(1, D, device=dev) * 0.1, torch.zeros(1, device=dev)]
prep = ops.prepare_amil_weights(*W)
x = (0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16)
flags = ops.amil_flags(True, dropout_h=True)
ws = ops.amil_bwd_workspace(N, prep, flags, dev)
lib = mmf.lib()
KD = 2 * D
grads = dict(dW1=torch.zeros(L, 1024, device=dev), db1=torch.zeros(L, device=dev), dWab=torch.zeros(KD, L, device=dev),
             dbab=torch.zeros(KD, device=dev), dwc=torch.zeros(D, device=dev), dbc=torch.zeros(1, device=dev))
gs = AmilGrads(*[grads[k].data_ptr() for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc")])
wst = prep.struct()
A_raw, parts, _ = ops.amil_partials_train(x, prep, flags, 1, workspace=ws)
M, ml = ops.amil_combine(parts, L, True)
dM = torch.randn(L, device=dev) * 0.1
S = torch.cuda.current_stream().cuda_stream
names = {0: "start", 1: "after cluster sync", 2: "mma: first stage landed", 3: "mma: all issued", 4: "epi: acc ready",
         5: "epi: done", 6: "after final cluster sync"}


names_fused = {0: "start", 1: "after cluster sync + PDL wait", 2: "mma: first A stage ready", 3: "mma: all issued",
               4: "epi: phase A done (ds, mask)", 5: "epi: all slices transformed", 6: "epi: acc ready", 7: "epi: done",
               8: "slice 0: A tile landed", 9: "slice 0: transformed", 10: "slice 1: A tile landed", 11: "slice 1: transformed",
               12: "slice 2: A tile landed", 13: "slice 2: transformed"}


def run(which):
    if which == "fused":
        ops.amil_partials_train(x, prep, flags, 1, workspace=ws)
        torch.cuda.synchronize()
        if buf_holder: buf_holder[0].zero_()
        check(lib.mmf_amil_bwd_gate_hidden_stashed(N, C.byref(wst), L, D, flags, 1, A_raw.data_ptr(), ml.data_ptr(), M.data_ptr(), dM.data_ptr(), None, C.byref(gs), ws.data_ptr(), ws.numel(), S))
    elif which == "hidden":
        check(lib.mmf_amil_bwd_hidden(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, A_raw.data_ptr(), ml.data_ptr(), dM.data_ptr(), C.byref(gs), ws.data_ptr(), ws.numel(), S))
    else:
        check(lib.mmf_amil_bwd_wgrad(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, C.byref(gs), None, ws.data_ptr(), ws.numel(), S))


buf_holder = []
run("fused")   # leaves dG / dU in the workspace for the stand-alone hidden / wgrad stages
for which in ("fused", "hidden", "wgrad"):
    for _ in range(3):
        run(which)
    torch.cuda.synchronize()
    buf = torch.zeros(512, 16, dtype=torch.int64, device=dev)
    buf_holder[:] = [buf]
    lib.mmf_debug_set_timing_buffer(buf.data_ptr())
    run(which)
    torch.cuda.synchronize()
    lib.mmf_debug_set_timing_buffer(None)
    t = buf.cpu().double()
    used = t[:, 0] > 0
    t = t[used]
    rel = t - t[:, :1]
    print(f"== {which}: {int(used.sum())} CTAs; cycles since CTA start (median leader / median peer / max) ==")
    nm = names_fused if which == "fused" else names
    for k in sorted(nm):
        lead = rel[0::2, k][t[0::2, k] > 0]; peer = rel[1::2, k][t[1::2, k] > 0]
        allv = rel[:, k][t[:, k] > 0]
        f = lambda v: f"{v.median().item():9.0f}" if len(v) else "      nan"
        print(f"  {k} {nm[k]:32s} {f(lead)} {f(peer)} {allv.max().item() if len(allv) else float('nan'):9.0f}")
