"""Per-stage device times of the hot path (each stage captured 8x in a CUDA graph over rotating bags, so host
launch overhead is excluded), plus the raw pinned-host -> device copy rate that bounds the e2e number."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200 import ops
from multimodalfusion_b200._lib import AmilGrads, check
L, D, N = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384))
NB = 8
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
Wk = torch.randn(4, L, device=dev) * 0.05; bk = torch.zeros(4, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(NB)]
flags = ops.amil_flags(True, dropout_h=True)
ws = ops.amil_bwd_workspace(N, prep, flags, dev)
lib = mmf.lib()
hbuf = torch.empty(N, L, dtype=torch.bfloat16, device=dev)
KD = 2 * D
grads = dict(dW1=torch.zeros(L, 1024, device=dev), db1=torch.zeros(L, device=dev), dWab=torch.zeros(KD, L, device=dev),
             dbab=torch.zeros(KD, device=dev), dwc=torch.zeros(D, device=dev), dbc=torch.zeros(1, device=dev))
gs = AmilGrads(*[grads[k].data_ptr() for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc")])
wst = prep.struct()
A_raw, parts, _ = ops.amil_partials_train(xs[0], prep, flags, 1, workspace=ws)
t = ops.amil_head_nll_step(parts, Wk, bk, Y, c, 0.0)
M, ml, dM = t["M"], t["ml"], t["dM"]
dWk, dbk = torch.zeros(4, L, device=dev), torch.zeros(4, device=dev)


def S():
    return torch.cuda.current_stream().cuda_stream


fbuf = ops.FusedStepBuffers(N, prep, flags, 4, dev)
fbuf.pack_head(Wk)
flat = torch.zeros(4096, device=dev)
head = fbuf.head_struct(Wk, bk, Y, c, 0.0, 1e-7, 1.0, dWk, dbk)
fbuf.workspace = ws   # share the stash
ops.amil_fused_step(xs[0], prep, flags, 1, fbuf, Wk, bk, Y, c, 0.0, grads, dWk=dWk, dbk=dbk)


def fwd_train_head(x, h=head):
    check(lib.mmf_amil_fwd_train_head(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, 1, fbuf.A_raw.data_ptr(),
                                      fbuf.partials.data_ptr(), ws.data_ptr(), ws.numel(), flat.data_ptr(), flat.numel(),
                                      C.byref(h), S()))


def hidden_headproj(x):
    from multimodalfusion_b200._lib import MMF_STASHED
    # (gate+hidden with the head-projected phase A, then wgrad: subtract the wgrad stage)
    check(lib.mmf_amil_bwd_head(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags | MMF_STASHED, 1, fbuf.A_raw.data_ptr(),
                                fbuf.partials.data_ptr(), C.byref(head), None, C.byref(gs), None, ws.data_ptr(), ws.numel(), S()))


def hidden_only(x):
    from multimodalfusion_b200._lib import MMF_STASHED
    check(lib.mmf_amil_bwd_gate_hidden_head(N, C.byref(wst), L, D, flags | MMF_STASHED, 1, fbuf.A_raw.data_ptr(),
                                            fbuf.partials.data_ptr(), C.byref(head), None, C.byref(gs), ws.data_ptr(), ws.numel(), S()))


stages = {
    "fwd_train_head": fwd_train_head,
    "gate_hidden_headproj": hidden_only,
    "bwd_head(hid+wgrad)": hidden_headproj,
    "fwd": lambda x: ops.amil_partials(x, prep, flags, 1),
    "fwd_hstash": lambda x: ops.amil_partials(x, prep, flags, 1, h_stash=hbuf),
    "fwd_train": lambda x: ops.amil_partials_train(x, prep, flags, 1, workspace=ws),
    "head_step": lambda x: ops.amil_head_nll_step(parts, Wk, bk, Y, c, 0.0, dWk=dWk, dbk=dbk),
    "gate_hidden_fused": lambda x: check(lib.mmf_amil_bwd_gate_hidden_stashed(N, C.byref(wst), L, D, flags, 1, A_raw.data_ptr(), ml.data_ptr(), M.data_ptr(), dM.data_ptr(), None, C.byref(gs), ws.data_ptr(), ws.numel(), S())),
    "gate_recompute": lambda x: check(lib.mmf_amil_bwd_gate(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, 1, A_raw.data_ptr(), ml.data_ptr(), M.data_ptr(), dM.data_ptr(), None, C.byref(gs), ws.data_ptr(), ws.numel(), S())),
    "hidden": lambda x: check(lib.mmf_amil_bwd_hidden(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, A_raw.data_ptr(), ml.data_ptr(), dM.data_ptr(), C.byref(gs), ws.data_ptr(), ws.numel(), S())),
    "wgrad": lambda x: check(lib.mmf_amil_bwd_wgrad(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, C.byref(gs), None, ws.data_ptr(), ws.numel(), S())),
    "zero_grads": lambda x: [g.zero_() for g in grads.values()],
}
only = os.environ.get("STAGES")
res = {}
for name, fn in stages.items():
    if only and name not in only.split(","):
        continue
    for i in range(2):
        fn(xs[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NB):
            fn(xs[i])
    for _ in range(3):
        g.replay()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / NB)
    res[name] = statistics.median(ts)
    print(f"{name:16s} {res[name]:8.2f} us")
if not only:
    hx = [x.cpu().pin_memory() for x in xs[:4]]
    dst = torch.empty_like(xs[0])
    for _ in range(3):
        dst.copy_(hx[0], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        dst.copy_(hx[i % 4], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(f"H2D pinned {N * 2048 / 2**20:.0f} MiB: {ms * 1e3:.0f} us = {N * 2048 / ms / 1e6:.1f} GB/s -> e2e ceiling {N / ms / 1e3:.1f} M patches/s")
