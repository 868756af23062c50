"""clock64 phase stamps (MMF_STAMP) of the fused training step's kernels: the forward with the folded head
(amil_tile2_kernel, stamps 0-14) and the head-projected gate + hidden backward (amil_hidden_fused_kernel, stamps 0-13).
Prints median and max over CTAs (the folded head runs in ONE CTA: its cost shows in the max)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import multimodalfusion_b200 as mmf
from multimodalfusion_b200._lib import require_debug_stamps; require_debug_stamps()   # needs a -DMMF_DEBUG_STAMPS=1 build (MMF_LIB_PATH)
from multimodalfusion_b200 import ops
from multimodalfusion_b200._lib import AmilGrads, MMF_STASHED, check
L, D, N, K = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384)), 4
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
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
flat = torch.zeros((sum(sizes) + 3) // 4 * 4, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
gs = AmilGrads(*[grads[k].data_ptr() for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc")])
buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
buf.pack_head(Wk)
head = buf.head_struct(Wk, bk, Y, c, 0.0, 1e-7, 1.0, vs[6], vs[7])
wst = prep.struct()
lib = mmf.lib()
S = lambda: torch.cuda.current_stream().cuda_stream


def fwd(x, with_head=True):
    if with_head:
        check(lib.mmf_amil_fwd_train_head(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, 1, buf.A_raw.data_ptr(),
                                          buf.partials.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(),
                                          flat.data_ptr(), flat.numel(), C.byref(head), S()))
    else:
        check(lib.mmf_amil_fwd_train(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags, 1, buf.A_raw.data_ptr(),
                                     buf.partials.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(),
                                     flat.data_ptr(), flat.numel(), S()))


def hid(x):
    check(lib.mmf_amil_bwd_gate_hidden_head(N, C.byref(wst), L, D, flags | MMF_STASHED, 1, buf.A_raw.data_ptr(),
                                            buf.partials.data_ptr(), C.byref(head), None, C.byref(gs), buf.workspace.data_ptr(),
                                            buf.workspace.numel(), S()))


def bwd(x):
    check(lib.mmf_amil_bwd_head(x.data_ptr(), N, 1024, C.byref(wst), L, D, flags | MMF_STASHED, 1, buf.A_raw.data_ptr(),
                                buf.partials.data_ptr(), C.byref(head), None, C.byref(gs), None, buf.workspace.data_ptr(), buf.workspace.numel(), S()))


fwd_names = {0: "start", 1: "after cluster sync + griddep wait", 2: "producer: GEMM1 loads issued", 3: "producer: acc1 seen",
             5: "mma: first stage landed", 6: "mma: GEMM1 issued", 7: "mma: H ready", 8: "mma: GEMM2 (+z) issued",
             9: "epi: vectors staged", 10: "epi: acc1 seen", 11: "epi: EPI1 done", 12: "epi: EPI2 + z done",
             13: "epi: tail (+ folded head) done", 14: "after final cluster sync"}
if os.environ.get("HEAD_STAMPS") == "1":
    hid_names = None
hid_names_head = {0: "start", 1: "after cluster sync + griddep wait", 2: "mma: first A stage ready", 3: "mma: all issued",
                  4: "workers: phase A done", 5: "workers: mainloop done", 6: "workers: acc seen", 7: "workers: epilogue done",
                  8: "head: inputs requested", 9: "head: global max known", 10: "head: sums reduced",
                  13: "head: dlogits / dM done"}
hid_names = {0: "start", 1: "after cluster sync + griddep wait", 2: "mma: first A stage ready", 3: "mma: all issued",
             4: "workers: phase A done", 5: "workers: mainloop done", 6: "workers: acc seen", 7: "workers: epilogue done",
             8: "slice 0 landed", 9: "slice 0 transformed", 10: "slice 1 landed", 11: "slice 1 transformed",
             12: "slice 2 landed", 13: "slice 2 transformed"}


def report(title, names, t):
    rel = t - t[:, :1]
    print(f"== {title}: cycles since CTA start over {t.shape[0]} CTAs ==")
    for k in sorted(names):
        col = rel[:, k][t[:, k] > 0]
        if len(col):
            print(f"  {k:2d} {names[k]:36s} median {col.median().item():9.0f}   max {col.max().item():9.0f}")


tiles = 2 * ((N + 255) // 256)
for i in range(3):
    fwd(xs[i]); bwd(xs[i])
torch.cuda.synchronize()
for title, with_head in (("forward, training form, NO head / z", False), ("forward + z side MMA + mask words", True)):
    dbg = torch.zeros(tiles + 1, 16, dtype=torch.int64, device=dev)
    lib.mmf_debug_set_timing_buffer(dbg.data_ptr())
    fwd(xs[3], with_head)
    torch.cuda.synchronize()
    lib.mmf_debug_set_timing_buffer(None)
    report(title, fwd_names, dbg.cpu().double()[:tiles])
fwd(xs[3])
dbg = torch.zeros(tiles, 16, dtype=torch.int64, device=dev)
lib.mmf_debug_set_timing_buffer(dbg.data_ptr())
hid(xs[3])
torch.cuda.synchronize()
lib.mmf_debug_set_timing_buffer(None)
report("gate + hidden backward (head-projected)", hid_names_head if os.environ.get("HEAD_STAMPS") == "1" else hid_names, dbg.cpu().double())
