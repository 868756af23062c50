"""Minimal driver for ncu: a few eager training steps of the hot path on a 16k bag
(MODE=stash: mmf_amil_fwd_train + stashed backward; MODE=recompute: mmf_amil_fwd + recompute backward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200 import ops
L, D, N = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384))
MODE, STEPS = os.environ.get("MODE", "stash"), int(os.environ.get("STEPS", 4))
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
Wk = torch.randn(4, L, device=dev) * 0.05; bk = torch.zeros(4, device=dev)
Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(3)]
flags = ops.amil_flags(True, dropout_h=True)
ws = ops.amil_bwd_workspace(N, prep, flags, dev)
for i in range(STEPS):
    x = xs[i % 3]
    if MODE == "stash":
        A_raw, parts, st = ops.amil_partials_train(x, prep, flags, 1, workspace=ws)
    else:
        (A_raw, parts), st = ops.amil_partials(x, prep, flags, 1), None
    t = ops.amil_head_nll_step(parts, Wk, bk, Y, c, 0.0)
    g = ops.amil_backward(x, prep, flags, 1, A_raw, t["ml"], t["M"], t["dM"], stash=st)
torch.cuda.synchronize()
print("ok", t["loss"].item(), g["dW1"].abs().max().item())
