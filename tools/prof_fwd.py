"""Minimal driver for ncu: a few launches of the fused AMIL tile kernel (fwd + bwd-gate) on a 16k bag."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200 import ops
L, D, N = int(os.environ.get("L", 512)), int(os.environ.get("D", 384)), int(os.environ.get("N", 16384))
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(3)]
flags = ops.amil_flags(True, dropout_h=True)
dM = torch.randn(L, device=dev) * 0.1
for i in range(4):
    A_raw, M, ml = ops.amil_forward(xs[i % 3], prep, flags, 1)
    ops.amil_backward(xs[i % 3], prep, flags, 1, A_raw, ml, M, dM)
torch.cuda.synchronize()
print("ok", M[:4].tolist())
