"""Minimal driver for ncu / compute-sanitizer: a few eager training steps of the hot path on one bag, through the
fused 3-launch step (mmf_amil_fwd_train_head -> mmf_amil_bwd_head = head + gate + hidden kernel, grouped wgrad).
env: N (16384), L / D (512 / 384), STEPS (4), K (4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200 import ops
L, D, N, K = (int(os.environ.get(k, d)) for k, d in (("L", 512), ("D", 384), ("N", 16384), ("K", 4)))
STEPS = int(os.environ.get("STEPS", 4))
dev = torch.device("cuda")
torch.manual_seed(0)
W1 = torch.randn(L, 1024, device=dev) * 0.03; b1 = torch.randn(L, device=dev) * 0.05
Wa = torch.randn(D, L, device=dev) * 0.05; ba = torch.randn(D, device=dev) * 0.05
Wb = torch.randn(D, L, device=dev) * 0.05; bb = torch.randn(D, device=dev) * 0.05
wc = torch.randn(1, D, device=dev) * 0.1; bc = torch.zeros(1, device=dev)
Wk = torch.randn(K, L, device=dev) * 0.05; bk = torch.zeros(K, device=dev)
Y, c = torch.tensor([min(2, K - 1)], device=dev), torch.tensor([0.0], device=dev)
prep = ops.prepare_amil_weights(W1, b1, Wa, ba, Wb, bb, wc, bc)
xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(3)]
flags = ops.amil_flags(True, dropout_h=True)
KD = 2 * D
sizes = [L * 1024, L, KD * L, KD, D, 1, K * L, K]
flat = torch.zeros((sum(sizes) + 3) // 4 * 4, device=dev)
vs, o = [], 0
for sz in sizes:
    vs.append(flat[o:o + sz]); o += sz
grads = dict(dW1=vs[0].view(L, 1024), db1=vs[1], dWab=vs[2].view(KD, L), dbab=vs[3], dwc=vs[4], dbc=vs[5])
buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
for i in range(STEPS):
    loss = ops.amil_fused_step(xs[i % 3], prep, flags, 1, buf, Wk, bk, Y, c, 0.0, grads, dWk=vs[6].view(K, L), dbk=vs[7],
                               zero=flat)
torch.cuda.synchronize()
print("ok", loss.item(), grads["dW1"].abs().max().item())
