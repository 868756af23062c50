import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from multimodalfusion_b200 import ops
from oracle import cases
import test_gpu_fused_step as T
dev = torch.device("cuda")
N, L, D, gated, drop, K, y, c_val = 300, 512, 384, True, 2, 4, 2, 0.0
seed = 0xF00D + N
W, Wk, bk = T._rand(L, D, gated, K, N + L + K)
x = cases.features(N, 700 + N)
xb = x.to(dev).to(torch.bfloat16)
prep = ops.prepare_amil_weights(*[None if t is None else t.to(dev) for t in W])
Wkd, bkd = Wk.to(dev), bk.to(dev)
flags = ops.amil_flags(gated) | drop
Y, c = torch.tensor([y], device=dev), torch.tensor([c_val], device=dev)
flat, grads, dWk, dbk = T._grad_bufs(L, D, gated, K, dev)
buf = ops.FusedStepBuffers(N, prep, flags, K, dev)
ops.amil_fused_step(xb, prep, flags, seed, buf, Wkd, bkd, Y, c, 0.15, grads, dWk=dWk, dbk=dbk, loss_scale=0.5, zero=flat)
torch.cuda.synchronize()
print("fused ml", buf.ml.tolist(), "M[:4]", buf.M[:4].tolist(), "loss", buf.loss.item())
print("fused partials[:, :4]", buf.partials[:, :4].tolist())
for rep in range(6):
    flat2, grads2, dWk2, dbk2 = T._grad_bufs(L, D, gated, K, dev)
    junk = torch.full((3, L + 2), 7.0 + rep, device=dev); jp = junk.data_ptr(); del junk
    j2 = torch.full((N,), -3.0, device=dev); del j2
    A_raw, parts, ws = ops.amil_partials_train(xb, prep, flags, seed, zero=flat2)
    t = ops.amil_head_nll_step(parts, Wkd, bkd, Y, c, 0.15, dWk=dWk2, dbk=dbk2)
    ops.amil_backward(xb, prep, flags, seed, A_raw, t["ml"], t["M"], t["dM"] * 0.5, grads=grads2, stash=ws)
    torch.cuda.synchronize()
    print(rep, "reused junk block:", parts.data_ptr() == jp, "modular ml", t["ml"].tolist(), "M[:2]", t["M"][:2].tolist(), "loss", t["loss"].item())
