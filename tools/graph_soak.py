import sys, types, torch
sys.path.insert(0, ".")
from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
from multimodalfusion_b200.utils import get_optim
dev = torch.device("cuda")
torch.manual_seed(0)
model = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
opt = get_optim(model, types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5))
g = torch.Generator().manual_seed(1)
ns = torch.randint(17, 156, (48,), generator=g).tolist()
pats = [{m: (0.5 * torch.randn(n, 1024, device=dev).abs()).to(torch.bfloat16) for m in model.modalities} for n in ns]
Ys = torch.randint(0, 4, (48,), generator=g)
losses, mem = [], []
for epoch in range(8):
    tot = 0.0
    for i, p in enumerate(pats):
        out = model.graphed_fused_step(opt, Y=Ys[i:i + 1].to(dev), c=torch.tensor([0.0], device=dev), alpha=0.0, **p)
        tot += out[4].item()
    losses.append(tot / len(pats)); mem.append(torch.cuda.memory_allocated() >> 20)
print("mean loss per epoch:", [round(l, 4) for l in losses])
print("allocated MiB per epoch:", mem, "graphs:", len(model._graph_family.graphs), "host step:", opt.host_step())
assert all(l == l for l in losses) and losses[-1] < losses[0], "loss must be finite and go down on a fixed cohort"
assert mem[-1] <= mem[1] + 8, "memory must not grow once every size has its graph"
print("soak ok")
