"""Per-kernel durations of one radiology patient step (BASELINE config 2) under ncu (serialised, warm caches):
  ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/radio_step_kernels.py
The tensor-core kernels are launched with programmatic dependent launch: a profiler timeline overlaps them, ncu does not."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
from multimodalfusion_b200.utils import get_optim
dev = torch.device("cuda")
torch.manual_seed(0)
N = int(os.environ.get("N", "155"))
model = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
model.enable_fused_step()
opt = get_optim(model, types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5))
bag = {m: (0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for m in model.modalities}
Y, c = torch.tensor([1], device=dev), torch.tensor([0.0], device=dev)
for _ in range(int(os.environ.get("STEPS", "4"))):
    model.fused_step(Y=Y, c=c, alpha=0.0, **bag)
    opt.step(zero_grad=False)
torch.cuda.synchronize()
print("ok")
