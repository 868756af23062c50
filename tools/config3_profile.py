"""Kernel-level breakdown of one BASELINE config-3 cohort step (multimodal Kronecker head, B = 512, Cox loss, fused Adam):
launch count and GPU time per kernel from torch.profiler — is the step launch-bound or kernel-bound?"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from multimodalfusion_b200.models import coxranking_models_pretrained as cox_heads
from multimodalfusion_b200.utils import CoxSurvLoss, get_optim
dev = torch.device("cuda")
torch.manual_seed(0)
args = types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5)
B = 512
head = cox_heads.multimodal_pretrained(mode="radio_path_omic", train_type="kronecker", n_classes=4).to(dev).train()
opt3 = get_optim(head, args)
emb = [torch.randn(B, 256, device=dev) for _ in range(3)]
times = (torch.empty(B, device=dev).exponential_(1 / 30.0).clamp_(0, 250) * 2).round() / 2
cens = (torch.rand(B, device=dev) < 0.46).float()
lf = CoxSurvLoss()


def it():
    risk, _, _ = head(*emb)
    loss = lf(risks=risk, times=times, c=cens)
    loss.backward()
    opt3.step()
    opt3.zero_grad(set_to_none=True)


for _ in range(3):
    it()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        it()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type.name == "CUDA"]
tot = sum(e.device_time_total for e in ev) / 4
print(f"launches per step: {sum(e.count for e in ev) / 4:.0f}, GPU time per step: {tot:.0f} us")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]:
    print(f"  {e.device_time_total / 4:8.1f} us  x{e.count / 4:4.1f}  {e.key[:110]}")
