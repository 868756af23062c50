"""Gradient-accumulation window of small bags: gc batch-1 fused steps vs ONE varlen-packed window step
(MIL_Attention_fc_surv_path.fused_window_step), fwd + nll_surv + bwd, gradients accumulated in .grad. Wall clock incl. host."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200.models import MIL_Attention_fc_surv_path
dev = torch.device("cuda")
torch.manual_seed(0)
for preset, lo, hi, gc in (("small", 80, 156, 32), ("small", 500, 3000, 32), ("big", 2000, 12000, 16)):
    model = MIL_Attention_fc_surv_path(gate_path=True, model_size_wsi=preset, dropout=True, n_classes=4).to(dev).train()
    model.enable_fused_step()
    g = torch.Generator().manual_seed(1)
    ns = torch.randint(lo, hi, (gc,), generator=g).tolist()
    bags = [(0.5 * torch.randn(n, 1024, device=dev).abs()).to(torch.bfloat16) for n in ns]
    Y = torch.randint(0, 4, (gc,), device=dev)
    c = (torch.rand(gc, device=dev) < 0.4).float()

    def loop():
        for i, b in enumerate(bags):
            model.fused_step(path_features=b, Y=Y[i:i + 1], c=c[i:i + 1], alpha=0.0, loss_scale=1.0 / gc, accumulate=i > 0)

    def window():
        model.fused_window_step(bags, Y, c, alpha=0.0)

    res = {}
    for name, fn, reps in (("loop", loop, 10), ("window", window, 20)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(); res[name] = (time.perf_counter() - t0) / reps
    print(json.dumps({"preset": preset, "gc": gc, "rows": sum(ns), "bag_rows": f"{lo}..{hi - 1}",
                      "ms_per_window_loop_of_fused_steps": res["loop"] * 1e3, "ms_per_window_packed": res["window"] * 1e3,
                      "speedup": res["loop"] / res["window"], "us_per_bag_packed": res["window"] / gc * 1e6}))

# radiology (BASELINE config 2 shapes) with a gradient-accumulation window of 32 patients
from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
gc = 32
model = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
model.enable_fused_step()
g = torch.Generator().manual_seed(2)
ns = torch.randint(80, 156, (gc,), generator=g).tolist()
patients = [{m: (0.5 * torch.randn(n, 1024, device=dev).abs()).to(torch.bfloat16) for m in model.modalities} for n in ns]
Y = torch.randint(0, 4, (gc,), device=dev)
c = (torch.rand(gc, device=dev) < 0.4).float()


def loop_r():
    for i, p_ in enumerate(patients):
        model.fused_step(Y=Y[i:i + 1], c=c[i:i + 1], alpha=0.0, loss_scale=1.0 / gc, accumulate=i > 0, **p_)


def window_r():
    model.fused_window_step(patients, Y, c, alpha=0.0)


res = {}
for name, fn, reps in (("loop", loop_r, 10), ("window", window_r, 20)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); res[name] = (time.perf_counter() - t0) / reps
print(json.dumps({"model": "radio_attention_mil, 4 modalities", "gc": gc, "slices": sum(ns), "ms_per_window_loop_of_fused_steps": res["loop"] * 1e3,
                  "ms_per_window_packed": res["window"] * 1e3, "speedup": res["loop"] / res["window"],
                  "us_per_patient_packed": res["window"] / gc * 1e6}))
