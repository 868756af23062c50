"""BASELINE.json config 5 at cohort scale: how many PAIRS of slides does the bf16 tensor-core path order differently from
the reference's fp32 arithmetic? 10 000 synthetic slides (N ~ logN(median 8k) clipped to [500, 64k], ~90 M patches) run
(a) through MIL_Attention_fc_surv_path.infer_cohort (varlen-packed fused forward, plain bf16 operands) and (b) through the
reference's own op sequence — nn.Linear / ReLU / tanh / sigmoid / softmax / mm / sigmoid / cumprod in torch fp32 on the GPU
(models/model_attention_mil_path.py:50-72, TF32 off) on the SAME weights and features. Reports max |risk difference|, the
number of discordant pairs (Kendall distance) overall and among pairs the fp32 run separates by more than 1e-3 / 1e-4, and
the c-index difference against synthetic survival times. A REPORT, not a test: the CPU oracle cannot run 90 M patches."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200.models import MIL_Attention_fc_surv_path
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
SLIDES, PRESET, BATCH = int(os.environ.get("SLIDES", 10000)), os.environ.get("PRESET", "small"), int(os.environ.get("BATCH", 64))
torch.manual_seed(0)
model = MIL_Attention_fc_surv_path(gate_path=True, model_size_wsi=PRESET, n_classes=4).to(dev).eval()
with torch.no_grad():      # biases away from the zero init so that every term of the path matters
    for p in model.parameters():
        if p.dim() == 1:
            p.add_(0.05 * torch.randn_like(p))
g = torch.Generator().manual_seed(1)
sizes = torch.exp(torch.randn(SLIDES, generator=g) * 0.8 + math.log(8000)).clamp(500, 64000).long().tolist()
pool_rows = 1 << 20
gd = torch.Generator(device=dev).manual_seed(3)
pool = torch.empty(pool_rows, 1024, dtype=torch.bfloat16, device=dev)
for r0 in range(0, pool_rows, 1 << 17):
    pool[r0:r0 + (1 << 17)] = (0.5 * torch.randn(1 << 17, 1024, device=dev, generator=gd).abs()).to(torch.bfloat16)
offs = [int(o) for o in (torch.rand(SLIDES, generator=g) * (pool_rows - 64000)).long().tolist()]
bag = lambda i: pool[offs[i]:offs[i] + sizes[i]]
fc, attn = model.attention_net_WSI[0], model.attention_net_WSI[3]
Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
Wk, bk = model.classifier.weight, model.classifier.bias


@torch.no_grad()
def reference_risk(x):
    h = torch.relu(x.float() @ fc.weight.t() + fc.bias)
    a, gt = torch.tanh(h @ Wa.t() + ba), torch.sigmoid(h @ Wb.t() + bb)
    A = torch.softmax(((a * gt) @ wc.t() + bc).t(), dim=1)
    hz = torch.sigmoid((A @ h) @ Wk.t() + bk)
    return -torch.cumprod(1 - hz, dim=1).sum()


torch.cuda.synchronize(); t0 = time.perf_counter()
ours = torch.empty(SLIDES, device=dev)
with torch.no_grad():
    for i0 in range(0, SLIDES, BATCH):
        ids = range(i0, min(SLIDES, i0 + BATCH))
        hz, S, _, _ = model.infer_cohort([bag(i) for i in ids])
        ours[i0:i0 + len(ids)] = -S.sum(dim=1)
torch.cuda.synchronize(); t_ours = time.perf_counter() - t0
t0 = time.perf_counter()
ref = torch.stack([reference_risk(bag(i)) for i in range(SLIDES)])
torch.cuda.synchronize(); t_ref = time.perf_counter() - t0

d_ref = ref[:, None] - ref[None, :]
d_our = ours[:, None] - ours[None, :]
upper = torch.triu(torch.ones(SLIDES, SLIDES, dtype=torch.bool, device=dev), diagonal=1)
disc = (torch.sign(d_ref) != torch.sign(d_our)) & upper
pairs = int(upper.sum().item())
times = (torch.empty(SLIDES, device=dev).exponential_(1 / 30.0) * torch.exp(-ref)).cpu()     # riskier slides die earlier
event = torch.ones(SLIDES)
sys.path.insert(0, ROOT)
from multimodalfusion_b200.utils.optim import concordance_index
out = {"slides": SLIDES, "patches": sum(sizes), "preset": PRESET,
       "max_abs_risk_diff": (ours - ref).abs().max().item(), "risk_range": (ref.max() - ref.min()).item(),
       "pairs": pairs, "discordant_pairs": int(disc.sum().item()),
       "discordant_pairs_ref_gap_gt_1e-4": int((disc & (d_ref.abs() > 1e-4)).sum().item()),
       "discordant_pairs_ref_gap_gt_1e-3": int((disc & (d_ref.abs() > 1e-3)).sum().item()),
       "largest_ref_gap_of_a_discordant_pair": (d_ref.abs() * disc).max().item(),
       "kendall_tau": 1.0 - 2.0 * disc.sum().item() / pairs,
       "cindex_ours": concordance_index(ours, times, event), "cindex_ref": concordance_index(ref, times, event),
       "seconds_ours_varlen_bf16": t_ours, "seconds_reference_ops_fp32_gpu": t_ref}
print(json.dumps(out))
