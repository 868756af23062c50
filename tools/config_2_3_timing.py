"""BASELINE.json configs 2 and 3 through the drop-in modules (eager, batch-1 loop as in the reference):
  2. radio_attention_mil: 256 synthetic patients, 4 modalities x [N,1024], N ~ U{80..155}, nll_surv, Adam step per patient;
  3. multimodal Kronecker head + Cox / ranking loss over a 512-patient cohort of 256-d embeddings, fwd + bwd + Adam."""
import json, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
from multimodalfusion_b200.models import coxranking_models_pretrained as cox_heads
from multimodalfusion_b200.utils import CoxSurvLoss, NLLSurvLoss, RankingSurvLoss, get_optim
dev = torch.device("cuda")
torch.manual_seed(0)
args = types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5)
# ---- config 2
model = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
opt = get_optim(model, args)
loss_fn = NLLSurvLoss(alpha=0.0)
g = torch.Generator().manual_seed(1)
ns = torch.randint(80, 156, (256,), generator=g).tolist()
names = model.modalities
bags = [{m: (0.5 * torch.randn(n, 1024, device=dev).abs()).to(torch.bfloat16) for m in names} for n in ns[:32]]
Y, c = torch.tensor([1], device=dev), torch.tensor([0.0], device=dev)


def radio_epoch(count):
    for i in range(count):
        hz, S, _, _ = model(**bags[i % len(bags)])
        loss = loss_fn(hazards=hz, S=S, Y=Y, c=c)
        loss.backward()
        opt.step(zero_grad=True)


radio_epoch(8)
torch.cuda.synchronize()
t0 = time.perf_counter(); radio_epoch(256); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"config": "2 radio_attention_mil batch-1 training loop", "patients": 256, "ms_per_patient": dt / 256 * 1e3,
                  "patients_per_s": 256 / dt, "slices_per_s": sum(ns) / dt,
                  "note": "eager launches from Python (fwd + loss + bwd + fused Adam per patient), wall clock incl. host"}))
# ---- config 2 through the fused step (no autograd graph): reduce_dim + 3-launch step + dx GEMM + reduce_dim wgrad + Adam
model.enable_fused_step()
opt = get_optim(model, args)


def radio_epoch_fused(count):
    for i in range(count):
        model.fused_step(Y=Y, c=c, alpha=0.0, **bags[i % len(bags)])
        opt.step(zero_grad=False)     # (the next fused step clears the gradients inside its forward kernel)


radio_epoch_fused(8)
torch.cuda.synchronize()
t0 = time.perf_counter(); radio_epoch_fused(256); torch.cuda.synchronize(); dt = time.perf_counter() - t0
# launches per patient, counted with the profiler's kernel events of 4 patients
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    radio_epoch_fused(4); torch.cuda.synchronize()
launches = sum(e.count for e in prof.key_averages() if e.device_type.name == "CUDA") / 4
if os.environ.get("MMF_CFG2_KERNELS"):
    for e_ in sorted((e for e in prof.key_averages() if e.device_type.name == "CUDA"), key=lambda e: -e.device_time_total)[:30]:
        print(f"  {e_.device_time_total / 4:8.1f} us  x{e_.count / 4:4.1f}  {e_.key[:150]}")
print(json.dumps({"config": "2 radio_attention_mil batch-1 training loop, fused_step", "patients": 256, "ms_per_patient": dt / 256 * 1e3,
                  "patients_per_s": 256 / dt, "kernel_launches_per_patient": launches,
                  "note": "MIL_Attention_fc_surv_radio.fused_step + fused Adam per patient, eager launches from Python, wall clock incl. host"}))
# ---- config 2, the same patient step (fused_step + fused Adam) replayed as ONE CUDA graph per slice count
model.train()
opt = get_optim(model, args)


def radio_epoch_graphed(count):
    for i in range(count):
        model.graphed_fused_step(opt, Y=Y, c=c, alpha=0.0, **bags[i % len(bags)])


radio_epoch_graphed(2 * len(bags))      # every slice count of the 32 patients is captured here (first visit: eager step + capture)
torch.cuda.synchronize()
t0 = time.perf_counter(); radio_epoch_graphed(256); torch.cuda.synchronize(); dt = time.perf_counter() - t0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
entry = next(iter(model._graph_family.graphs.values()))
ev0.record()
for _ in range(50):
    entry.graph.replay()
ev1.record(); torch.cuda.synchronize()
print(json.dumps({"config": "2 radio_attention_mil batch-1 training loop, graphed_fused_step", "patients": 256,
                  "ms_per_patient": dt / 256 * 1e3, "patients_per_s": 256 / dt, "graphs": len(model._graph_family.graphs),
                  "graph_replay_gpu_ms": ev0.elapsed_time(ev1) / 50,
                  "launches_per_patient_from_python": 4,
                  "note": "fused_step + fused Adam per patient as one CUDA-graph launch per slice count (device-resident Adam step "
                          "count and dropout seeds), + 1 stack and 2 small copies into the static inputs; wall clock incl. host"}))
# ---- config 3 (COHORT=32: the batch size of the reference's own command, commands/commands.sh:69-72 `--batch_size 32`)
B = int(os.environ.get("COHORT", 512))
head = cox_heads.multimodal_pretrained(mode="radio_path_omic", train_type="kronecker", n_classes=4).to(dev).train()
opt3 = get_optim(head, args)
emb = [torch.randn(B, 256, device=dev) for _ in range(3)]
times = (torch.empty(B, device=dev).exponential_(1 / 30.0).clamp_(0, 250) * 2).round() / 2
cens = (torch.rand(B, device=dev) < 0.46).float()
for name, lf in (("cox", CoxSurvLoss()), ("ranking", RankingSurvLoss())):
    def it():
        risk, _, _ = head(*emb)
        loss = lf(risks=risk.reshape(-1), times=times, c=cens) if name == "ranking" else lf(risks=risk, times=times, c=cens)
        loss.backward()
        opt3.step()
        opt3.zero_grad(set_to_none=True)    # (torch's default, as the reference's optimizer.zero_grad(): the next backward's
        return loss.detach()                # gradients are adopted, not added — no accumulate kernels)
    for _ in range(3):
        it()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        l = it()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(json.dumps({"config": f"3 multimodal kronecker head + {name} loss, B={B} cohort", "ms_per_step": dt * 1e3,
                      "patients_per_s": B / dt, "loss": l.item(),
                      "note": "fwd + loss + bwd + fused Adam, eager, wall clock; the reference's Cox / ranking host loops alone take 2.5 s / 7.2 s at B=512 (SURVEY.md)"}))

    # the same cohort step replayed as one CUDA graph (static embeddings; device-resident Adam step count / dropout seeds)
    from multimodalfusion_b200.graphs import GraphedStep
    graphed = GraphedStep(it, optimizers=(opt3,), modules=(head,))
    for _ in range(3):
        graphed()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        l = graphed()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(json.dumps({"config": f"3 multimodal kronecker head + {name} loss, B={B} cohort, CUDA-graph replay", "ms_per_step": dt * 1e3,
                      "patients_per_s": B / dt, "loss": l.item(),
                      "note": "GraphedStep(fwd + loss + bwd + fused Adam): one graph launch per cohort step, wall clock"}))
