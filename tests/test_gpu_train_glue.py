"""Training-step glue kernels against plain PyTorch / the oracle on the CPU:
fused multi-tensor Adam (+ weight decay, + l1_reg_all folded in, + fused zero_grad) == torch.optim.Adam on
`loss + lambda * sum|W|`; device concordance index == the oracle's restatement of sksurv's; and a short training
loop through the drop-in model with the fused optimizer."""
import types

import pytest
import torch

from helpers import rel_err
from oracle import amil_oracle as O
from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda")


@pytest.mark.parametrize("wd,l1,gscale", [(0.0, 0.0, 1.0), (1e-5, 0.0, 1.0), (1e-2, 1e-3, 0.25)])
def test_fused_adam_matches_torch_adam(dev, wd, l1, gscale):
    from multimodalfusion_b200.utils import FusedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(512, 1024), (512,), (768, 512), (1, 384), (1,), (4, 512), (3, 7, 5)] + [(17,)] * 60   # > 48 tensors
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    ours = [t.detach().clone().to(dev).requires_grad_(True) for t in ref]
    opt_ref = torch.optim.Adam(ref, lr=2e-3, weight_decay=wd)
    opt = FusedAdam(ours, lr=2e-3, weight_decay=wd, l1_lambda=l1)
    for step in range(6):
        grads = [torch.randn(s, generator=g) for s in shapes]
        l1_ref = sum(t.detach().abs().sum() for t in ref)
        for t, gr in zip(ref, grads):
            t.grad = gr * gscale + l1 * torch.sign(t.detach())     # d/dW (loss/gc + lambda * sum|W|)
        for t, gr in zip(ours, grads):
            t.grad = gr.to(dev)
        opt_ref.step()
        v0 = ours[0]._version
        opt.step(zero_grad=True, grad_scale=gscale)
        assert ours[0]._version > v0, "in-place update must invalidate cached weight copies"
        assert abs(opt.l1_value.item() / l1_ref.item() - 1) < 1e-5
        assert all(torch.count_nonzero(t.grad).item() == 0 for t in ours)
    for a, b in zip(ours, ref):
        assert rel_err(a, b) < 2e-6


@pytest.mark.parametrize("B", [2, 37, 512, 3000])
def test_cindex_matches_oracle(dev, B):
    from multimodalfusion_b200.utils import concordance_index
    g = torch.Generator().manual_seed(B)
    risk = torch.randn(B, generator=g)
    risk[B // 3] = risk[0]                       # an exact tie in risk
    times, c = cases.cohort_labels(B, 11)        # ties in time included
    event = 1 - c
    got = concordance_index(risk.to(dev), times.to(dev), event.to(dev))
    want = O.concordance_index(risk, times, event)
    assert (got != got and want != want) or abs(got - want) < 1e-12
    # discretised times (months): events tied in time with censored samples are comparable pairs (sksurv semantics)
    times_m = torch.floor(times / times.max() * 12)
    got = concordance_index(risk.to(dev), times_m.to(dev), event.to(dev))
    want = O.concordance_index(risk, times_m, event)
    assert (got != got and want != want) or abs(got - want) < 1e-12


def test_short_training_loop_with_fused_optimizer(dev):
    """utils/core_utils.py:200-247 with the drop-in model, NLLSurvLoss, get_optim(adam, reg=1e-5): the loss on a
    fixed bag goes down and the fc / attention weights move (weight cache invalidation through the fused step)."""
    from multimodalfusion_b200.models import MIL_Attention_fc_surv_path
    from multimodalfusion_b200.utils import NLLSurvLoss, get_optim
    torch.manual_seed(0)
    model = MIL_Attention_fc_surv_path(gate_path=True, model_size_wsi="small", n_classes=4).to(dev).eval()
    opt = get_optim(model, types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5))
    x = cases.features(700, 3).to(dev)
    Y, c = torch.tensor([1], device=dev), torch.tensor([0.0], device=dev)
    loss_fn = NLLSurvLoss(alpha=0.0)
    w0 = model.attention_net_WSI[0].weight.detach().clone()
    losses = []
    for _ in range(8):
        hz, S, _, _ = model(path_features=x)
        loss = loss_fn(hazards=hz, S=S, Y=Y, c=c)
        loss.backward()
        opt.step(zero_grad=True)
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 1e-3, losses
    assert (model.attention_net_WSI[0].weight - w0).abs().max().item() > 0


@pytest.mark.parametrize("n,nq", [(1, 1), (1000, 1000), (5000, 300)])
def test_percentiles_match_scipy(dev, n, nq):
    """utils/wsi_utils.py:171-174 / utils/heatmap_utils.py:32-34: scipy.stats.percentileofscore per score."""
    from scipy.stats import percentileofscore
    from multimodalfusion_b200.utils import to_percentiles
    g = torch.Generator().manual_seed(n)
    ref = torch.randn(n, generator=g).round(decimals=2)          # rounding -> plenty of ties
    q = ref if nq == n else torch.randn(nq, generator=g).round(decimals=2)
    got = to_percentiles(q.to(dev), None if nq == n else ref.to(dev)).cpu()
    want = torch.tensor([percentileofscore(ref.numpy(), float(v)) for v in q.numpy()], dtype=torch.float32)
    assert torch.allclose(got, want, rtol=0, atol=1e-4)
