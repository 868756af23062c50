"""GPU tests of the CUDA-graph replayed training steps (multimodalfusion_b200.graphs): device-resident Adam step count
and dropout seeds (mmf_step_state_advance, mmf_adam_step_multi_dev, MMF_SEED_DEVICE), graph replays == the eager loop
with the same seed sequence. Reference loop: utils/core_utils.py:184-247 (batch-1, optimizer step per patient)."""
import copy
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def splitmix64_next(x):
    """Host restatement of train_glue.cuh:splitmix64_next."""
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return (z ^ (z >> 31)) & 0x3FFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda")


def test_step_state_advance(dev):
    from multimodalfusion_b200.graphs import StepState
    torch.manual_seed(3)
    st = StepState(dev)
    init = st.buf.tolist()
    assert init[0] == 0 and all(0 <= v < 2 ** 62 for v in init[1:])
    st.set_step(41)
    st.advance(); st.advance()
    now = st.buf.tolist()
    assert now[0] == 43
    for a, b in zip(init[1:], now[1:]):
        assert b == splitmix64_next(splitmix64_next(a))


def test_adam_device_step_matches_host_step(dev):
    from multimodalfusion_b200.graphs import StepState, device_step_state
    from multimodalfusion_b200.utils.optim import FusedAdam
    torch.manual_seed(0)
    pa = [torch.randn(257, 33, device=dev).requires_grad_(), torch.randn(19, device=dev).requires_grad_()]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa, ob = FusedAdam(pa, lr=1e-2, weight_decay=1e-3), FusedAdam(pb, lr=1e-2, weight_decay=1e-3)
    st = StepState(dev)
    for it in range(5):
        gs = [torch.randn_like(p) for p in pa]
        for p, q, g in zip(pa, pb, gs):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        if it == 0:
            ob.step()                      # eager first step (creates the optimizer state), as GraphedStep does
            st.set_step(ob.host_step())
        else:
            st.advance()
            with device_step_state(st):
                ob.step()
            ob.note_graph_step()
    assert ob.host_step() == 5 and int(st.buf[0].item()) == 5
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=0, atol=1e-7), (p - q).abs().max().item()


def test_device_seed_equals_host_seed(dev):
    """A seed read from device memory (MMF_SEED_DEVICE) gives the bit-identical dropout masks of the same host seed:
    fused forward (scores, partials) and the Kronecker encoder with post-fusion dropout."""
    from multimodalfusion_b200 import ops
    from multimodalfusion_b200._lib import SEED_DEVICE_BIT
    from test_gpu_fused_step import _rand
    params, _, _ = _rand(512, 384, True, 4, 5)
    prep = ops.prepare_amil_weights(*[None if p is None else p.to(dev) for p in params])
    x = (0.5 * torch.randn(700, 1024, device=dev).abs()).to(torch.bfloat16)
    flags = ops.amil_flags(True, dropout_h=True, dropout_attn=True)
    seed = 0x1234_5678_9ABC_DEF
    word = torch.tensor([seed], dtype=torch.int64, device=dev)
    A0, P0 = ops.amil_partials(x, prep, flags, seed)
    A1, P1 = ops.amil_partials(x, prep, flags, SEED_DEVICE_BIT | word.data_ptr())
    assert torch.equal(A0, A1) and torch.equal(P0, P1)
    A2, _ = ops.amil_partials(x, prep, flags, seed + 1)
    assert not torch.equal(A0, A2)
    o = [torch.rand(6, 17, device=dev) for _ in range(3)]
    W, b = torch.randn(64, 17 ** 3, device=dev) * 0.02, torch.randn(64, device=dev) * 0.1
    k0 = ops.kron_enc_fwd(o, W, b, dropout=True, seed=seed)
    k1 = ops.kron_enc_fwd(o, W, b, dropout=True, seed=SEED_DEVICE_BIT | word.data_ptr())
    assert torch.equal(k0, k1)
    dout = torch.randn_like(k0)
    g0 = ops.kron_enc_bwd(o, W, k0, dout, dropout=True, seed=seed)
    g1 = ops.kron_enc_bwd(o, W, k0, dout, dropout=True, seed=SEED_DEVICE_BIT | word.data_ptr())
    for a, c in zip(g0[0] + [g0[1], g0[2]], g1[0] + [g1[1], g1[2]]):
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)


def _patched_seeds(monkeypatch, host_seeds):
    """_seed_from_torch of the fused step: device words under capture, the given host seeds otherwise."""
    from multimodalfusion_b200 import graphs
    from multimodalfusion_b200.models import _fused_step

    def seed():
        st = graphs.current_state()
        return st.new_seed() if st is not None else host_seeds.pop(0)
    monkeypatch.setattr(_fused_step, "_seed_from_torch", seed)


@pytest.mark.parametrize("kind,N", [("path", 700), ("radio", 130)])
def test_graphed_fused_step_equals_eager_loop(dev, monkeypatch, kind, N):
    """Six patients of one size: eager first step + five graph replays == six eager steps that are handed the seed
    sequence the device state produces (losses and parameters; the wgrad's TMA reduce-adds are unordered fp32 sums)."""
    from multimodalfusion_b200.models import MIL_Attention_fc_surv_path, MIL_Attention_fc_surv_radio
    from multimodalfusion_b200.utils import get_optim
    args = types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5)
    torch.manual_seed(11)
    if kind == "path":
        mg = MIL_Attention_fc_surv_path(model_size_wsi="small", n_classes=4).to(dev).train()
        names = ["path_features"]
    else:
        mg = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
        names = list(mg.modalities)
    me = copy.deepcopy(mg)
    mg.enable_fused_step(); me.enable_fused_step()
    og, oe = get_optim(mg, args), get_optim(me, args)
    steps = 6
    bags = [{n: (0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for n in names} for _ in range(steps)]
    Ys = [torch.tensor([i % 4], device=dev) for i in range(steps)]
    cs = [torch.tensor([float(i % 2)], device=dev) for i in range(steps)]
    first_seed = 4242
    host = [first_seed]
    _patched_seeds(monkeypatch, host)
    losses_g = []
    for i in range(steps):
        out = mg.graphed_fused_step(og, Y=Ys[i], c=cs[i], alpha=0.0, **bags[i])
        losses_g.append(out[4].item())
        if i == 0:
            w = int(mg._graph_family.state.buf[1].item())    # seed word 1 before the first replay
    assert og.host_step() == steps
    seq, s = [first_seed], w
    for _ in range(steps - 1):
        s = splitmix64_next(s)
        seq.append(s)
    assert int(mg._graph_family.state.buf[1].item()) == seq[-1]
    host[:] = seq
    losses_e = []
    for i in range(steps):
        out = me.fused_step(Y=Ys[i], c=cs[i], alpha=0.0, **bags[i])
        oe.step(zero_grad=False)
        losses_e.append(out[4].item())
    assert losses_g == pytest.approx(losses_e, rel=2e-3), (losses_g, losses_e)
    lr = args.lr
    for (n, p), q in zip(mg.named_parameters(), me.parameters()):
        d = (p - q).abs()
        assert d.max().item() <= 2.5 * lr, n                       # (an Adam step is at most ~lr per element)
        # (unordered fp32 sums in the wgrad: Adam turns the noise of a near-zero gradient into a fraction of a step)
        assert (d > 0.1 * lr).float().mean().item() < 1e-2, n
    # a different seed sequence must NOT reproduce the losses (the masks matter)
    assert abs(losses_g[1] - losses_g[0]) > 0 and len(set(losses_g)) == steps


def test_graphed_radio_family_many_sizes(dev):
    """Patients of several slice counts (incl. a tiny exact-fp32 bag): one graph per size, shared step count and seeds."""
    from multimodalfusion_b200.models import MIL_Attention_fc_surv_radio
    from multimodalfusion_b200.utils import get_optim
    torch.manual_seed(2)
    model = MIL_Attention_fc_surv_radio(gate_radio=True, dropout=True, n_classes=4).to(dev).train()
    model.enable_fused_step()
    opt = get_optim(model, types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5))
    sizes = [100, 155, 40, 100, 155, 40, 100, 129, 129]
    Y, c = torch.tensor([2], device=dev), torch.tensor([0.0], device=dev)
    p0 = [p.detach().clone() for p in model.parameters()]
    for n in sizes:
        bag = {m: (0.5 * torch.randn(n, 1024, device=dev).abs()).to(torch.bfloat16) for m in model.modalities}
        hz, S, Y_hat, A_raw, loss = model.graphed_fused_step(opt, Y=Y, c=c, alpha=0.0, **bag)
        assert A_raw.shape[-1] == n and torch.isfinite(loss).item() and torch.isfinite(hz).all().item()
    assert len(model._graph_family.graphs) == 4
    assert opt.host_step() == len(sizes) and int(model._graph_family.state.buf[0].item()) == len(sizes)
    assert all((p - q).abs().max().item() > 0 for p, q in zip(model.parameters(), p0) if p.numel() > 1)
    # eager use after replays sees the replayed parameters (the weight-copy caches were invalidated)
    model.eval()
    bag = {m: (0.5 * torch.randn(100, 1024, device=dev).abs()).to(torch.bfloat16) for m in model.modalities}
    with torch.no_grad():
        hz_a = model(**bag)[0].clone()
        model.attention_net_radio._mmf_prep = None
        hz_b = model(**bag)[0]
    assert torch.allclose(hz_a, hz_b, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("loss_kind", ["cox", "ranking"])
def test_graphed_cohort_head_step(dev, loss_kind):
    """BASELINE config 3 (Kronecker head + Cox / ranking loss over a cohort, autograd + FusedAdam) as a replayed graph.
    eval-mode heads (no dropout anywhere): replays == eager steps; train mode: runs, fresh seeds per replay."""
    from multimodalfusion_b200.graphs import GraphedStep
    from multimodalfusion_b200.models import coxranking_models_pretrained as cox_heads
    from multimodalfusion_b200.utils import CoxSurvLoss, RankingSurvLoss, get_optim
    args = types.SimpleNamespace(opt="adam", lr=2e-4, reg=1e-5)
    B = 96
    torch.manual_seed(5)
    hg = cox_heads.multimodal_pretrained(mode="radio_path_omic", train_type="kronecker", n_classes=4).to(dev).eval()
    he = copy.deepcopy(hg)
    og, oe = get_optim(hg, args), get_optim(he, args)
    lf = CoxSurvLoss() if loss_kind == "cox" else RankingSurvLoss()
    semb = [torch.zeros(B, 256, device=dev) for _ in range(3)]
    times = ((torch.empty(B, device=dev).exponential_(1 / 30.0).clamp_(0, 250) * 2).round() / 2)
    cens = (torch.rand(B, device=dev) < 0.46).float()

    def step(head, opt, emb):
        risk = head(*emb)[0]
        loss = lf(risks=risk.reshape(-1), times=times, c=cens)
        loss.backward()
        opt.step(zero_grad=True)
        return loss.detach()

    graphed = GraphedStep(lambda: step(hg, og, semb), optimizers=(og,), modules=(hg,))
    lg, le = [], []
    for it in range(5):
        emb = [torch.randn(B, 256, device=dev) for _ in range(3)]
        for s_, e_ in zip(semb, emb):
            s_.copy_(e_)
        lg.append(graphed().item())
        le.append(step(he, oe, emb).item())
    assert graphed.replays == 4 and og.host_step() == 5
    assert lg == pytest.approx(le, rel=1e-4, abs=1e-6), (lg, le)
    for p, q in zip(hg.parameters(), he.parameters()):
        assert (p - q).abs().max().item() <= 2.5 * args.lr
    # train mode: the post-fusion dropout seed is a device word that moves on with every replay
    ht = cox_heads.multimodal_pretrained(mode="radio_path_omic", train_type="kronecker", n_classes=4).to(dev).train()
    ot = get_optim(ht, args)
    gt = GraphedStep(lambda: step(ht, ot, semb), optimizers=(ot,), modules=(ht,))
    seen = []
    for it in range(4):
        assert torch.isfinite(gt()).item()
        seen.append(tuple(gt.state.buf[1:3].tolist()))
    assert len(set(seen)) == len(seen)          # the seed words move on with every replay
    assert ot.host_step() == 4 and int(gt.state.buf[0].item()) == 4
