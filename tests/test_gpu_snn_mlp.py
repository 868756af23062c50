"""GPU parity of the fused SNN MLP (mmf_snn_mlp_fwd / _bwd, csrc/snn_mlp.cuh) against torch fp32 autograd of the reference's
own op sequence (models/model_modules.py:64-68: Linear -> SELU -> AlphaDropout per block) with the SAME keep masks —
alpha-dropout restated from its definition (oracle: torch.nn.functional.alpha_dropout's affine) — outputs, weight / bias
gradients and the input gradient; then MaxNet end to end (fused trunk vs block by block)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

ALPHA_P = -1.7580993408473766


def alpha_drop(y, keep, p):
    a = ((1 - p) * (1 + p * ALPHA_P ** 2)) ** -0.5
    return a * (y * keep + ALPHA_P * (1 - keep)) - a * ALPHA_P * p


def reference_mlp(x, layers):
    for W, b, keep, p in layers:
        x = torch.selu(x @ W.t() + b)
        if keep is not None:
            x = alpha_drop(x, keep, p)
    return x


def test_alpha_dropout_restatement_matches_torch():
    torch.manual_seed(0)
    y = torch.randn(4000, 64)
    out = torch.nn.functional.alpha_dropout(y, 0.25, True)
    keep = (out != (((1 - 0.25) * (1 + 0.25 * ALPHA_P ** 2)) ** -0.5 * ALPHA_P * (1 - 0.25))).float()   # dropped -> the constant
    torch.testing.assert_close(alpha_drop(y, keep, 0.25), out, rtol=1e-5, atol=1e-6)
    assert abs(keep.mean().item() - 0.75) < 0.01


@pytest.mark.parametrize("widths,B,masks,need_dx", [
    ((37, 256, 256), 1, False, True), ((37, 256, 256), 70, True, False), ((80, 256, 256), 512, True, True),
    ((1500, 1024, 256), 5, True, True), ((300, 64), 9, False, False), ((20, 128, 96, 64, 32), 33, True, True)])
def test_snn_mlp_vs_reference_ops(widths, B, masks, need_dx):
    from multimodalfusion_b200.autograd import SnnMlp
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(sum(widths) + B)
    x = torch.randn(B, widths[0], generator=g)
    layers, p = [], 0.25
    for d_in, d_out in zip(widths[:-1], widths[1:]):
        keep = (torch.rand(B, d_out, generator=g) > p).float() if masks else None
        layers.append((torch.randn(d_out, d_in, generator=g) * d_in ** -0.5, torch.randn(d_out, generator=g) * 0.1, keep, p))
    dout = torch.randn(B, widths[-1], generator=g)
    xr = x.clone().requires_grad_(need_dx)
    lr = [(W.clone().requires_grad_(), b.clone().requires_grad_(), k, p_) for W, b, k, p_ in layers]
    ref = reference_mlp(xr, lr)
    ref.backward(dout)
    xg = x.to(dev).requires_grad_(need_dx)
    lg = [(W.to(dev).requires_grad_(), b.to(dev).requires_grad_(), None if k is None else k.to(dev), p_) for W, b, k, p_ in layers]
    out = SnnMlp.apply(xg, tuple(p_ if k is not None else 0.0 for _, _, k, p_ in lg), tuple(k for _, _, k, _ in lg),
                       *[t for W, b, _, _ in lg for t in (W, b)])
    out.backward(dout.to(dev))
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=3e-5, atol=3e-6)
    for i, ((Wg, bg, _, _), (Wr, br, _, _)) in enumerate(zip(lg, lr)):
        for n, a, r in (("dW", Wg.grad, Wr.grad), ("db", bg.grad, br.grad)):
            assert (a.cpu() - r).abs().max().item() <= 5e-5 * (r.abs().max().item() + 1e-6) + 1e-6, (i, n)
    if need_dx:
        assert (xg.grad.cpu() - xr.grad).abs().max().item() <= 5e-5 * (xr.grad.abs().max().item() + 1e-6) + 1e-6
    else:
        assert xg.grad is None


@pytest.mark.parametrize("size,B,one_d", [("small", 6, False), ("big", 3, False), ("small", 1, True)])
def test_maxnet_fused_trunk_equals_block_by_block(size, B, one_d, monkeypatch):
    """MaxNet (eval mode) with the fused SNN trunk == the same model run block by block on the Dense kernels: hazards and
    every parameter gradient."""
    from multimodalfusion_b200 import ops
    from multimodalfusion_b200.models import MaxNet
    dev = torch.device("cuda")
    torch.manual_seed(3)
    model = MaxNet(input_dim=80, model_size_omic=size, bag_loss="nll_surv", n_classes=4).to(dev).eval()
    x = torch.randn(80, device=dev) if one_d else torch.randn(B, 80, device=dev)

    def run(fused):
        model.zero_grad(set_to_none=True)
        if not fused:
            monkeypatch.setattr(ops, "snn_mlp_supported", lambda *_: False)
        hz, S, _, _ = model(genomic_features=x)
        (hz.sum() + S.sum()).backward()
        monkeypatch.undo()
        return hz.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}

    h1, g1 = run(True)
    h2, g2 = run(False)
    torch.testing.assert_close(h1, h2, rtol=1e-5, atol=1e-6)
    for n in g1:
        assert (g1[n] - g2[n]).abs().max().item() <= 1e-4 * (g2[n].abs().max().item() + 1e-6) + 1e-7, n
