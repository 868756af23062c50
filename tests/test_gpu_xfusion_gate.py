"""GPU parity of the fused XlinearFusion gate kernels (mmf_xfusion_gate_fwd / _bwd, csrc/xfusion_gate.cuh) against the
oracle's fp32 restatement (oracle.xfusion_gate; reference models/model_modules.py:156-166): outputs, every weight / bias
gradient and the embedding gradients, 2-4 modalities, both widths of the reference (256 / 16 and 1024 / 64), ragged batch
sizes, with and without the dropout scale mask; then the whole XlinearFusion module (fused gate vs the layer-by-layer path)."""
import pytest
import torch

from oracle import amil_oracle as O

pytestmark = pytest.mark.gpu


def _params(m, dim, g):
    ps = []
    for _ in range(m):
        ps.append((torch.randn(16, dim, generator=g) * dim ** -0.5, torch.randn(16, generator=g) * 0.1,
                   torch.randn(16, dim * m, generator=g) * (dim * m) ** -0.5, torch.randn(16, generator=g) * 0.1,
                   torch.randn(16, 16, generator=g) * 0.25, torch.randn(16, generator=g) * 0.1))
    return ps


@pytest.mark.parametrize("m,dim,B,use_mask,need_dv", [
    (3, 256, 512, True, False), (3, 256, 70, False, True), (2, 256, 1, False, True), (4, 256, 13, True, True),
    (4, 1024, 1, False, True), (4, 1024, 9, True, False), (3, 256, 2048, True, True)])
def test_xfusion_gate_vs_oracle(m, dim, B, use_mask, need_dv):
    from multimodalfusion_b200 import ops
    from multimodalfusion_b200.autograd import XfusionGate
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(100 * m + B)
    vs = [torch.randn(B, dim, generator=g) for _ in range(m)]
    ps = _params(m, dim, g)
    mask = (torch.rand(m, B, 16, generator=g) > 0.7).float() / 0.3 if use_mask else None
    d_o = torch.randn(m, B, 17, generator=g)
    # oracle (CPU, fp32 autograd)
    vr = [v.clone().requires_grad_(need_dv) for v in vs]
    pr = [tuple(t.clone().requires_grad_() for t in p_) for p_ in ps]
    o_ref = O.xfusion_gate(vr, [((p_[0], p_[1]), (p_[2], p_[3]), (p_[4], p_[5])) for p_ in pr], mask)
    o_ref.backward(d_o)
    # kernels
    vg = [v.to(dev).requires_grad_(need_dv) for v in vs]
    pg = [tuple(t.to(dev).requires_grad_() for t in p_) for p_ in ps]
    assert ops.xfusion_gate_supported(vg, pg)
    o = XfusionGate.apply(m, None if mask is None else mask.to(dev), *vg, *[t for p_ in pg for t in p_])
    assert o.shape == (m, B, 17) and torch.all(o[:, :, 16] == 1)
    o.backward(d_o.to(dev))
    torch.testing.assert_close(o.detach().cpu(), o_ref.detach(), rtol=2e-5, atol=2e-6)
    names = ["dWh", "dbh", "dWz", "dbz", "dWo", "dbo"]
    for i in range(m):
        for n, a, b in zip(names, pg[i], pr[i]):
            scale = b.grad.abs().max().item() + 1e-6
            assert (a.grad.cpu() - b.grad).abs().max().item() <= 3e-5 * scale + 1e-6, (i, n)
        if need_dv:
            scale = vr[i].grad.abs().max().item() + 1e-6
            assert (vg[i].grad.cpu() - vr[i].grad).abs().max().item() <= 3e-5 * scale + 1e-6, (i, "dv")
        else:
            assert vg[i].grad is None


@pytest.mark.parametrize("m,train", [(3, False), (2, False), (4, False), (3, True)])
def test_xlinear_fusion_module_fused_gate_equals_layerwise(m, train, monkeypatch):
    """XlinearFusion with the fused gate == the same module run layer by layer (Dense kernels): forward and parameter
    gradients; eval mode exactly comparable, train mode with the dropout draws pinned by a fixed mask."""
    from multimodalfusion_b200 import ops
    from multimodalfusion_b200.models.model_modules import XlinearFusion
    dev = torch.device("cuda")
    torch.manual_seed(7)
    fus = XlinearFusion(num_modalities=m, dropout_rate=0.0 if train else 0.25).to(dev)
    fus.train(train)
    B = 37
    vs = [torch.randn(B, 256, device=dev) for _ in range(m)]

    def run(fused: bool):
        fus.zero_grad(set_to_none=True)
        if not fused:
            monkeypatch.setattr(ops, "xfusion_gate_supported", lambda *_: False)
        out = fus(v_list=[v.clone() for v in vs])
        out.square().sum().backward()
        monkeypatch.undo()
        return out.detach().clone(), {n: p.grad.clone() for n, p in fus.named_parameters() if p.grad is not None}

    o1, g1 = run(True)
    o2, g2 = run(False)
    torch.testing.assert_close(o1, o2, rtol=1e-4, atol=1e-5)
    assert g1.keys() == g2.keys()
    for n in g1:
        scale = g2[n].abs().max().item() + 1e-6
        assert (g1[n] - g2[n]).abs().max().item() <= 1e-4 * scale + 1e-6, n
