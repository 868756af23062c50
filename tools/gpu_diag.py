"""GPU bring-up diagnostics: runs every kernel family against the oracle and prints error metrics
(no asserts). Each group runs in its own subprocess so a faulting kernel cannot poison the rest.

    python tools/gpu_diag.py            # all groups
    python tools/gpu_diag.py amil_fwd   # one group, in-process
"""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True

GROUPS = ["basic", "gemm", "amil_fwd", "amil_bwd", "amil_big", "amil_ungated", "small", "models", "timing"]


def rel(a, b):
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def run_group(name):
    import torch
    from multimodalfusion_b200 import ops
    from oracle import amil_oracle as O
    from oracle import cases
    dev = torch.device("cuda")
    torch.manual_seed(0)

    def bfr(t):
        return t.to(torch.bfloat16).float()

    def rand_amil(L, D, gated, scale=1.0):
        W1 = torch.randn(L, 1024) * (2.0 / (1024 + L)) ** 0.5
        b1 = torch.randn(L) * 0.05
        Wa = torch.randn(D, L) * (2.0 / (L + D)) ** 0.5
        ba = torch.randn(D) * 0.05
        Wb = torch.randn(D, L) * (2.0 / (L + D)) ** 0.5 if gated else None
        bb = torch.randn(D) * 0.05 if gated else None
        wc = torch.randn(1, D) * (2.0 / (D + 1)) ** 0.5 * scale
        bc = torch.randn(1) * 0.05
        return W1, b1, Wa, ba, Wb, bb, wc, bc

    def amil_case(N, L, D, gated, flags_extra=0, seed=0, check_bwd=True, tag=""):
        W = rand_amil(L, D, gated)
        x = cases.features(N, 5)
        Wd = [None if t is None else t.to(dev) for t in W]
        prep = ops.prepare_amil_weights(*Wd)
        flags = ops.amil_flags(gated) | flags_extra
        xb = x.to(dev).to(torch.bfloat16)
        A_raw, M, ml = ops.amil_forward(xb, prep, flags, seed)
        torch.cuda.synchronize()
        # oracle with the operands the kernel sees (bf16 W, bf16 h)
        W1, b1, Wa, ba, Wb, bb, wc, bc = W
        hs = as_ = gs = None
        if flags_extra & 2:
            hs = O.dropout_scale_mask(seed, 0, N, L)
        if flags_extra & 4:
            as_ = O.dropout_scale_mask(seed, 1, N, D)
            gs = O.dropout_scale_mask(seed, 2, N, D)
        s, h, a, g = O.fc_attention(x, bfr(W1), b1, bfr(Wa), ba, None if Wb is None else bfr(Wb), bb, wc, bc,
                                    h_scale=hs, a_scale=as_, g_scale=gs, round_h=True)
        Mo, m, l = O.softmax_pool(s, h)
        s32, h32, _, _ = O.fc_attention(x, W1, b1, Wa, ba, Wb, bb, wc, bc, h_scale=hs, a_scale=as_, g_scale=gs)
        M32, _, _ = O.softmax_pool(s32, h32)
        print(f"[{tag}] N={N} L={L} D={D} gated={gated}: A_raw rel(bf16-oracle)={rel(A_raw, s):.2e} "
              f"rel(fp32)={rel(A_raw, s32):.2e} | M rel(bf16-oracle)={rel(M, Mo):.2e} rel(fp32)={rel(M, M32):.2e} "
              f"| m {ml[0].item():.5f} vs {m.item():.5f}  l {ml[1].item():.5f} vs {l.item():.5f}", flush=True)
        if not check_bwd:
            return
        dM = torch.randn(L) * 0.1
        dA = torch.randn(N) * 0.01
        gr = ops.amil_backward(xb, prep, flags | 8, seed, A_raw, ml, M, dM.to(dev), dA.to(dev))
        torch.cuda.synchronize()
        go = O.amil_backward(x, bfr(W1), bfr(Wa), None if Wb is None else bfr(Wb), wc, s, h, a, g, Mo, m, l, dM, dA,
                             drop_h=bool(flags_extra & 2), a_scale=as_, g_scale=gs, need_dx=True)
        for k in ("dW1", "db1", "dWab", "dbab", "dwc", "dbc", "dx"):
            print(f"    grad {k:5s} rel={rel(gr[k], go[k]):.2e}  |ref|max={go[k].abs().max().item():.3e}", flush=True)

    if name == "basic":
        from multimodalfusion_b200 import _lib
        print("version", _lib.lib().mmf_version(), torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
        x = torch.randn(1000, 1027)
        y = ops.to_bf16(x.to(dev))
        print("cast exact:", torch.equal(y.cpu(), x.to(torch.bfloat16)))
        Wab = torch.randn(512, 256).to(torch.bfloat16).to(dev)
        packed = torch.empty_like(Wab)
        from multimodalfusion_b200._lib import check, lib
        check(lib().mmf_pack_wab(Wab.data_ptr(), packed.data_ptr(), 256, 256, 1, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = torch.cat([torch.cat([Wab[c * 128:(c + 1) * 128], Wab[256 + c * 128:256 + (c + 1) * 128]]) for c in range(2)])
        print("pack exact:", torch.equal(packed, ref))
    elif name == "gemm":
        for (M, Kseg, nseg, N) in [(128, 64, 1, 256), (300, 128, 2, 256), (155, 1024, 4, 1024)]:
            segs = [bfr(torch.randn(M, Kseg) * 0.5) for _ in range(nseg)]
            W = bfr(torch.randn(N, Kseg * nseg) * 0.05)
            b = torch.randn(N)
            y = ops.linear_bf16([s.to(dev).to(torch.bfloat16) for s in segs], W.to(dev).to(torch.bfloat16), b.to(dev),
                                torch.float32)
            torch.cuda.synchronize()
            ref = torch.cat(segs, 1) @ W.t() + b
            print(f"[gemm KK] M={M} K={Kseg}x{nseg} N={N}: rel={rel(y, ref):.2e}", flush=True)
        for (M, N, Kseg, nseg) in [(64, 128, 256, 1), (300, 128, 256, 1), (1000, 256, 1024, 4)]:
            dY = bfr(torch.randn(M, N) * 0.1)
            segs = [bfr(torch.randn(M, Kseg) * 0.5) for _ in range(nseg)]
            dW = torch.zeros(N, Kseg * nseg, device=dev)
            db = torch.zeros(N, device=dev)
            ops.linear_bf16_wgrad(dY.to(dev).to(torch.bfloat16), [s.to(dev).to(torch.bfloat16) for s in segs], dW, db)
            torch.cuda.synchronize()
            ref = dY.t() @ torch.cat(segs, 1)
            print(f"[gemm MN/MN wgrad] M={M} N={N} K={Kseg}x{nseg}: dW rel={rel(dW, ref):.2e} db rel={rel(db, dY.sum(0)):.2e}",
                  flush=True)
    elif name == "amil_fwd":
        for N in (1, 128, 200, 1000):
            amil_case(N, 256, 256, True, check_bwd=False, tag="fwd")
        amil_case(300, 256, 256, True, flags_extra=2 | 4, seed=0x1234567, check_bwd=False, tag="fwd-dropout")
    elif name == "amil_bwd":
        amil_case(200, 256, 256, True, tag="bwd")
        amil_case(129, 256, 256, True, tag="bwd")
        amil_case(300, 256, 256, True, flags_extra=2 | 4, seed=0x1234567, tag="bwd-dropout")
    elif name == "amil_big":
        amil_case(300, 512, 384, True, tag="big")
        amil_case(1000, 256, 384, True, tag="mm-big")
    elif name == "amil_ungated":
        amil_case(200, 256, 256, False, tag="ungated")
        amil_case(300, 512, 384, False, tag="ungated-big")
    elif name == "small":
        B, I, Oo = 37, 186, 256
        x, W, b = torch.randn(B, I), torch.randn(Oo, I) * 0.1, torch.randn(Oo) * 0.1
        for act, fn in ((0, lambda t: t), (1, torch.relu), (2, torch.selu), (3, torch.sigmoid), (4, torch.tanh)):
            y = ops.dense_fwd(x.to(dev), W.to(dev), b.to(dev), act)
            xr, Wr, br = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yr = fn(xr @ Wr.t() + br)
            dy = torch.randn_like(yr)
            yr.backward(dy)
            dx, dW, db = ops.dense_bwd(x.to(dev), W.to(dev), act, y, dy.to(dev))
            torch.cuda.synchronize()
            print(f"[dense act={act}] y={rel(y, yr):.2e} dx={rel(dx, xr.grad):.2e} dW={rel(dW, Wr.grad):.2e} "
                  f"db={rel(db, br.grad):.2e}", flush=True)
        for m in (2, 3):
            E, H, B = 17, 64, 9
            o = [torch.rand(B, E).requires_grad_(True) for _ in range(m)]
            W = (torch.randn(H, E ** m) * 0.05).requires_grad_(True)
            b = (torch.randn(H) * 0.1).requires_grad_(True)
            fused = o[0]
            for t in o[1:]:
                fused = (fused[:, :, None] * t[:, None, :]).flatten(1)
            ref = torch.relu(fused @ W.t() + b)
            dout = torch.randn_like(ref)
            ref.backward(dout)
            od = [t.detach().to(dev) for t in o]
            out = ops.kron_enc_fwd(od, W.detach().to(dev), b.detach().to(dev))
            d_o, dW, db = ops.kron_enc_bwd(od, W.detach().to(dev), out, dout.to(dev))
            torch.cuda.synchronize()
            print(f"[kron m={m}] out={rel(out, ref):.2e} dW={rel(dW, W.grad):.2e} db={rel(db, b.grad):.2e} "
                  + " ".join(f"do{i}={rel(d_o[i], o[i].grad):.2e}" for i in range(m)), flush=True)
        # hazard head + nll
        B, Lin, K = 5, 256, 4
        Mx = torch.randn(B, Lin).requires_grad_(True)
        Wk = (torch.randn(K, Lin) * 0.1).requires_grad_(True)
        bk = (torch.randn(K) * 0.1).requires_grad_(True)
        hz, S, Yh = O.hazard_head(Mx, Wk, bk)
        Y = torch.randint(0, K, (B,)); c = (torch.rand(B) < 0.5).float()
        loss = O.nll_surv_loss(hz, S, Y, c, alpha=0.15)
        loss.backward()
        hz_g, S_g, Y_g = ops.hazard_head_fwd(Mx.detach().to(dev), Wk.detach().to(dev), bk.detach().to(dev))
        l_g, dh_g, dS_g = ops.nll_surv(hz_g, S_g, Y.to(dev), c.to(dev), 0.15)
        dM_g, dWk_g, dbk_g = ops.hazard_head_bwd(Mx.detach().to(dev), Wk.detach().to(dev), hz_g, S_g, dh_g, dS_g)
        torch.cuda.synchronize()
        print(f"[head] haz={rel(hz_g, hz):.2e} S={rel(S_g, S):.2e} Yhat_eq={torch.equal(Y_g.cpu(), Yh)} "
              f"loss={abs(l_g.item() - loss.item()):.2e} dM={rel(dM_g, Mx.grad):.2e} dWk={rel(dWk_g, Wk.grad):.2e} "
              f"dbk={rel(dbk_g, bk.grad):.2e}", flush=True)
        for B in (2, 64, 200, 512, 2048):
            r = torch.randn(B).requires_grad_(True)
            times, c = cases.cohort_labels(B, B)
            lo = O.cox_loss(r, times, c)
            lo.backward()
            lg, dg = ops.cox(r.detach().to(dev), times.to(dev), c.to(dev))
            torch.cuda.synchronize()
            print(f"[cox B={B}] loss {lg.item():.6f} vs {lo.item():.6f}  dtheta rel={rel(dg, r.grad):.2e}", flush=True)
        for B, phi, red in ((2, "sigmoid", "mean"), (33, "sigmoid", "mean"), (40, "relu", "sum"), (512, "sigmoid", "mean")):
            r = torch.randn(B).requires_grad_(True)
            times, c = cases.cohort_labels(B, B + 1)
            lo = O.ranking_loss(r, times, c, phi, red).reshape(())
            if lo.requires_grad:
                lo.backward()
            gref = r.grad if r.grad is not None else torch.zeros(B)
            lg, dg, npairs = ops.ranking(r.detach().to(dev), times.to(dev), c.to(dev), phi, red)
            torch.cuda.synchronize()
            print(f"[rank B={B} {phi} {red}] loss {lg.item():.6f} vs {lo.item():.6f} pairs={npairs.item()} "
                  f"dr maxabs err={(dg.cpu() - gref).abs().max().item():.2e}", flush=True)
    elif name == "models":
        from helpers import build_head_model, build_omic_model, build_path_model, build_radio_model
        from multimodalfusion_b200.utils import CoxSurvLoss, NLLSurvLoss, RankingSurvLoss
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.pt"), weights_only=False)
        for nm, cfg in cases.PATH_CASES.items():
            gd = gold["path"][nm]
            model = build_path_model(cfg).to(dev)
            x = cases.path_bag(cfg).to(dev)
            Y, c = cases.labels(cfg)
            hz, S, Yh, A = model(path_features=x)
            loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hz, S=S, Y=Y.to(dev), c=c.to(dev))
            model.zero_grad(); loss.backward(); torch.cuda.synchronize()
            worst = 0.0
            for k, p in model.named_parameters():
                fp = gd["grads"][k]
                idx = cases._sample_idx(p.numel())
                e = (p.grad.reshape(-1).cpu()[idx] - fp["vals"]).abs().max().item() / max(fp["vals"].abs().max().item(), 1e-30)
                if fp["vals"].abs().max().item() > 1e-6:
                    worst = max(worst, e)
            print(f"[path {nm}] A={rel(A, gd['A_raw']):.2e} hz={rel(hz, gd['hazards']):.2e} S={rel(S, gd['S']):.2e} "
                  f"loss={abs(loss.item() - gd['loss'].item()):.2e} worst-grad-rel={worst:.2e}", flush=True)
        for nm, cfg in cases.RADIO_CASES.items():
            gd = gold["radio"][nm]
            model = build_radio_model(cfg).to(dev)
            bags = {k: v.to(dev) for k, v in cases.radio_bags(cfg).items()}
            Y, c = cases.labels(cfg)
            hz, S, Yh, A = model(**bags)
            loss = NLLSurvLoss(alpha=cfg["alpha"])(hazards=hz, S=S, Y=Y.to(dev), c=c.to(dev))
            model.zero_grad(); loss.backward(); torch.cuda.synchronize()
            worst = 0.0
            for k, p in model.named_parameters():
                fp = gd["grads"][k]
                idx = cases._sample_idx(p.numel())
                e = (p.grad.reshape(-1).cpu()[idx] - fp["vals"]).abs().max().item() / max(fp["vals"].abs().max().item(), 1e-30)
                if fp["vals"].abs().max().item() > 1e-6:
                    worst = max(worst, e)
            print(f"[radio {nm}] A={rel(A, gd['A_raw']):.2e} hz={rel(hz, gd['hazards']):.2e} "
                  f"loss={abs(loss.item() - gd['loss'].item()):.2e} worst-grad-rel={worst:.2e}", flush=True)
        for nm, cfg in cases.OMIC_CASES.items():
            gd = gold["omic"][nm]
            model = build_omic_model(cfg).to(dev)
            x = cases.omic_batch(cfg).to(dev).requires_grad_(True)
            times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
            risk = model(genomic_features=x)[0]
            loss = CoxSurvLoss()(risks=risk, times=times.to(dev), c=c.to(dev))
            model.zero_grad(); loss.backward(); torch.cuda.synchronize()
            print(f"[omic {nm}] risk={rel(risk, gd['risk']):.2e} loss={abs(loss.item() - gd['loss'].item()):.2e} "
                  f"dx={rel(x.grad, gd['dx']):.2e}", flush=True)
        for nm, cfg in cases.HEAD_CASES.items():
            gd = gold["heads"][nm]
            model = build_head_model(cfg).to(dev)
            hr, hp, ho = [t.to(dev).requires_grad_(True) for t in cases.embeddings(cfg)]
            times, c = cases.cohort_labels(cfg["B"], cfg["seed"])
            res = model(hr, hp, ho)
            if cfg["kind"] == "cox":
                risk = res[0]
                loss = (CoxSurvLoss()(risks=risk, times=times.to(dev), c=c.to(dev)) if cfg["loss"] == "cox"
                        else RankingSurvLoss()(risks=risk.reshape(-1), times=times.to(dev), c=c.to(dev)))
            else:
                risk, hz, S = res
                Yl = (torch.arange(cfg["B"]) % 4).to(dev)
                loss = NLLSurvLoss(alpha=0.15)(hazards=hz, S=S, Y=Yl, c=c.to(dev))
            model.zero_grad(); loss.backward(); torch.cuda.synchronize()
            dins = [t.grad for t in (hr, hp, ho)]
            derr = max(rel(d, gdd) for d, gdd in zip(dins, gd["d_inputs"]) if d is not None and gdd is not None)
            print(f"[head {nm}] risk={rel(risk, gd['risk']):.2e} loss={abs(loss.item() - gd['loss'].item()):.2e} "
                  f"d_inputs={derr:.2e}", flush=True)
    elif name == "timing":
        for (L, D) in ((512, 384), (256, 256)):
            N = 16384
            W = rand_amil(L, D, True)
            prep = ops.prepare_amil_weights(*[t.to(dev) for t in W])
            xs = [(0.5 * torch.randn(N, 1024, device=dev).abs()).to(torch.bfloat16) for _ in range(6)]
            flags = ops.amil_flags(True)
            dM = torch.randn(L, device=dev) * 0.1
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for it in range(3):
                A_raw, M, ml = ops.amil_forward(xs[it % 6], prep, flags, 0)
                ops.amil_backward(xs[it % 6], prep, flags, 0, A_raw, ml, M, dM)
            torch.cuda.synchronize()
            tf = tb = 0.0
            reps = 6
            for it in range(reps):
                ev[0].record()
                A_raw, M, ml = ops.amil_forward(xs[it % 6], prep, flags, 0)
                ev[1].record()
                ops.amil_backward(xs[it % 6], prep, flags, 0, A_raw, ml, M, dM)
                ev[2].record()
                torch.cuda.synchronize()
                tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
            tf, tb = tf / reps * 1e3, tb / reps * 1e3
            F = 2 * N * (2 * 1024 * L + 6 * L * D)
            print(f"[timing L={L} D={D} N={N}] fwd {tf:.1f} us  bwd {tb:.1f} us  total {tf + tb:.1f} us  "
                  f"-> {N / (tf + tb) * 1e6 / 1e6:.1f} M patches/s, {F / (tf + tb) / 1e6:.1f} TFLOP/s algorithmic", flush=True)


def main():
    if len(sys.argv) > 1:
        try:
            run_group(sys.argv[1])
        except Exception:
            traceback.print_exc()
            sys.exit(1)
        return
    for g in GROUPS:
        t0 = time.time()
        print(f"===== {g} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=300, capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[{g}] EXIT {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"[{g}] TIMEOUT\n{(e.stdout or b'')[-2000:]}")
        print(f"===== {g} done in {time.time() - t0:.1f}s =====", flush=True)


if __name__ == "__main__":
    main()
