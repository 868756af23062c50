"""Shared test helpers: build the drop-in models for a seeded case and pull their weights out in the
layout the oracle functions take."""
import torch

from multimodalfusion_b200.models import (MIL_Attention_fc_surv_path, MIL_Attention_fc_surv_radio, MaxNet)
from multimodalfusion_b200.models import coxranking_models_pretrained as cox_heads
from multimodalfusion_b200.models import nll_models_pretrained as nll_heads
from oracle import cases


def build_path_model(cfg):
    torch.manual_seed(cfg["seed"])
    model = MIL_Attention_fc_surv_path(gate_path=cfg["gated"], model_size_wsi=cfg["size"],
                                       dropout=cfg["dropout"], n_classes=cfg["K"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    if cfg.get("peaky"):
        cases.make_peaky(model, cfg["peaky"])
    return model


def build_radio_model(cfg):
    torch.manual_seed(cfg["seed"])
    model = MIL_Attention_fc_surv_radio(gate_radio=cfg["gated"], dropout=cfg["dropout"], n_classes=cfg["K"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def build_radio_tensor_model(cfg):
    torch.manual_seed(cfg["seed"])
    model = MIL_Attention_fc_surv_radio(radio_fusion="tensor", gate_radio=cfg["gated"], dropout=cfg["dropout"],
                                        n_classes=cfg["K"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def build_mm_model(cfg):
    """MM_MIL_Attention_fc_surv seeded and perturbed like oracle/make_goldens_mm.py builds the reference's."""
    from multimodalfusion_b200.models.model_mm_attention_mil import MM_MIL_Attention_fc_surv
    torch.manual_seed(cfg["seed"])
    model = MM_MIL_Attention_fc_surv(input_dim=cfg["d"], radio_fusion="concat", fusion=cfg["fusion"], gate=True,
                                     gate_path=True, gate_omic=True, gate_radio=True, model_size_radio="small",
                                     model_size_wsi="small", model_size_omic="small", dropout=False, n_classes=4,
                                     mode=cfg["mode"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def build_omic_model(cfg):
    torch.manual_seed(cfg["seed"])
    model = MaxNet(cfg["d_in"], bag_loss=cfg["bag_loss"], n_classes=4).eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def build_head_model(cfg):
    torch.manual_seed(cfg["seed"])
    mod = cox_heads if cfg["kind"] == "cox" else nll_heads
    model = mod.multimodal_pretrained(mode=cfg["mode"], train_type="kronecker", n_classes=4).eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def build_head2_model(cfg):
    """fcnn / Highway fusion head for a HEAD2 case: seeded like the reference, biases and BatchNorm state perturbed."""
    torch.manual_seed(cfg["seed"])
    mod = cox_heads if cfg["kind"] == "cox" else nll_heads
    model = mod.multimodal_pretrained(mode=cfg["mode"], train_type=cfg["train_type"], n_classes=4,
                                      n_layers=cfg["n_layers"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    cases.perturb_bn(model, cfg["seed"])
    return model


def build_unimodal_model(cfg):
    """unimonal_pretrained head for a UNI case: seeded like the reference, biases and BatchNorm state perturbed."""
    torch.manual_seed(cfg["seed"])
    mod = cox_heads if cfg["kind"] == "cox" else nll_heads
    model = mod.unimonal_pretrained(mode=cfg["mode"], train_type=cfg["train_type"], n_classes=4,
                                    n_layers=cfg["n_layers"]).eval()
    cases.perturb_biases(model, cfg["seed"])
    cases.perturb_bn(model, cfg["seed"])
    return model


def unimodal_input(cfg):
    hr, hp, ho = cases.embeddings(cfg)
    return {"radio": hr, "path": hp, "omic": ho}[cfg["mode"]]


def build_xfusion4(cfg):
    """XlinearFusion with the reference's default ctor (4 modalities), seeded and perturbed like the golden generator."""
    from multimodalfusion_b200.models.model_modules import XlinearFusion
    torch.manual_seed(cfg["seed"])
    model = XlinearFusion().eval()
    cases.perturb_biases(model, cfg["seed"])
    return model


def amil_weights(seq):
    """(W1, b1, Wa, ba, Wb, bb, wc, bc) as detached fp32 tensors from an attention_net_* Sequential."""
    fc, attn = seq[0], seq[3]
    Wa, ba, Wb, bb, wc, bc = attn.amil_weights()
    out = [fc.weight, fc.bias, Wa, ba, Wb, bb, wc, bc]
    return [None if t is None else t.detach().clone().float() for t in out]


def rel_err(a, b):
    """max|a-b| / max|b| (the per-tensor relative error used by the parity bars)."""
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def xfusion_params(xf):
    red = []
    for blk in xf.reduce:
        red.append(((blk[0][0].weight.detach(), blk[0][0].bias.detach()),
                    (blk[1][0].weight.detach(), blk[1][0].bias.detach()),
                    (blk[2][0].weight.detach(), blk[2][0].bias.detach())))
    e1 = (xf.encoder1[0].weight.detach(), xf.encoder1[0].bias.detach())
    e2 = (xf.encoder2[0].weight.detach(), xf.encoder2[0].bias.detach())
    return red, e1, e2
