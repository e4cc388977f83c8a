"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch
from torch import nn

from oracle import fusion_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances stated by BASELINE.json north_star
TOL_FP32 = 1e-4   # fp32 forward: relative
TOL_BF16 = 2e-2   # bf16 tensor-core path: relative


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max(|got-ref|) / max(|ref|): the relative error the tolerances are quoted in."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def dropin():
    import swinfuse
    swinfuse.install_dropin()
    return swinfuse


def build_model(cfg: fo.FusionConfig = None, device="cuda", act=None):
    """The drop-in MyModel, constructed with the reference's 14 kwargs (a016:26-40)."""
    dropin()
    from a013_ModelDefinition import MyModel
    cfg = cfg or fo.FusionConfig()
    m = MyModel(window_size=cfg.window_size, merging_size=cfg.merging_size, in_dims_list=cfg.in_dims_list,
                out_dims_list=cfg.out_dims_list, att_num_heads=cfg.att_num_heads,
                att_dims_per_head_ratio=cfg.att_dims_per_head_ratio, attention_drop_ratio=0,
                linear_after_att_drop_ratio=0, mlp_hidden_dims_ratio=cfg.mlp_hidden_dims_ratio,
                mlp_activation_func=act if act is not None else nn.ELU(inplace=True), mlp_drop_ratio=0,
                final_layer_att_dims_per_head_ratio=1, final_conv_layer_kernel_size=cfg.final_conv_layer_kernel_size,
                final_layer_mlp_hidden_dims_ratio=1)
    return m.to(device) if device else m


def wa_cases():
    g = golden("window_attention.npz")
    return g, json.loads(bytes(g["cases"]).decode())
