"""Import the REAL reference modules from /root/reference  --  TEST INFRASTRUCTURE.

Only usable in the authoring container (``/root/reference`` does not exist on the GPU box).
Used by ``oracle/make_golden.py`` to generate the committed fixtures and by the CPU tests
that pin the oracle (skipped when the reference tree is absent).

``a013_ModelDefinition`` imports ``a008_loss`` which imports ``kornia`` (absent): a stub
``kornia`` package backed by ``oracle/kornia_restatement.py`` is injected so the model file
imports.  Nothing is written to /root/reference.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SWINFUSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "a013_ModelDefinition.py"))


def _install_kornia_stub() -> None:
    if "kornia" in sys.modules:
        return
    from oracle import kornia_restatement as kr

    kornia = types.ModuleType("kornia")
    losses = types.ModuleType("kornia.losses")
    filters = types.ModuleType("kornia.filters")
    losses.ssim_loss, losses.MS_SSIMLoss, losses.PSNRLoss = kr.ssim_loss, kr.MS_SSIMLoss, kr.PSNRLoss
    filters.Canny, filters.Sobel = kr.Canny, kr.Sobel
    kornia.losses, kornia.filters = losses, filters
    sys.modules["kornia"], sys.modules["kornia.losses"], sys.modules["kornia.filters"] = kornia, losses, filters


_REF_MODULES = ["A000_CONFIG", "a001_WindowAttention", "a002_AutoPathWinAtt", "a003_AutoPathMLP",
                "a004_AddAndLayerNormWithOtherModule", "a005_BasicBlock", "a006_PaddingOperation", "a007_utils",
                "a008_loss", "a009_NormalAndShiftWinsBlockPair", "a010_StateRecorder", "a011_PatchOperation",
                "a012_SelfAndCrossBlockPair", "a013_ModelDefinition"]


class ReferenceModules:
    """Context manager: puts /root/reference first on sys.path, imports the reference's flat
    modules under their own names, and on exit removes them from sys.modules again so the
    drop-in modules of the same names can be imported afterwards."""

    def __enter__(self):
        if not reference_available():
            raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
        _install_kornia_stub()
        self._saved = {m: sys.modules.pop(m) for m in _REF_MODULES if m in sys.modules}
        sys.path.insert(0, REFERENCE_ROOT)
        self.mods = {m: importlib.import_module(m) for m in _REF_MODULES}
        return self

    def __getattr__(self, name):
        return self.__dict__["mods"][name]

    def __exit__(self, *exc):
        sys.path.remove(REFERENCE_ROOT)
        for m in _REF_MODULES:
            sys.modules.pop(m, None)
        sys.modules.update(self._saved)
        return False


def build_reference_model(ref: ReferenceModules, cfg=None, act=None):
    """MyModel built the way a016:26-40 does, with ``nn.ELU()`` (the default in-place ELU
    cannot back-propagate on torch 2.11 -- SURVEY.md appendix D.1; forward is identical)."""
    from torch import nn
    from oracle.fusion_oracle import FusionConfig

    cfg = cfg or FusionConfig()
    return ref.a013_ModelDefinition.MyModel(
        window_size=cfg.window_size, merging_size=cfg.merging_size, in_dims_list=cfg.in_dims_list,
        out_dims_list=cfg.out_dims_list, att_num_heads=cfg.att_num_heads,
        att_dims_per_head_ratio=cfg.att_dims_per_head_ratio, attention_drop_ratio=0.0,
        linear_after_att_drop_ratio=0.0, mlp_hidden_dims_ratio=cfg.mlp_hidden_dims_ratio,
        mlp_activation_func=act if act is not None else nn.ELU(), mlp_drop_ratio=0.0,
        final_layer_att_dims_per_head_ratio=1, final_conv_layer_kernel_size=cfg.final_conv_layer_kernel_size,
        final_layer_mlp_hidden_dims_ratio=1)
