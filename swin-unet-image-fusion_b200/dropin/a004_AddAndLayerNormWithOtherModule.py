"""Drop-in for a004_AddAndLayerNormWithOtherModule.py: the pre-norm residual wrapper
``x + other(LN_C(x))`` (a004:20-48).  When ``other_module`` is one of this package's attention /
MLP modules, LayerNorm and the residual add are folded into that module's kernel call."""
import torch
from torch import nn

from a007_utils import *  # noqa: F401,F403
from swinfuse import ops


def _ln(layer: nn.LayerNorm):
    return (layer.weight, layer.bias)


class AddAndLayerNormWithOtherModule(nn.Module):
    def __init__(self, normalized_shape: list, use_dual_path: bool, other_module: nn.Module):
        super().__init__()
        self.normalized_shape, self.use_dual_path, self.other_module = normalized_shape, use_dual_path, other_module
        self.norm_layer_1 = nn.LayerNorm(normalized_shape=normalized_shape)
        if use_dual_path:
            self.norm_layer_2 = nn.LayerNorm(normalized_shape=normalized_shape)

    def forward(self, x, y):
        dual = self.use_dual_path or y is not None
        fused = getattr(self.other_module, "fused", None)
        if fused is not None:
            if dual and self.norm_layer_2.eps != self.norm_layer_1.eps:
                raise ops.SwinFuseError("AddAndLayerNormWithOtherModule: both LayerNorms must share one eps")
            return fused(x, y if dual else None, _ln(self.norm_layer_1), _ln(self.norm_layer_2) if dual else None,
                         eps=self.norm_layer_1.eps)
        # generic other_module: forward-only composition of the standalone operators (sf_layernorm has no adjoint
        # outside the fused operators): refuse to drop gradients silently
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in
                                           (x, y, self.norm_layer_1.weight, self.norm_layer_1.bias)):
            raise ops.SwinFuseError("AddAndLayerNormWithOtherModule: the non-fused composition is forward-only; wrap an "
                                    "AutoPathWinAtt / AutoPathMLP (fused operators own the gradients) or run under no_grad")
        if dual:
            nx, ny = my_layer_norm(x, self.norm_layer_1, y, self.norm_layer_2)
            ox, oy = self.other_module(nx, ny)
            return ops.add(x, ox), ops.add(y, oy)
        return ops.add(x, self.other_module(x=my_layer_norm(x, self.norm_layer_1), y=None))

    def forward_(self, x, y):
        return self(x, y)


def my_layer_norm(x, layer_x, y=None, layer_y=None):
    """LN over the channel dim of (B,C,H,W) tensors (a004:54-72); forward only."""
    nx = ops.layernorm(x, layer_x.weight, layer_x.bias, layer_x.eps)
    if y is None:
        return nx
    return nx, ops.layernorm(y, layer_y.weight, layer_y.bias, layer_y.eps)
