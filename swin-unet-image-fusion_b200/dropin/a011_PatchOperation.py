"""Drop-in for a011_PatchOperation.py: ``PatchMergingAndLinearLayer``.

encoder: 2x2 space-to-depth -> 1x1 conv -> LayerNorm -> ELU                  (a011:87-93, 236-239)
decoder: 1x1 conv -> LayerNorm(4*C_out) -> depth-to-space -> ELU ("anti patch merging", a011:111-117, 241)

One sf_patch_fwd call per path; parameters stay in nn.Conv2d / nn.LayerNorm members named as in
the reference (mlp_layer_{x,y}, layer_norm_{x,y}, buffer_to_show_device)."""
import torch
from torch import nn

from a003_AutoPathMLP import check_elu
from a004_AddAndLayerNormWithOtherModule import my_layer_norm  # noqa: F401  (re-exported like the reference)
from a007_utils import *  # noqa: F401,F403
from a010_StateRecorder import StateRecorder
from swinfuse import ops


class PatchMergingAndLinearLayer(nn.Module):
    def __init__(self, belongs_to_encoder: bool, use_dual_path: bool, in_dims: int, out_dims: int,
                 patch_merging_size_recorder: StateRecorder, merging_or_unmerging_size: tuple,
                 activation_func: nn.Module = None):
        super().__init__()
        activation_func = activation_func if activation_func is not None else nn.ELU()
        check_elu(activation_func, "PatchMergingAndLinearLayer")
        self.belongs_to_encoder, self.use_dual_path = belongs_to_encoder, use_dual_path
        self.in_dims, self.out_dims = in_dims, out_dims
        self.patch_merging_recorder = patch_merging_size_recorder
        self.merging_or_unmerging_size = merging_or_unmerging_size
        self.activation_func = activation_func
        mh, mw = merging_or_unmerging_size
        if belongs_to_encoder:
            self.conv_in_dims, self.conv_out_dims = in_dims * mh * mw, out_dims
        else:
            self.conv_in_dims, self.conv_out_dims = in_dims, out_dims * mh * mw
        for path in ("x", "y") if use_dual_path else ("x",):
            setattr(self, f"mlp_layer_{path}", nn.Conv2d(self.conv_in_dims, self.conv_out_dims, kernel_size=1))
            setattr(self, f"layer_norm_{path}", nn.LayerNorm(normalized_shape=self.conv_out_dims))
        self.register_buffer(name="buffer_to_show_device", tensor=torch.zeros(size=(1,)))
        self.precision = None

    def _one(self, path: str, t):
        conv, ln = getattr(self, f"mlp_layer_{path}"), getattr(self, f"layer_norm_{path}")
        return ops.patch_layer(t, w=conv.weight, b=conv.bias, ln_gamma=ln.weight, ln_beta=ln.bias,
                               encoder=self.belongs_to_encoder, merging_size=self.merging_or_unmerging_size,
                               out_dims=self.out_dims, eps=ln.eps, precision=self.precision)

    # index-only helpers of the reference API
    def do_patch_merging_for_one_tensor(self, feature):
        return ops.patch_merge(feature, self.merging_or_unmerging_size)

    def undo_patch_merging_for_one_tensor(self, feature):
        return ops.patch_unmerge(feature, self.merging_or_unmerging_size)

    def forward(self, x, y=None):
        if y is not None:
            return ops.dual_path(lambda: self._one("x", x), lambda: self._one("y", y))
        return self._one("x", x)

    def forward_(self, x, y):
        return self(x, y)
