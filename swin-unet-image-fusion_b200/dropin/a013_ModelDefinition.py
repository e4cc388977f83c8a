"""Drop-in for a013_ModelDefinition.py: ``MyModel`` -- the dual-path (IR + visible) Swin-UNet.

Same 14 constructor kwargs, same module tree (encoder_list / decoder_list / final_layer) and
therefore the same 3,139 state_dict keys as the reference (a013:17-207).  ``forward(in_x, in_y)``
follows a013:209-230: 5 encoder stages [pad(2x2), patch merge, pad(7x7), 4 blocks], skip stack,
5 decoder stages [4 blocks, crop, anti patch merge, crop] with ``x += skip`` fused into the
preceding crop kernel, then the conv head (sf_head_fwd).  Every arithmetic op runs in
libswinfuse.so; feature maps stay in NHWC memory between operators."""
import math
from collections import deque

import torch
from torch import nn

from a003_AutoPathMLP import check_elu
from a006_PaddingOperation import MyPadding
from a010_StateRecorder import StateRecorder
from a011_PatchOperation import PatchMergingAndLinearLayer
from a012_SelfAndCrossBlockPair import SelfAndCrossBlockPair
from a009_NormalAndShiftWinsBlockPair import NormalAndShiftWinsBlockPair  # noqa: F401
from swinfuse import ops


class MyModel(nn.Module):
    def __init__(self, window_size: tuple, merging_size: tuple, in_dims_list: list, out_dims_list: list,
                 att_num_heads: int, att_dims_per_head_ratio: float, attention_drop_ratio: float,
                 linear_after_att_drop_ratio: float, mlp_hidden_dims_ratio: int, mlp_activation_func: nn.Module,
                 mlp_drop_ratio: float, final_layer_att_dims_per_head_ratio: float, final_conv_layer_kernel_size: int,
                 final_layer_mlp_hidden_dims_ratio: int):
        super().__init__()
        check_elu(mlp_activation_func, "MyModel")
        self.window_size, self.merging_size = window_size, merging_size
        self.in_dims_list, self.out_dims_list = in_dims_list, out_dims_list
        self.att_num_heads, self.att_dims_per_head_ratio = att_num_heads, att_dims_per_head_ratio
        self.attention_drop_ratio, self.linear_after_att_drop_ratio = attention_drop_ratio, linear_after_att_drop_ratio
        self.mlp_hidden_dims_ratio, self.mlp_activation_func, self.mlp_drop_ratio = mlp_hidden_dims_ratio, mlp_activation_func, mlp_drop_ratio
        self.final_layer_att_dims_per_head_ratio = final_layer_att_dims_per_head_ratio
        self.final_layer_conv_kernel_size = final_conv_layer_kernel_size
        self.final_layer_mlp_hidden_dims_ratio = final_layer_mlp_hidden_dims_ratio
        self.feature_shape_recorder, self.padding_size_recorder = StateRecorder(), StateRecorder()
        self.patch_merging_size_recorder = StateRecorder()
        self.u_net_intermediate_result_recorder = StateRecorder()
        self.encoder_list, self.decoder_list = self.generate_model_using_deque()
        self.final_layer = self.get_final_layer()

    def get_final_layer(self):
        k = self.final_layer_conv_kernel_size
        return nn.Sequential(
            nn.Conv2d(2, 2, kernel_size=k, padding="same", padding_mode="reflect"),
            nn.BatchNorm2d(2),
            self.mlp_activation_func,
            nn.Conv2d(2, 1, kernel_size=k, padding="same", padding_mode="reflect"))

    def do_final_layer(self, x, y):
        conv1, bn, _, conv2 = self.final_layer
        use_batch_stats = bn.training or bn.running_mean is None   # nn.BatchNorm2d decides by its OWN mode flag
        if bn.momentum is None or not bn.track_running_stats or not bn.affine:
            raise ops.SwinFuseError("MyModel: the head expects the default nn.BatchNorm2d(2) (a013:133)")
        out = ops.final_head(x, y, w1=conv1.weight, b1=conv1.bias, bn_gamma=bn.weight, bn_beta=bn.bias,
                             running_mean=bn.running_mean, running_var=bn.running_var, w2=conv2.weight, b2=conv2.bias,
                             training=use_batch_stats, eps=bn.eps, momentum=bn.momentum)
        if use_batch_stats:
            bn.num_batches_tracked += 1  # bookkeeping counter of nn.BatchNorm2d
        return out

    def generate_model_using_deque(self):
        enc, dec = deque(), deque()
        for j in range(len(self.in_dims_list) - 1, -1, -1):
            common = dict(window_size=self.window_size, feature_shape_recorder=self.feature_shape_recorder,
                          padding_size_recorder=self.padding_size_recorder, merging_size=self.merging_size,
                          patch_merging_size_recorder=self.patch_merging_size_recorder,
                          att_num_heads=self.att_num_heads,
                          att_dims_per_head=math.floor(self.out_dims_list[j] * self.att_dims_per_head_ratio),
                          attention_drop_ratio=self.attention_drop_ratio,
                          linear_after_att_drop_ratio=self.linear_after_att_drop_ratio,
                          mlp_activation_func=self.mlp_activation_func, mlp_drop_ratio=self.mlp_drop_ratio)
            enc.appendleft(get_encoder_or_decoder_block(
                mode="encoder", in_dims=self.in_dims_list[j], out_dims=self.out_dims_list[j],
                mlp_hidden_dims=self.out_dims_list[j] * self.mlp_hidden_dims_ratio, **common))
            dec.append(get_encoder_or_decoder_block(
                mode="decoder", in_dims=self.out_dims_list[j], out_dims=self.in_dims_list[j],
                mlp_hidden_dims=self.in_dims_list[j] * self.mlp_hidden_dims_ratio, **common))
        return nn.ModuleList(enc), nn.ModuleList(dec)

    def forward(self, in_x, in_y) -> torch.Tensor:
        skips = self.u_net_intermediate_result_recorder
        skips.delete_all()
        self.feature_shape_recorder.delete_all()
        self.padding_size_recorder.delete_all()
        x, y = ops.as_fmap(in_x, "MyModel.in_x"), ops.as_fmap(in_y, "MyModel.in_y")
        last = len(self.encoder_list) - 1
        for i, stage in enumerate(self.encoder_list):
            for m in stage:
                x, y = m(x=x, y=y)
            if i < last:
                skips.record((x, y))
        for j, stage in enumerate(self.decoder_list):
            blocks, crop_win, unmerge, crop_merge = stage
            x, y = blocks(x=x, y=y)
            x, y = crop_win(x=x, y=y)
            x, y = unmerge(x=x, y=y)
            # the next stage starts with `x += skip` (a013:222-225): fold it into this crop
            skip = skips.read() if j < len(self.decoder_list) - 1 else None
            x, y = crop_merge(x, y, skip=skip)
        return self.do_final_layer(x, y)

    def forward_(self, in_x, in_y) -> torch.Tensor:
        return self(in_x, in_y)


def get_encoder_or_decoder_block(mode: str, window_size: tuple, feature_shape_recorder: StateRecorder,
                                 padding_size_recorder: StateRecorder, merging_size: tuple, in_dims: int,
                                 out_dims: int, patch_merging_size_recorder: StateRecorder, att_num_heads: int,
                                 att_dims_per_head: int, attention_drop_ratio: float,
                                 linear_after_att_drop_ratio: float, mlp_hidden_dims: int,
                                 mlp_activation_func: nn.Module, mlp_drop_ratio: float) -> nn.ModuleList:
    if mode not in ("encoder", "decoder"):
        raise ValueError("mode must be either encoder or decoder")
    enc = mode == "encoder"
    rec = dict(feature_shape_recorder=feature_shape_recorder, padding_size_recorder=padding_size_recorder)
    stage = nn.ModuleList([
        MyPadding(belongs_to_encoder=enc, window_size=merging_size, use_dual_path=True, **rec),
        PatchMergingAndLinearLayer(belongs_to_encoder=enc, use_dual_path=True, merging_or_unmerging_size=merging_size,
                                   in_dims=in_dims, out_dims=out_dims,
                                   patch_merging_size_recorder=patch_merging_size_recorder,
                                   activation_func=mlp_activation_func),
        MyPadding(belongs_to_encoder=enc, window_size=window_size, use_dual_path=True, **rec),
        SelfAndCrossBlockPair(in_out_dims=out_dims if enc else in_dims, num_heads=att_num_heads,
                              dims_per_head=att_dims_per_head, window_size=window_size, use_dual_path=True,
                              use_qkv_bias=True, attention_drop_ratio=attention_drop_ratio,
                              linear_after_att_drop_ratio=linear_after_att_drop_ratio,
                              mlp_hidden_dims=mlp_hidden_dims, mlp_activation_func=mlp_activation_func,
                              mlp_drop_ratio=mlp_drop_ratio)])
    return stage if enc else stage[::-1]
