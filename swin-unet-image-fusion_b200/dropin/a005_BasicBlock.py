"""Drop-in for a005_BasicBlock.py: one dual-path Swin block = attention residual (stage_1) then
MLP residual (stage_2) (a005:127-145).  Sub-modules are registered under the same two parents as
the reference (auto_path_win_att + stage_1.other_module, ...) so state_dict keys match."""
import torch
from torch import nn

from a002_AutoPathWinAtt import AutoPathWinAtt
from a003_AutoPathMLP import AutoPathMLP
from a004_AddAndLayerNormWithOtherModule import AddAndLayerNormWithOtherModule


class BasicBlock(nn.Module):
    def __init__(self, in_out_dims: int, num_heads: int, dims_per_head: int, window_size: tuple,
                 use_cyclic_shift: bool, use_dual_path: bool, use_cross_attr: bool, use_qkv_bias: bool,
                 attention_drop_ratio: float, linear_after_att_drop_ratio: float, mlp_hidden_dims: int,
                 mlp_activation_func: nn.Module, mlp_drop_ratio: float):
        super().__init__()
        self.in_out_dims, self.num_heads, self.dims_per_head, self.window_size = in_out_dims, num_heads, dims_per_head, window_size
        self.use_cyclic_shift, self.use_dual_path, self.use_cross_attr, self.use_qkv_bias = use_cyclic_shift, use_dual_path, use_cross_attr, use_qkv_bias
        self.attention_drop_ratio, self.linear_after_att_drop_ratio = attention_drop_ratio, linear_after_att_drop_ratio
        self.mlp_hidden_dims, self.mlp_activation_func, self.mlp_drop_ratio = mlp_hidden_dims, mlp_activation_func, mlp_drop_ratio
        self.input_compatibility_with_cross_option = None
        self.auto_path_win_att = AutoPathWinAtt(
            in_out_dims=in_out_dims, num_heads=num_heads, dims_per_head=dims_per_head, window_size=window_size,
            use_cyclic_shift=use_cyclic_shift, use_dual_path=use_dual_path, use_cross_att=use_cross_attr,
            use_qkv_bias=use_qkv_bias, attention_drop_ratio=attention_drop_ratio,
            linear_after_att_drop_ratio=linear_after_att_drop_ratio)
        self.auto_path_mlp = AutoPathMLP(in_out_dims=in_out_dims, hidden_dims=mlp_hidden_dims,
                                         activation_func=mlp_activation_func, use_dual_path=use_dual_path,
                                         drop_ratio=mlp_drop_ratio)
        self.stage_1 = AddAndLayerNormWithOtherModule([in_out_dims], use_dual_path, self.auto_path_win_att)
        self.stage_2 = AddAndLayerNormWithOtherModule([in_out_dims], use_dual_path, self.auto_path_mlp)

    def check_compatibility_between_cross_and_path_option(self):
        if self.use_cross_attr:
            assert self.use_dual_path is True

    def check_input_compatibility_with_option(self, x, y):
        """First call only (a005:89-125).  The reference prints and exit()s; here it raises."""
        if self.input_compatibility_with_cross_option is not None:
            return
        if x is None:
            raise ValueError("BasicBlock: x must not be None")
        if self.use_dual_path != (y is not None):
            raise ValueError("BasicBlock: inputs do not match the use_dual_path option")
        if self.use_cross_attr and y is not None and not torch.cuda.is_current_stream_capturing() and torch.equal(x, y):
            raise ValueError("BasicBlock: cross attention needs two different modalities (x == y everywhere)")
        self.input_compatibility_with_cross_option = True

    def forward(self, x, y=None):
        self.check_input_compatibility_with_option(x=x, y=y)
        if self.use_dual_path or y is not None:
            x, y = self.stage_1(x, y)
            return self.stage_2(x, y)
        return self.stage_2(x=self.stage_1(x=x, y=None), y=None)

    def forward_(self, x, y):
        return self(x, y)
