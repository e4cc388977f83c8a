"""Drop-in for a009_NormalAndShiftWinsBlockPair.py: a regular-window BasicBlock followed by a
shifted-window BasicBlock (a009:57-109)."""
from torch import nn

from a005_BasicBlock import BasicBlock


class NormalAndShiftWinsBlockPair(nn.Module):
    def __init__(self, in_out_dims: int, num_heads: int, dims_per_head: int, window_size: tuple, use_dual_path: bool,
                 use_cross_attr: bool, use_qkv_bias: bool, attention_drop_ratio: float,
                 linear_after_att_drop_ratio: float, mlp_hidden_dims: int, mlp_activation_func: nn.Module,
                 mlp_drop_ratio: float):
        super().__init__()
        kw = dict(in_out_dims=in_out_dims, num_heads=num_heads, dims_per_head=dims_per_head, window_size=window_size,
                  use_dual_path=use_dual_path, use_cross_attr=use_cross_attr, use_qkv_bias=use_qkv_bias,
                  attention_drop_ratio=attention_drop_ratio, linear_after_att_drop_ratio=linear_after_att_drop_ratio,
                  mlp_hidden_dims=mlp_hidden_dims, mlp_activation_func=mlp_activation_func,
                  mlp_drop_ratio=mlp_drop_ratio)
        for k, v in kw.items():
            setattr(self, k, v)
        self.normal_window_block = BasicBlock(use_cyclic_shift=False, **kw)
        self.shifted_window_block = BasicBlock(use_cyclic_shift=True, **kw)

    def forward(self, x, y=None):
        if self.use_dual_path:
            x, y = self.normal_window_block(x=x, y=y)
            return self.shifted_window_block(x=x, y=y)
        return self.shifted_window_block(x=self.normal_window_block(x=x, y=None), y=None)

    def forward_(self, x, y):
        return self(x, y)
