"""Drop-in for a012_SelfAndCrossBlockPair.py: a self-attention block pair followed by a
cross-attention block pair -- four BasicBlocks per stage, in the fixed order self/normal,
self/shifted, cross/normal, cross/shifted (a012:40-78)."""
from torch import nn

from a009_NormalAndShiftWinsBlockPair import NormalAndShiftWinsBlockPair


class SelfAndCrossBlockPair(nn.Module):
    def __init__(self, in_out_dims: int, num_heads: int, dims_per_head: int, window_size: tuple, use_dual_path: bool,
                 use_qkv_bias: bool, attention_drop_ratio: float, linear_after_att_drop_ratio: float,
                 mlp_hidden_dims: int, mlp_activation_func: nn.Module, mlp_drop_ratio: float):
        super().__init__()
        kw = dict(in_out_dims=in_out_dims, num_heads=num_heads, dims_per_head=dims_per_head, window_size=window_size,
                  use_dual_path=use_dual_path, use_qkv_bias=use_qkv_bias, attention_drop_ratio=attention_drop_ratio,
                  linear_after_att_drop_ratio=linear_after_att_drop_ratio, mlp_hidden_dims=mlp_hidden_dims,
                  mlp_activation_func=mlp_activation_func, mlp_drop_ratio=mlp_drop_ratio)
        for k, v in kw.items():
            setattr(self, k, v)
        self.self_att_block = NormalAndShiftWinsBlockPair(use_cross_attr=False, **kw)
        self.cross_att_block = NormalAndShiftWinsBlockPair(use_cross_attr=True, **kw)

    def forward(self, x, y=None):
        if self.use_dual_path:
            x, y = self.self_att_block(x=x, y=y)
            return self.cross_att_block(x=x, y=y)
        return self.cross_att_block(x=self.self_att_block(x=x, y=None), y=None)

    def forward_(self, x, y):
        return self(x, y)
