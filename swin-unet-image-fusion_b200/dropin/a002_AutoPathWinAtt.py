"""Drop-in for a002_AutoPathWinAtt.py: owns window_attention_x / window_attention_y and routes
self (x<-x, y<-y) vs cross (x: Q=x, KV=y; y: Q=y, KV=x) attention (a002:58-82)."""
from torch import nn

from a001_WindowAttention import WindowAttention
from swinfuse import ops


class AutoPathWinAtt(nn.Module):
    def __init__(self, in_out_dims: int, num_heads: int, dims_per_head: int, window_size: tuple,
                 use_cyclic_shift: bool, use_dual_path: bool, use_cross_att: bool, use_qkv_bias: bool,
                 attention_drop_ratio: float, linear_after_att_drop_ratio: float):
        super().__init__()
        self.in_out_dims, self.num_heads, self.dims_per_head, self.window_size = in_out_dims, num_heads, dims_per_head, window_size
        self.use_cyclic_shift, self.use_dual_path, self.use_cross_att, self.use_qkv_bias = use_cyclic_shift, use_dual_path, use_cross_att, use_qkv_bias
        self.attention_drop_ratio, self.linear_after_att_drop_ratio = attention_drop_ratio, linear_after_att_drop_ratio

        def make():
            return WindowAttention(in_out_dims=in_out_dims, num_heads=num_heads, dims_per_head=dims_per_head,
                                   window_size=window_size, use_cyclic_shift=use_cyclic_shift,
                                   use_cross_attention=use_cross_att, use_qkv_bias=use_qkv_bias,
                                   attention_drop_ratio=attention_drop_ratio,
                                   linear_after_att_drop_ratio=linear_after_att_drop_ratio)

        self.window_attention_x = make()
        if use_dual_path:
            self.window_attention_y = make()

    def fused(self, x, y, ln_x, ln_y, eps: float = 1e-5):
        """x + Attn_x(LN_x(x), ...), y + Attn_y(LN_y(y), ...) with LN and residual inside the operator."""
        if not self.use_dual_path:
            return self.window_attention_x.fused(x, None, ln_q=ln_x, ln_kv=ln_x, residual=x, eps=eps)
        if self.use_cross_att:  # both directions read the pre-update tensors (a002:70-73)
            return ops.dual_path(lambda: self.window_attention_x.fused(x, y, ln_q=ln_x, ln_kv=ln_y, residual=x, eps=eps),
                                 lambda: self.window_attention_y.fused(y, x, ln_q=ln_y, ln_kv=ln_x, residual=y, eps=eps))
        return ops.dual_path(lambda: self.window_attention_x.fused(x, None, ln_q=ln_x, ln_kv=ln_x, residual=x, eps=eps),
                             lambda: self.window_attention_y.fused(y, None, ln_q=ln_y, ln_kv=ln_y, residual=y, eps=eps))

    def forward(self, x, y):
        if not self.use_dual_path:
            return self.window_attention_x(q=x, k=x, v=x)
        if self.use_cross_att:
            return self.window_attention_x(q=x, k=y, v=y), self.window_attention_y(q=y, k=x, v=x)
        return self.window_attention_x(q=x, k=x, v=x), self.window_attention_y(q=y, k=y, v=y)

    def forward_(self, x, y):
        return self(x, y)
