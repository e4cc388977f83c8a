"""Drop-in for a001_WindowAttention.py: same class, constructor, parameters and state_dict keys;
``forward`` runs the fused sm_100a window-attention operator (sf_window_attn_fwd / _bwd).

Reference semantics kept (a001:448-474): optional cyclic shift by window//2, window partition,
per-head QK^T scaled by d^-1/2 AFTER the product, + relative-position bias (key minus query, one
table shared by all heads), shift mask overwriting with -1e10, softmax, PV, output projection,
window reverse, un-shift.  Shift / partition / head split / reverse are index math inside the
kernels -- no rolled or partitioned copies are materialised.
"""
import torch
from torch import nn

from swinfuse import ops
from swinfuse._lib import SwinFuseError


def _same_tensor(a, b) -> bool:
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride())


class WindowAttention(nn.Module):
    def __init__(self, in_out_dims: int, num_heads: int, dims_per_head: int, window_size: tuple,
                 use_cyclic_shift: bool, use_cross_attention: bool, use_qkv_bias: bool, attention_drop_ratio: float,
                 linear_after_att_drop_ratio: float):
        super().__init__()
        self.in_out_dims, self.num_heads, self.dims_per_head = in_out_dims, num_heads, dims_per_head
        self.window_size = tuple(window_size)
        self.use_cyclic_shift, self.use_cross_attention, self.use_qkv_bias = use_cyclic_shift, use_cross_attention, use_qkv_bias
        self.attention_drop_ratio, self.linear_after_att_drop_ratio = attention_drop_ratio, linear_after_att_drop_ratio
        self.qk_scale = dims_per_head ** -0.5
        self.attention_drop_layer = nn.Dropout(attention_drop_ratio)
        self.linear_drop_layer = nn.Dropout(linear_after_att_drop_ratio)
        self.feature_shape_hw: tuple = tuple()
        inner = num_heads * dims_per_head
        # registration order matters for state_dict key order (a001:42-82)
        self.q_for_heads = nn.Linear(in_out_dims, inner, bias=use_qkv_bias)
        self.k_for_heads = nn.Linear(in_out_dims, inner, bias=use_qkv_bias)
        self.v_for_heads = nn.Linear(in_out_dims, inner, bias=use_qkv_bias)
        self.linear_projection = nn.Linear(inner, in_out_dims)
        self.relative_position_bias_indices = self.get_initial_relative_position_indices()
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.randn(2 * wh - 1, 2 * ww - 1))
        self.mask_for_cyclic_shift = torch.tensor([])
        self.precision = None  # None -> swinfuse default

    # ---- reference helper API (host side; the kernels recompute these as index math) ----------
    def get_initial_relative_position_indices(self):
        wh, ww = self.window_size
        t = torch.arange(wh * ww)
        coords = torch.stack([t // ww, t % ww])               # (2, t)
        rel = coords[:, None, :] - coords[:, :, None]          # [.., query i, key j] = key - query
        rel[0] += wh - 1
        rel[1] += ww - 1
        return rel

    def get_new_relative_position_bias(self):
        return ops.relative_position_bias(self.relative_position_bias_table, self.window_size)

    def initialize_feature_shape_hw(self, q):
        self.feature_shape_hw = tuple(q.shape[-2:])  # derived every call

    def initialize_mask_for_cyclic_shift(self):
        h, w = self.feature_shape_hw
        self.mask_for_cyclic_shift = ops.shift_mask(h, w, self.window_size, self.relative_position_bias_table.device)

    def _check_dropout(self):
        if self.training and (self.attention_drop_ratio > 0 or self.linear_after_att_drop_ratio > 0):
            raise SwinFuseError("WindowAttention: non-zero dropout is not supported by the fused kernel "
                                "(the reference config uses 0, A000_CONFIG.py:61-62)")

    def fused(self, q_src, kv_src, ln_q=None, ln_kv=None, residual=None, eps: float = 1e-5):
        """LN -> attention -> (+ residual) in one operator call; used by BasicBlock."""
        self._check_dropout()
        self.initialize_feature_shape_hw(q_src)
        return ops.window_attention(
            q_src, kv_src,
            wq=self.q_for_heads.weight, bq=self.q_for_heads.bias, wk=self.k_for_heads.weight, bk=self.k_for_heads.bias,
            wv=self.v_for_heads.weight, bv=self.v_for_heads.bias, wo=self.linear_projection.weight,
            bo=self.linear_projection.bias, bias_table=self.relative_position_bias_table, num_heads=self.num_heads,
            head_dim=self.dims_per_head, window_size=self.window_size, shift=self.use_cyclic_shift, ln_q=ln_q,
            ln_kv=ln_kv, residual=residual, eps=eps, precision=self.precision)

    def forward(self, q, k, v):
        if not _same_tensor(k, v):
            raise SwinFuseError("WindowAttention: k and v must be the same tensor (they always are in the "
                                "reference, a002:68-79); distinct k/v sources are not supported")
        return self.fused(q, None if _same_tensor(q, k) else k)

    def forward_(self, q, k, v):
        return self(q, k, v)
