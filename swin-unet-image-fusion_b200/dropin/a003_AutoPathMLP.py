"""Drop-in for a003_AutoPathMLP.py: per-path MLP of two 1x1 convolutions with ELU between
(a003:21-31).  Parameters live in the same nn.Conv2d / nn.Sequential members as the reference so
that state_dict keys (mlp_x_1, sequence_x.0, ...) and init_params (a016:382-390) line up; the
forward pass is one fused sf_mlp_fwd call per path."""
from torch import nn

from a007_utils import *  # noqa: F401,F403  (the reference re-exports these)
from swinfuse import ops
from swinfuse._lib import SwinFuseError


def check_elu(act: nn.Module, who: str) -> None:
    if not isinstance(act, nn.ELU) or float(act.alpha) != 1.0:
        raise SwinFuseError(f"{who}: only nn.ELU(alpha=1) is implemented by the fused kernels "
                            f"(A000_CONFIG.py:64), got {act!r}")


class AutoPathMLP(nn.Module):
    def __init__(self, in_out_dims: int, hidden_dims: int, activation_func: nn.Module, use_dual_path: bool,
                 drop_ratio: float):
        super().__init__()
        check_elu(activation_func, "AutoPathMLP")
        self.in_out_dims, self.hidden_dims, self.activation_func = in_out_dims, hidden_dims, activation_func
        self.use_dual_path, self.drop_ratio = use_dual_path, drop_ratio
        for path in ("x", "y") if use_dual_path else ("x",):
            fc1 = nn.Conv2d(in_out_dims, hidden_dims, kernel_size=1)
            fc2 = nn.Conv2d(hidden_dims, in_out_dims, kernel_size=1)
            d1, d2 = nn.Dropout(p=drop_ratio), nn.Dropout(p=drop_ratio)
            setattr(self, f"mlp_{path}_1", fc1)
            setattr(self, f"mlp_{path}_2", fc2)
            setattr(self, f"dropout_{path}_1", d1)
            setattr(self, f"dropout_{path}_2", d2)
            setattr(self, f"sequence_{path}", nn.Sequential(fc1, activation_func, d1, fc2, d2))
        self.precision = None

    def fused_path(self, path: str, t, ln=None, residual=None, eps: float = 1e-5):
        if self.training and self.drop_ratio > 0:
            raise SwinFuseError("AutoPathMLP: non-zero dropout is not supported (A000_CONFIG.py:65 uses 0)")
        fc1, fc2 = getattr(self, f"mlp_{path}_1"), getattr(self, f"mlp_{path}_2")
        return ops.mlp(t, w1=fc1.weight, b1=fc1.bias, w2=fc2.weight, b2=fc2.bias, ln=ln, residual=residual,
                       eps=eps, precision=self.precision)

    def fused(self, x, y, ln_x, ln_y, eps: float = 1e-5):
        if self.use_dual_path or y is not None:
            return ops.dual_path(lambda: self.fused_path("x", x, ln_x, x, eps), lambda: self.fused_path("y", y, ln_y, y, eps))
        return self.fused_path("x", x, ln_x, x, eps)

    def forward(self, x, y):
        if self.use_dual_path or y is not None:
            return self.fused_path("x", x), self.fused_path("y", y)
        return self.fused_path("x", x)

    def forward_(self, x, y):
        return self(x, y)
