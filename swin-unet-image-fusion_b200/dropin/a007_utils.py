"""Drop-in for a007_utils.py (a007:7-26): NCHW <-> NHWC *views*.  The kernels work on NHWC
memory directly, so these are only kept for callers that import them."""
from torch import Tensor


def put_channel_dim_to_the_last_position(tensor: Tensor) -> Tensor:
    return tensor.permute(0, 2, 3, 1)


def put_channel_dim_to_the_second_position(tensor: Tensor) -> Tensor:
    return tensor.permute(0, 3, 1, 2)
