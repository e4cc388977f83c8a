"""Drop-in for a006_PaddingOperation.py: encoder side reflect-pads bottom/right up to a multiple
of ``window_size`` (sf_pad_reflect), decoder side pops the recorded size and crops it back
(sf_crop).  Sizes travel through the shared LIFO recorders exactly as in a006:167-187; unlike the
reference they are re-derived from the input on every call, also in training mode (the
reference freezes them at the first training forward -- SURVEY appendix D.2)."""
from typing import Optional

from torch import nn, Tensor

from a010_StateRecorder import StateRecorder
from swinfuse import ops


class MyPadding(nn.Module):
    def __init__(self, belongs_to_encoder: bool, window_size: tuple, use_dual_path: bool,
                 feature_shape_recorder: StateRecorder, padding_size_recorder: StateRecorder):
        super().__init__()
        self.belongs_to_encoder, self.window_size, self.use_dual_path = belongs_to_encoder, window_size, use_dual_path
        self.feature_shape_hw: tuple = tuple()
        self.padding_size: tuple = tuple()
        self.feature_shape_recorder, self.padding_size_recorder = feature_shape_recorder, padding_size_recorder

    @staticmethod
    def calculate_padding_size(current_length, window_size):
        return (window_size - current_length % window_size) % window_size

    def get_feature_shape_hw(self):
        return self.feature_shape_hw

    def set_feature_shape_hw(self, new_feature_shape_hw):
        self.feature_shape_hw = new_feature_shape_hw

    def get_padding_size(self):
        return self.padding_size

    def set_padding_size(self, new_padding_size):
        self.padding_size = new_padding_size

    def do_padding_for_one_tensor(self, tensor: Tensor) -> Tensor:
        return ops.pad_reflect(tensor, *self.padding_size)

    def undo_padding_for_one_tensor(self, tensor: Tensor, add: Optional[Tensor] = None) -> Tensor:
        return ops.crop(tensor, *self.padding_size, add=add)

    def do_padding(self, x: Tensor, y: Optional[Tensor]):
        if y is None:
            return self.do_padding_for_one_tensor(x), None
        return ops.dual_path(lambda: self.do_padding_for_one_tensor(x), lambda: self.do_padding_for_one_tensor(y))

    def undo_padding(self, x, y, add_x=None, add_y=None):
        if y is None:
            return self.undo_padding_for_one_tensor(x, add_x), None
        return ops.dual_path(lambda: self.undo_padding_for_one_tensor(x, add_x), lambda: self.undo_padding_for_one_tensor(y, add_y))

    def forward(self, x, y, skip=None):
        """``skip`` (decoder only, optional (skip_x, skip_y)): U-Net skip tensors added right after
        the crop in the same kernel (a013:222-225)."""
        if self.belongs_to_encoder:
            h, w = x.shape[-2:]
            self.feature_shape_hw = (h, w)
            self.padding_size = (self.calculate_padding_size(h, self.window_size[0]),
                                 self.calculate_padding_size(w, self.window_size[1]))
            self.feature_shape_recorder.record(self.feature_shape_hw)
            self.padding_size_recorder.record(self.padding_size)
            x, y = self.do_padding(x, y if self.use_dual_path else None)
        else:
            self.feature_shape_hw = self.feature_shape_recorder.read()
            self.padding_size = self.padding_size_recorder.read()
            sx, sy = skip if skip is not None else (None, None)
            x, y = self.undo_padding(x, y if self.use_dual_path else None, sx, sy)
        return (x, y) if self.use_dual_path else x

    def forward_(self, x, y):
        return self(x, y)
