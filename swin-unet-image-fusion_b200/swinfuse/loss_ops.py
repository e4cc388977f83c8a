"""Fusion loss of a008_loss.py on the device: libswinfuse's sf_fusion_loss kernels.

    L = r_s * ssim_scale * [w_ir MS(f, ir) + (1 - w_ir) MS(f, vis)]
      + r_t * texture_scale * mean |Sobel(f) - max(Sobel(ir), Sobel(vis))|
      + r_i * intensity_scale * mean |f - max(ir, vis)|          (A000_CONFIG.py:34-52, a008:226-282)

One C-ABI call computes the value (and the three scaled terms a016 logs) together with d L / d fusion; autograd only
multiplies by the upstream scalar (sf_scale_by_scalar).  MS(.,.) and Sobel are kornia's MS_SSIMLoss() / Sobel() with
default arguments, restated (kornia is not vendored or pinned by the reference: parity against kornia itself is
unpinned, parity against oracle/kornia_restatement.py is tested on the GPU).  CUDA tensors only: no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import SwinFuseError, check


def _check_inputs(fusion, ir, vis):
    for name, t in (("fusion", fusion), ("ir", ir), ("vis", vis)):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise SwinFuseError(f"fusion loss: {name} must be a CUDA tensor (libswinfuse has no CPU path)")
        if t.dtype != torch.float32:
            raise SwinFuseError(f"fusion loss: {name} must be float32, got {t.dtype}")
    if fusion.dim() != 4 or fusion.shape[1] != 1 or ir.shape != fusion.shape or vis.shape != fusion.shape:
        raise SwinFuseError(f"fusion loss expects three (B,1,H,W) tensors, got {tuple(fusion.shape)}, "
                            f"{tuple(ir.shape)}, {tuple(vis.shape)}")


def _plane(t: torch.Tensor) -> torch.Tensor:
    # one channel: NCHW-contiguous and channels-last memory coincide, .contiguous() is then free
    return t.detach().contiguous()


class _FusionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fusion, ir, vis, cfg):
        _check_inputs(fusion, ir, vis)
        f, a, v = _plane(fusion), _plane(ir), _plane(vis)
        b, _, h, w = f.shape
        need_grad = fusion.requires_grad
        out = torch.empty(4, dtype=torch.float32, device=f.device)
        total = torch.empty((), dtype=torch.float32, device=f.device)
        grad = torch.empty_like(f) if need_grad else None
        p = _lib.FusionLossParams(f.data_ptr(), a.data_ptr(), v.data_ptr(), out.data_ptr(), total.data_ptr(),
                                  None if grad is None else grad.data_ptr(), b, h, w, int(cfg["clamp01"]),
                                  cfg["w_ir"], cfg["ssim_scale"], cfg["texture_scale"], cfg["intensity_scale"],
                                  cfg["r_ssim"], cfg["r_texture"], cfg["r_intensity"])
        lib = _lib.load()
        nbytes = lib.sf_fusion_loss_workspace_bytes(C.byref(p))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=f.device)
        check(lib.sf_fusion_loss(C.byref(p), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream),
              "sf_fusion_loss")
        ctx.grad = grad
        ctx.fusion_shape = fusion.shape
        ctx.mark_non_differentiable(out)
        return total, out

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        if ctx.grad is None:
            return None, None, None, None
        g = torch.empty_like(ctx.grad)
        gs = g_total.detach().to(torch.float32).contiguous()
        check(_lib.load().sf_scale_by_scalar(ctx.grad.data_ptr(), gs.data_ptr(), g.data_ptr(), g.numel(),
                                              torch.cuda.current_stream().cuda_stream), "sf_scale_by_scalar")
        return g.view(ctx.fusion_shape), None, None, None


class FusionLoss(torch.nn.Module):
    """forward(fusion, ir, vis) -> scalar loss (a008 MyLoss.calcu_total_loss without the bookkeeping).

    ``clamp01=True`` folds a016:153's ``torch.clamp(fusion, 0, 1)`` into the kernels (value and gradient mask).
    ``last_terms`` holds the device tensor (total, ssim, texture, intensity) of the last call, scaled as a008 logs them."""

    def __init__(self, fus_ir_ssim_weight=0.2, ssim_scale=0.305, texture_scale=250.0, intensity_scale=45.0,
                 ratios=(1 / 3, 1 / 3, 1 / 3), clamp01=False):
        super().__init__()
        self.cfg = dict(w_ir=float(fus_ir_ssim_weight), ssim_scale=float(ssim_scale), texture_scale=float(texture_scale),
                        intensity_scale=float(intensity_scale), r_ssim=float(ratios[0]), r_texture=float(ratios[1]),
                        r_intensity=float(ratios[2]), clamp01=bool(clamp01))
        self.last_terms = None

    def forward(self, fusion, ir, vis):
        total, terms = _FusionLossFn.apply(fusion, ir, vis, self.cfg)
        self.last_terms = terms
        return total
