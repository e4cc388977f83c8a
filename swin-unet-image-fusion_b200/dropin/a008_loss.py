"""Drop-in for the reference's a008_loss.py: ``MyLoss`` with the same attributes, ``calcu_total_loss`` signature and
bookkeeping (a008:13-62, 226-311), computed by libswinfuse's sf_fusion_loss kernels (value, the three logged terms
and the gradient w.r.t. the fused image in one C-ABI call).

The reference takes MS-SSIM+L1 and Sobel from kornia (a008:3-4, not vendored, not pinned); the kernels follow the
restatement in oracle/kornia_restatement.py.  Only the repo-default loss configuration is built (MS-SSIM, Sobel,
no PSNR term: A000_CONFIG.py:34-52); any other switch raises instead of falling back.  No kornia import is needed.
"""
import numpy as np
import torch
from torch import Tensor, nn

from a010_StateRecorder import StateRecorder

try:  # the caller's configuration module, exactly as the reference reads it (a008:9)
    import A000_CONFIG as MyConfig
except ImportError:  # stand-alone use: the repo defaults, A000_CONFIG.py:34-52
    class MyConfig:
        CHOOSE_MS_SSIM = True
        FUS_IR_SSIM_WEIGHT = 0.2
        CHOOSE_CANNY_ELSE_SOBEL = False
        USE_PSNR = False
        FUS_IR_PSNR_WEIGHT = 0.4
        SSIM_SCALE, TEXTURE_SCALE, INTENSITY_SCALE, PSNR_SCALE = 0.305, 250, 45, 0
        SSIM_LOSS_RATIO = TEXTURE_LOSS_RATIO = INTENSITY_LOSS_RATIO = 1 / 3
        PSNR_LOSS_RATIO = 0


class MyLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.use_multi_scale_ssim = MyConfig.CHOOSE_MS_SSIM
        self.choose_canny = MyConfig.CHOOSE_CANNY_ELSE_SOBEL
        self.use_psnr = MyConfig.USE_PSNR
        if not self.use_multi_scale_ssim or self.choose_canny or self.use_psnr:
            raise NotImplementedError("libswinfuse builds the default loss only: CHOOSE_MS_SSIM=True, "
                                      "CHOOSE_CANNY_ELSE_SOBEL=False, USE_PSNR=False (A000_CONFIG.py:34-39)")
        self.max_val = 1.0
        self.fus_ir_ssim_weight = MyConfig.FUS_IR_SSIM_WEIGHT
        self.fus_vis_ssim_weight = 1 - self.fus_ir_ssim_weight
        self.ssim_scale = MyConfig.SSIM_SCALE
        self.texture_scale = MyConfig.TEXTURE_SCALE
        self.intensity_scale = MyConfig.INTENSITY_SCALE
        self.psnr_scale = MyConfig.PSNR_SCALE
        self.ssim_loss_ratio = MyConfig.SSIM_LOSS_RATIO
        self.texture_loss_ratio = MyConfig.TEXTURE_LOSS_RATIO
        self.intensity_loss_ratio = MyConfig.INTENSITY_LOSS_RATIO
        self.psnr_loss_ratio = MyConfig.PSNR_LOSS_RATIO
        self.loss_recorder_in_detail = StateRecorder()
        self.mean_loss_recorder = StateRecorder()
        from swinfuse.loss_ops import FusionLoss
        self._kernel = FusionLoss(fus_ir_ssim_weight=self.fus_ir_ssim_weight, ssim_scale=self.ssim_scale,
                                  texture_scale=self.texture_scale, intensity_scale=self.intensity_scale,
                                  ratios=(self.ssim_loss_ratio, self.texture_loss_ratio, self.intensity_loss_ratio))

    def calcu_total_loss(self, fusion_images: Tensor, ir_images: Tensor, vis_images: Tensor):
        """a008:226-282: returns (total loss tensor for backward(), dict of the scaled terms rounded to 5 digits)."""
        total_loss = self._kernel(fusion_images, ir_images, vis_images)
        total, ssim, texture, intensity = (round(v, 5) for v in self._kernel.last_terms.tolist())  # one device->host read
        loss_state_dict = {"ssim_loss": ssim, "texture_loss": texture, "intensity_loss": intensity, "psnr_loss": 0.0,
                           "total_loss": total}
        self.loss_recorder_in_detail.record(loss_state_dict)
        return total_loss, loss_state_dict

    def calcu_history_mean_and_clear_and_save_to_mean_recorder(self) -> dict:
        """a008:284-311"""
        values = [list(d.values()) for d in self.loss_recorder_in_detail.record_stack]
        means = [round(float(np.mean(col)), 5) for col in zip(*values)]
        self.loss_recorder_in_detail.delete_all()
        keys = ["ssim_loss_mean", "texture_loss_mean", "intensity_loss_mean", "psnr_loss_mean", "total_loss_mean"]
        means_dict = dict(zip(keys, means))
        self.mean_loss_recorder.record(means_dict)
        return means_dict
