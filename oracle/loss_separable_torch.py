"""Fusion loss of a008_loss.py, separable / de-duplicated form -- TEST INFRASTRUCTURE.

Second restatement of the loss: the same restructuring the CUDA kernels use (csrc/loss_ops.cu), written with plain
torch ops so that the algebra (separable windows, shared mu_f / E[f^2], one copy per sigma) can be checked on the CPU
against the dense kornia-style formulation (oracle/kornia_restatement.py) by tests/test_loss_cpu.py.  Not used by the
product path.

    L = 1/3 * 0.305 * [0.2 MS(f, ir) + 0.8 MS(f, vis)]
      + 1/3 * 250   * mean |Sobel(f) - max(Sobel(ir), Sobel(vis))|
      + 1/3 * 45    * mean |f - max(ir, vis)|                       (A000_CONFIG.py:34-52, a008:226-282)

Status: the arithmetic of MS-SSIM+L1 and Sobel lives in the third-party ``kornia`` package, which the
reference does not pin and which is not installed here, so this is a from-memory restatement of the
published algorithms (PARITY UNPINNED).
The two MS-SSIM calls share mu_f and E[f^2], the 3x-duplicated sigma channels are de-duplicated and the
Gaussian windows are applied separably (mathematically identical to the kornia formulation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SIGMAS = (0.5, 1.0, 2.0, 4.0, 8.0)
K1, K2, ALPHA, COMPENSATION, DATA_RANGE = 0.01, 0.03, 0.025, 200.0, 1.0


class FusionLoss(torch.nn.Module):
    def __init__(self, fus_ir_ssim_weight=0.2, ssim_scale=0.305, texture_scale=250.0, intensity_scale=45.0,
                 ratios=(1 / 3, 1 / 3, 1 / 3)):
        super().__init__()
        size = int(4 * SIGMAS[-1] + 1)
        coords = torch.arange(size, dtype=torch.float32) - size // 2
        g1 = torch.stack([torch.exp(-(coords ** 2) / (2 * s ** 2)) for s in SIGMAS])
        g1 = g1 / g1.sum(dim=1, keepdim=True)
        # the 33x33 Gaussian windows are outer products g g^T: applied as a vertical then a horizontal 33-tap pass
        # (same zero padding, 16x fewer multiply-adds than the dense 2-D windows)
        self.register_buffer("g_v", g1[:, None, :, None].contiguous())               # (5,1,33,1)
        self.register_buffer("g_h", g1[:, None, None, :].contiguous())               # (5,1,1,33)
        kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
        self.register_buffer("sobel", torch.stack([kx, kx.t()])[:, None])             # (2,1,3,3)
        self.pad = size // 2
        self.w_ir, self.ssim_scale, self.texture_scale, self.intensity_scale, self.ratios = \
            fus_ir_ssim_weight, ssim_scale, texture_scale, intensity_scale, ratios
        self.c1, self.c2 = (K1 * DATA_RANGE) ** 2, (K2 * DATA_RANGE) ** 2

    def _blur(self, x):   # (B,1,H,W) -> (B,5,H,W), zero padding as kornia's MS_SSIMLoss
        return F.conv2d(F.conv2d(x, self.g_v, padding=(self.pad, 0)), self.g_h, padding=(0, self.pad), groups=len(SIGMAS))

    def _ms_ssim_l1(self, f, mu_f, e_ff, y):
        mu_y, e_yy, e_fy = self._blur(y), self._blur(y * y), self._blur(f * y)
        l = (2 * mu_f * mu_y + self.c1) / (mu_f * mu_f + mu_y * mu_y + self.c1)
        cs = (2 * (e_fy - mu_f * mu_y) + self.c2) / ((e_ff - mu_f * mu_f) + (e_yy - mu_y * mu_y) + self.c2)
        lm = l[:, -1] ** 3                      # the three duplicated sigma=8 channels
        # every sigma appears three times; an explicit product (prod()'s backward inspects the input for zeros on
        # the host, which breaks CUDA-graph capture of the training step)
        pics = (cs[:, 0] * cs[:, 1] * cs[:, 2] * cs[:, 3] * cs[:, 4]) ** 3
        d = (f - y).abs()
        l1 = F.conv2d(F.conv2d(d, self.g_v[-1:], padding=(self.pad, 0)), self.g_h[-1:], padding=(0, self.pad))[:, 0]
        return (COMPENSATION * (ALPHA * (1 - lm * pics) + (1 - ALPHA) * l1 / DATA_RANGE)).mean()

    def _sobel_mag(self, x):
        g = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), self.sobel)
        return torch.sqrt(g[:, 0:1] ** 2 + g[:, 1:2] ** 2 + 1e-6)

    def forward(self, fusion, ir, vis):
        mu_f, e_ff = self._blur(fusion), self._blur(fusion * fusion)
        ssim = (self.w_ir * self._ms_ssim_l1(fusion, mu_f, e_ff, ir)
                + (1 - self.w_ir) * self._ms_ssim_l1(fusion, mu_f, e_ff, vis)) * self.ssim_scale
        texture = (self._sobel_mag(fusion) - torch.max(self._sobel_mag(ir), self._sobel_mag(vis))).abs().mean() * self.texture_scale
        intensity = (fusion - torch.max(ir, vis)).abs().mean() * self.intensity_scale
        self.last_terms = (ssim.detach(), texture.detach(), intensity.detach())
        return ssim * self.ratios[0] + texture * self.ratios[1] + intensity * self.ratios[2]
