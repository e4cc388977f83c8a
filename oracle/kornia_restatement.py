"""Restatement of the two kornia pieces a008_loss.py uses  --  TEST INFRASTRUCTURE.

PARITY UNPINNED: ``kornia`` is a third-party dependency of the reference (imported at
a008:3-4, version not pinned by the reference -- it ships no requirements file) and is not
installed in this image, nor vendored under /root/reference.  The two classes below restate
the published algorithms of ``kornia.losses.MS_SSIMLoss`` (Zhao et al. "Loss functions for
image restoration", MS-SSIM + L1 mix) and ``kornia.filters.Sobel`` from memory of the
upstream source; they could not be checked against kornia itself.  Call sites in the
reference: a008:24 (``MS_SSIMLoss()`` default args), a008:37 (``Sobel()`` default args),
a008:109-110, a008:187-192.

This is NOT kornia and must not be labelled as such in any report.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class MS_SSIMLoss(nn.Module):
    def __init__(self, sigmas=(0.5, 1.0, 2.0, 4.0, 8.0), data_range=1.0, K=(0.01, 0.03), alpha=0.025,
                 compensation=200.0, reduction="mean"):
        super().__init__()
        self.DR = data_range
        self.C1 = (K[0] * data_range) ** 2
        self.C2 = (K[1] * data_range) ** 2
        self.pad = int(2 * sigmas[-1])
        self.alpha = alpha
        self.compensation = compensation
        self.reduction = reduction
        size = int(4 * sigmas[-1] + 1)
        g = torch.zeros(3 * len(sigmas), 1, size, size)
        for i, s in enumerate(sigmas):
            g2 = self._gauss_2d(size, s)
            g[3 * i + 0, 0], g[3 * i + 1, 0], g[3 * i + 2, 0] = g2, g2, g2
        self.register_buffer("_g_masks", g)

    @staticmethod
    def _gauss_1d(size: int, sigma: float) -> torch.Tensor:
        coords = torch.arange(size, dtype=torch.float) - size // 2
        g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
        return g / g.sum()

    def _gauss_2d(self, size: int, sigma: float) -> torch.Tensor:
        v = self._gauss_1d(size, sigma)
        return torch.outer(v, v)

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        ch = img1.shape[-3]
        g = self._g_masks
        mux = F.conv2d(img1, g, groups=ch, padding=self.pad)
        muy = F.conv2d(img2, g, groups=ch, padding=self.pad)
        mux2, muy2, muxy = mux * mux, muy * muy, mux * muy
        sigmax2 = F.conv2d(img1 * img1, g, groups=ch, padding=self.pad) - mux2
        sigmay2 = F.conv2d(img2 * img2, g, groups=ch, padding=self.pad) - muy2
        sigmaxy = F.conv2d(img1 * img2, g, groups=ch, padding=self.pad) - muxy
        lc = (2 * muxy + self.C1) / (mux2 + muy2 + self.C1)
        cs = (2 * sigmaxy + self.C2) / (sigmax2 + sigmay2 + self.C2)
        lm = lc[:, -1] * lc[:, -2] * lc[:, -3]
        pics = cs.prod(dim=1)
        loss_ms_ssim = 1 - lm * pics
        loss_l1 = F.l1_loss(img1, img2, reduction="none")
        gaussian_l1 = F.conv2d(loss_l1, g[-ch:], groups=ch, padding=self.pad).mean(1)
        loss = self.compensation * (self.alpha * loss_ms_ssim + (1 - self.alpha) * gaussian_l1 / self.DR)
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        return loss


class Sobel(nn.Module):
    def __init__(self, normalized: bool = True, eps: float = 1e-6):
        super().__init__()
        self.normalized = normalized
        self.eps = eps
        kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])
        k = torch.stack([kx, kx.t()])
        if normalized:
            k = k / k.abs().sum(dim=(-2, -1), keepdim=True)
        self.register_buffer("_k", k[:, None])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, c, h, w = x.shape
        xp = F.pad(x.reshape(b * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
        gxy = F.conv2d(xp, self._k.to(x.dtype))
        gx, gy = gxy[:, 0], gxy[:, 1]
        return torch.sqrt(gx * gx + gy * gy + self.eps).reshape(b, c, h, w)


def ssim_loss(*a, **k):  # a008:3 imports it; disabled by default (A000:34)
    raise NotImplementedError("ssim_loss is not restated (CHOOSE_MS_SSIM=True in A000_CONFIG.py:34)")


class PSNRLoss(nn.Module):  # a008:3; disabled by default (A000:39)
    def __init__(self, max_val: float = 1.0):
        super().__init__()
        self.max_val = max_val

    def forward(self, a, b):
        mse = F.mse_loss(a, b)
        return -10.0 * torch.log10(self.max_val ** 2 / mse)


class Canny(nn.Module):  # a008:4; disabled by default (A000:37)
    def forward(self, x):
        raise NotImplementedError("Canny is not restated (CHOOSE_CANNY_ELSE_SOBEL=False in A000_CONFIG.py:37)")


def total_loss(fusion: torch.Tensor, ir: torch.Tensor, vis: torch.Tensor, ms: MS_SSIMLoss, sobel: Sobel):
    """MyLoss.calcu_total_loss (a008:226-282) with A000_CONFIG.py:34-52 constants."""
    ssim = (0.2 * ms(fusion, ir) + 0.8 * ms(fusion, vis)) * 0.305
    tex = (sobel(fusion) - torch.max(sobel(ir), sobel(vis))).abs().mean() * 250
    inten = (fusion - torch.max(ir, vis)).abs().sum() / fusion.numel() * 45
    return ssim / 3 + tex / 3 + inten / 3
