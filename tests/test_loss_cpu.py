"""The restructured loss the CUDA kernels follow (oracle/loss_separable_torch.py: shared terms, de-duplicated
sigma channels, separable Gaussian windows) against the dense restatement of the a008 / kornia formulation
(oracle/kornia_restatement.py, parity unpinned: kornia itself is not installed).  Value and gradient."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))

from oracle import kornia_restatement as kr  # noqa: E402
from oracle.loss_separable_torch import FusionLoss  # noqa: E402


def test_fusion_loss_matches_dense_restatement_value_and_gradient():
    g = torch.Generator().manual_seed(7)
    f = torch.rand(2, 1, 48, 56, generator=g).requires_grad_(True)
    ir, vis = torch.rand(2, 1, 48, 56, generator=g), torch.rand(2, 1, 48, 56, generator=g)
    ref = kr.total_loss(f, ir, vis, kr.MS_SSIMLoss(), kr.Sobel())
    (gref,) = torch.autograd.grad(ref, f)
    got = FusionLoss()(f, ir, vis)
    (ggot,) = torch.autograd.grad(got, f)
    assert abs(float(got.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert float((ggot - gref).abs().max()) <= 1e-5 * float(gref.abs().max())
