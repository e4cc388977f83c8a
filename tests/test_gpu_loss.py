"""sf_fusion_loss on the GPU against the dense restatement of the a008 / kornia formulation
(oracle/kornia_restatement.py, run on the CPU; kornia itself is not installed: parity against kornia is unpinned).
Value, the three logged terms and the gradient w.r.t. the fused image; fp32 tolerance 2e-4 relative (the kernels
sum 33-tap windows in a different order than the dense 33x33 convolution and use 1/den instead of a division)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))

from oracle import kornia_restatement as kr  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 2e-4


def _oracle(x, ir, vis, clamp):
    x = x.clone().requires_grad_(True)
    f = torch.clamp(x, 0, 1) if clamp else x
    ms, sob = kr.MS_SSIMLoss(), kr.Sobel()
    total = kr.total_loss(f, ir, vis, ms, sob)
    (g,) = torch.autograd.grad(total, x)
    with torch.no_grad():
        ssim = (0.2 * ms(f, ir) + 0.8 * ms(f, vis)) * 0.305
        tex = (sob(f) - torch.max(sob(ir), sob(vis))).abs().mean() * 250
        inten = (f - torch.max(ir, vis)).abs().mean() * 45
    return total.detach(), torch.stack([total.detach(), ssim, tex, inten]), g


@pytest.mark.parametrize("shape,clamp", [((2, 1, 48, 56), False), ((1, 1, 37, 131), True), ((3, 1, 20, 20), True),
                                         ((2, 1, 256, 256), True), ((1, 1, 300, 129), False)])
def test_fusion_loss_value_terms_and_gradient_match_the_dense_restatement(shape, clamp):
    from swinfuse.loss_ops import FusionLoss
    g = torch.Generator().manual_seed(11)
    # the fused image leaves [0,1] on ~20% of the pixels so that the clamp (value and gradient mask) is exercised
    x = torch.rand(shape, generator=g) * 1.25 - 0.125 if clamp else torch.rand(shape, generator=g)
    ir, vis = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    ref_total, ref_terms, ref_g = _oracle(x, ir, vis, clamp)
    xd = x.cuda().requires_grad_(True)
    fn = FusionLoss(clamp01=clamp)
    total = fn(xd, ir.cuda(), vis.cuda())
    (got_g,) = torch.autograd.grad(total * 1.0, xd)
    terms = fn.last_terms.cpu()
    assert abs(float(total) - float(ref_total)) <= TOL * abs(float(ref_total))
    assert float((terms - ref_terms).abs().max()) <= TOL * float(ref_terms.abs().max())
    assert float((got_g.cpu() - ref_g).abs().max()) <= TOL * float(ref_g.abs().max())
    if clamp:
        outside = (x < 0) | (x > 1)
        assert outside.any() and float(got_g.cpu()[outside].abs().max()) == 0.0


def test_fusion_loss_upstream_gradient_scales_and_value_only_mode():
    from swinfuse.loss_ops import FusionLoss
    g = torch.Generator().manual_seed(5)
    x, ir, vis = (torch.rand(2, 1, 64, 64, generator=g).cuda() for _ in range(3))
    fn = FusionLoss()
    with torch.no_grad():
        v0 = fn(x, ir, vis)
    xg = x.clone().requires_grad_(True)
    v1 = fn(xg, ir, vis)
    (g1,) = torch.autograd.grad(v1, xg)
    xg2 = x.clone().requires_grad_(True)
    (g3,) = torch.autograd.grad(fn(xg2, ir, vis) * 3.0, xg2)
    assert float(v0) == float(v1)                      # deterministic two-stage sums
    assert torch.equal(g3, g1 * 3.0)


def test_fusion_loss_rejects_cpu_tensors():
    from swinfuse import SwinFuseError
    from swinfuse.loss_ops import FusionLoss
    x = torch.rand(1, 1, 32, 32)
    with pytest.raises(SwinFuseError):
        FusionLoss()(x, x, x)


def test_dropin_myloss_matches_the_restatement_and_keeps_the_a008_bookkeeping():
    """dropin/a008_loss.py: MyLoss.calcu_total_loss returns (tensor, dict rounded to 5 digits) like a008:226-282 and
    the history means of a008:284-311."""
    import swinfuse
    swinfuse.install_dropin()
    import importlib
    import a008_loss
    importlib.reload(a008_loss)
    g = torch.Generator().manual_seed(3)
    x, ir, vis = (torch.rand(2, 1, 40, 72, generator=g) for _ in range(3))
    ref_total, ref_terms, ref_g = _oracle(x, ir, vis, False)
    loss = a008_loss.MyLoss().cuda()
    xd = x.cuda().requires_grad_(True)
    total, d = loss.calcu_total_loss(xd, ir.cuda(), vis.cuda())
    total.backward()
    assert abs(float(total.detach()) - float(ref_total)) <= TOL * abs(float(ref_total))
    assert float((xd.grad.cpu() - ref_g).abs().max()) <= TOL * float(ref_g.abs().max())
    assert set(d) == {"ssim_loss", "texture_loss", "intensity_loss", "psnr_loss", "total_loss"}
    assert abs(d["ssim_loss"] - float(ref_terms[1])) <= 2e-4 * abs(float(ref_terms[1])) + 1e-5
    means = loss.calcu_history_mean_and_clear_and_save_to_mean_recorder()
    assert means["total_loss_mean"] == d["total_loss"] and not loss.loss_recorder_in_detail.record_stack


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_a016_training_loop_body_with_the_dropin_model_and_loss(precision):
    """The body of a016's training loop (a016:150-165) on the drop-in modules only: MyModel (kwargs of a016:26-40,
    init_params of a016:382-390) -> torch.clamp_ -> MyLoss.calcu_total_loss -> zero_grad / backward / torch.optim.Adam.step.
    The loss on a fixed batch must fall and the first step must agree between the fp32 and bf16 operator modes."""
    from torch import nn
    from oracle.make_golden import small_cfg
    from tests.util import build_model, dropin
    sw = dropin()
    import importlib
    import a008_loss
    importlib.reload(a008_loss)
    sw.set_default_precision(precision)
    try:
        torch.manual_seed(0)
        model = build_model(small_cfg(), act=nn.ELU()).train()

        def init_params(m):   # a016:382-390
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        model.apply(init_params)
        loss_fn = a008_loss.MyLoss().cuda()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        g = torch.Generator().manual_seed(9)
        ir, vis = torch.rand(2, 1, 64, 64, generator=g).cuda(), torch.rand(2, 1, 64, 64, generator=g).cuda()
        hist = []
        for _ in range(6):
            fusion = model(ir, vis)
            fusion = torch.clamp_(fusion, min=0, max=1)
            loss, detail = loss_fn.calcu_total_loss(fusion_images=fusion, ir_images=ir, vis_images=vis)
            opt.zero_grad()
            loss.backward()
            opt.step()
            hist.append(detail["total_loss"])
        assert all(np.isfinite(hist)), hist
        assert hist[-1] < hist[0], hist
        test_a016_training_loop_body_with_the_dropin_model_and_loss.first[precision] = hist[0]
        first = test_a016_training_loop_body_with_the_dropin_model_and_loss.first
        if len(first) == 2:
            assert abs(first["bf16"] - first["fp32"]) <= 2e-2 * abs(first["fp32"]), first
    finally:
        sw.set_default_precision("fp32")


test_a016_training_loop_body_with_the_dropin_model_and_loss.first = {}
