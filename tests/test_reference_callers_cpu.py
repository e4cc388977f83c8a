"""The reference's own callers against the drop-in modules (VERDICT r1 next #4d) -- CPU only.

Runs only where /root/reference exists (the authoring container; the GPU box has no reference tree):
  * the REAL a016_train.py / a017_test.py import unchanged with dropin/ ahead on sys.path and build MyModel from the
    REAL A000_CONFIG exactly as a016:26-40 / a017:19-36 do (33,145,973 parameters, the 3,139-key state dict);
  * the a008 glue (weights, scales, ratios, the 5-digit bookkeeping: a008:226-311) restated in
    oracle/kornia_restatement.total_loss equals the REAL a008_loss.MyLoss run over the kornia stub.  This pins the
    glue; the kornia arithmetic inside the stub stays unpinned (kornia is not installed -- DESIGN.md section 4).
"""
import importlib
import os
import sys
import types

import pytest
import torch

from oracle import kornia_restatement as kr
from oracle.ref_import import REFERENCE_ROOT, ReferenceModules, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the authoring container")

_CALLERS = ("A000_CONFIG", "a015_dataset", "a016_train", "a017_test")


def _inert_shims():
    """colorama / matplotlib / mpl_toolkits are not installed here: inert stand-ins so that the
    callers import; none of them is on the hot path."""
    shims = {}
    if "colorama" not in sys.modules:
        colorama = types.ModuleType("colorama")
        colorama.init = lambda *a, **k: None
        colorama.Fore = types.SimpleNamespace(CYAN="", GREEN="", YELLOW="", RED="")
        shims["colorama"] = colorama
    if "matplotlib" not in sys.modules:
        mpl, pyplot, axes = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot"), types.ModuleType("matplotlib.axes")
        axes.Axes = object
        mpl.pyplot, mpl.axes = pyplot, axes
        shims.update({"matplotlib": mpl, "matplotlib.pyplot": pyplot, "matplotlib.axes": axes})
    if "mpl_toolkits" not in sys.modules:
        tk, grid = types.ModuleType("mpl_toolkits"), types.ModuleType("mpl_toolkits.axes_grid1")
        grid.ImageGrid = object
        tk.axes_grid1 = grid
        shims.update({"mpl_toolkits": tk, "mpl_toolkits.axes_grid1": grid})
    return shims


def test_real_callers_import_and_build_the_dropin_model():
    import swinfuse
    from oracle.ref_import import _install_kornia_stub
    _install_kornia_stub()          # a016 does `from a008_loss import MyLoss`: the drop-in a008 needs no kornia, a015 neither
    shims = _inert_shims()
    sys.modules.update(shims)
    saved = {m: sys.modules.pop(m) for m in _CALLERS if m in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    dropin_dir = swinfuse.install_dropin()   # ahead of the reference tree
    try:
        a016 = importlib.import_module("a016_train")
        a017 = importlib.import_module("a017_test")
        assert os.path.dirname(os.path.abspath(a016.__file__)) == REFERENCE_ROOT
        assert os.path.dirname(os.path.abspath(a017.__file__)) == REFERENCE_ROOT
        # the model class the callers bound is the drop-in one, the configuration is the reference's own
        assert os.path.dirname(os.path.abspath(sys.modules[a016.MyModel.__module__].__file__)) == dropin_dir
        assert a017.MyModel is a016.MyModel
        CFG = a016.CFG
        assert os.path.dirname(os.path.abspath(CFG.__file__)) == REFERENCE_ROOT
        model = a016.MyModel(
            window_size=CFG.WINDOW_SIZE, merging_size=CFG.MERGING_SIZE, in_dims_list=CFG.IN_DIMS_LIST,
            out_dims_list=CFG.OUT_DIMS_LIST, att_num_heads=CFG.ATT_NUM_HEADS,
            att_dims_per_head_ratio=CFG.ATT_DIMS_PER_HEAD_RATIO, attention_drop_ratio=CFG.ATTENTION_DROP_RATIO,
            linear_after_att_drop_ratio=CFG.LINEAR_AFTER_ATT_DROP_RATIO, mlp_hidden_dims_ratio=CFG.MLP_HIDDEN_DIMS_RATIO,
            mlp_activation_func=CFG.MLP_ACTIVATION_FUNC, mlp_drop_ratio=CFG.MLP_DROP_RATIO,
            final_layer_att_dims_per_head_ratio=CFG.FINAL_LAYER_ATT_DIMS_PER_HEAD_RATIO,
            final_conv_layer_kernel_size=CFG.FINAL_CONV_LAYER_KERNEL_SIZE,
            final_layer_mlp_hidden_dims_ratio=CFG.FINAL_LAYER_MLP_HIDDEN_DIMS_RATIO)
        model.apply(a016.init_params)                      # a016:42: re-initialises by isinstance(nn.Linear / nn.Conv2d)
        assert sum(p.numel() for p in model.parameters()) == 33_145_973
        assert len(model.state_dict()) == 3139
        # the optimizer / scheduler a016:67-72 builds accept the drop-in model's parameters
        opt = torch.optim.Adam(model.parameters(), lr=CFG.LR)
        torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer=opt, T_0=CFG.SCHEDULER_T0, eta_min=CFG.MINIMUM_LR)
        # and the model refuses CPU tensors instead of computing them some other way
        with pytest.raises(swinfuse.SwinFuseError):
            model(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for m in _CALLERS:
            sys.modules.pop(m, None)
        sys.modules.update(saved)
        for m in shims:
            sys.modules.pop(m, None)


@pytest.mark.parametrize("shape", [(2, 1, 40, 56), (1, 1, 64, 64)])
def test_a008_glue_restatement_equals_the_real_myloss_over_the_stub(shape):
    g = torch.Generator().manual_seed(11)
    fusion = torch.rand(shape, generator=g).requires_grad_(True)
    ir, vis = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    with ReferenceModules() as ref:
        real = ref.a008_loss.MyLoss()                    # the reference's class; kornia.* resolves to the stub
        total, detail = real.calcu_total_loss(fusion_images=fusion, ir_images=ir, vis_images=vis)
        (g_real,) = torch.autograd.grad(total, fusion)
        means = real.calcu_history_mean_and_clear_and_save_to_mean_recorder()
    f2 = fusion.detach().clone().requires_grad_(True)
    mine = kr.total_loss(f2, ir, vis, kr.MS_SSIMLoss(), kr.Sobel())
    (g_mine,) = torch.autograd.grad(mine, f2)
    assert float(mine.detach()) == pytest.approx(float(total.detach()), rel=1e-6)
    assert torch.allclose(g_mine, g_real, rtol=1e-5, atol=1e-9)
    assert detail["total_loss"] == round(float(total.detach()), 5) and detail["psnr_loss"] == 0
    assert means["total_loss_mean"] == pytest.approx(detail["total_loss"], abs=1e-5)
