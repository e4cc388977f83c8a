"""Parity of the CUDA path with the oracle and with the reference-generated golden fixtures
(fp32 mode: <= 1e-4 relative, BASELINE.json north_star).  Everything goes through the drop-in
modules -> ctypes -> C ABI."""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import fusion_oracle as fo
from oracle.make_golden import small_cfg
from tests.util import TOL_FP32, build_model, dropin, golden, rel_err, rel_l2, wa_cases

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_native_library_is_the_one_running():
    sw = dropin()
    sw.ops.reset_launch_count()
    x = sw.ops.as_fmap(torch.rand(1, 4, 5, 5, device="cuda"))   # NCHW -> NHWC: 1 kernel
    sw.ops.pad_reflect(x, 2, 2)                                 # 1 kernel
    assert sw.ops.launch_count() == 2
    with open("/proc/self/maps") as f:
        assert "libswinfuse.so" in f.read()


def test_window_attention_golden_cases():
    dropin()
    from a001_WindowAttention import WindowAttention
    g, cases = wa_cases()
    for c in cases:
        t = c["tag"]
        wa = WindowAttention(in_out_dims=c["c"], num_heads=c["nh"], dims_per_head=c["d"], window_size=(7, 7),
                             use_cyclic_shift=c["shifted"], use_cross_attention=c["cross"], use_qkv_bias=True,
                             attention_drop_ratio=0.0, linear_after_att_drop_ratio=0.0).eval()
        wa.load_state_dict({k[len(t) + 3:]: T(g[k]) for k in g.files if k.startswith(t + "/p/")})
        wa = wa.cuda()
        q = T(g[t + "/q"]).cuda()
        kv = T(g[t + "/kv"]).cuda() if c["cross"] else q
        with torch.no_grad():
            out = wa(q, kv, kv)
        assert out.shape == q.shape
        assert rel_err(out, T(g[t + "/out"])) <= TOL_FP32, t


def test_fused_prenorm_residual_matches_composition():
    """BasicBlock stage_1 path (LN inside the operator + residual) against the oracle."""
    sw = dropin()
    ops = sw.ops
    g = torch.Generator().manual_seed(3)
    b, c, h, w, nh, d = 2, 24, 14, 21, 8, 3
    x, y = torch.randn(b, c, h, w, generator=g), torch.randn(b, c, h, w, generator=g)
    p = {"q_for_heads.weight": torch.randn(nh * d, c, generator=g) * 0.3, "q_for_heads.bias": torch.randn(nh * d, generator=g) * 0.1,
         "k_for_heads.weight": torch.randn(nh * d, c, generator=g) * 0.3, "k_for_heads.bias": torch.randn(nh * d, generator=g) * 0.1,
         "v_for_heads.weight": torch.randn(nh * d, c, generator=g) * 0.3, "v_for_heads.bias": torch.randn(nh * d, generator=g) * 0.1,
         "linear_projection.weight": torch.randn(c, nh * d, generator=g) * 0.3, "linear_projection.bias": torch.randn(c, generator=g) * 0.1,
         "relative_position_bias_table": torch.randn(13, 13, generator=g)}
    gx, bx = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    gy, by = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    for shift in (False, True):
        ref = x + fo.window_attention(fo.layer_norm_c(x, gx, bx), fo.layer_norm_c(y, gy, by), p, "", nh, d, (7, 7), shift)
        cu = {k: v.cuda() for k, v in p.items()}
        got = ops.window_attention(
            x.cuda(), y.cuda(), wq=cu["q_for_heads.weight"], bq=cu["q_for_heads.bias"], wk=cu["k_for_heads.weight"],
            bk=cu["k_for_heads.bias"], wv=cu["v_for_heads.weight"], bv=cu["v_for_heads.bias"],
            wo=cu["linear_projection.weight"], bo=cu["linear_projection.bias"],
            bias_table=cu["relative_position_bias_table"], num_heads=nh, head_dim=d, window_size=(7, 7), shift=shift,
            ln_q=(gx.cuda(), bx.cuda()), ln_kv=(gy.cuda(), by.cuda()), residual=x.cuda(), precision="fp32")
        assert rel_err(got, ref) <= TOL_FP32


def test_block_pair_patch_layers_and_padding_golden():
    dropin()
    from a006_PaddingOperation import MyPadding
    from a010_StateRecorder import StateRecorder
    from a011_PatchOperation import PatchMergingAndLinearLayer
    from a012_SelfAndCrossBlockPair import SelfAndCrossBlockPair
    g = golden("blocks_patch_pad.npz")
    pair = SelfAndCrossBlockPair(in_out_dims=16, num_heads=4, dims_per_head=4, window_size=(7, 7), use_dual_path=True,
                                 use_qkv_bias=True, attention_drop_ratio=0.0, linear_after_att_drop_ratio=0.0,
                                 mlp_hidden_dims=40, mlp_activation_func=nn.ELU(), mlp_drop_ratio=0.0).eval()
    pair.load_state_dict({k[len("pair/p/"):]: T(g[k]) for k in g.files if k.startswith("pair/p/")})
    pair = pair.cuda()
    with torch.no_grad():
        ox, oy = pair(T(g["pair/x"]).cuda(), T(g["pair/y"]).cuda())
    assert rel_err(ox, T(g["pair/ox"])) <= TOL_FP32 and rel_err(oy, T(g["pair/oy"])) <= TOL_FP32

    enc = PatchMergingAndLinearLayer(True, True, 6, 16, StateRecorder(), (2, 2), nn.ELU()).eval()
    dec = PatchMergingAndLinearLayer(False, True, 16, 6, StateRecorder(), (2, 2), nn.ELU()).eval()
    enc.load_state_dict({k[len("enc/p/"):]: T(g[k]) for k in g.files if k.startswith("enc/p/")})
    dec.load_state_dict({k[len("dec/p/"):]: T(g[k]) for k in g.files if k.startswith("dec/p/")})
    enc, dec = enc.cuda(), dec.cuda()
    with torch.no_grad():
        ex, ey = enc(T(g["enc/x"]).cuda(), T(g["enc/y"]).cuda())
        dx, dy = dec(ex, ey)
    for got, key in ((ex, "enc/ox"), (ey, "enc/oy"), (dx, "dec/ox"), (dy, "dec/oy")):
        assert got.shape == T(g[key]).shape
        assert rel_err(got, T(g[key])) <= TOL_FP32, key

    r1, r2 = StateRecorder(), StateRecorder()
    pe = MyPadding(True, (7, 7), True, r1, r2)
    pd = MyPadding(False, (7, 7), True, r1, r2)
    px, py = pe(T(g["pad/x"]).cuda(), T(g["pad/y"]).cuda())
    assert torch.equal(px.cpu(), T(g["pad/px"])) and torch.equal(py.cpu(), T(g["pad/py"]))
    cx, cy = pd(px, py)
    assert torch.equal(cx.cpu(), T(g["pad/cx"])) and torch.equal(cy.cpu(), T(g["pad/cy"]))
    assert r1.record_stack == [] and r2.record_stack == []


@pytest.mark.parametrize("tag,shape", [("b2_64", (2, 64, 64)), ("65x97", (1, 65, 97)), ("256", (1, 256, 256))])
def test_default_model_against_reference_outputs(tag, shape):
    g = golden("model_default.npz")
    m = build_model().eval()
    m.load_state_dict(fo.synth_state_dict(), strict=True)
    ir, vis = fo.synth_inputs(*shape)
    with torch.no_grad():
        out = m(ir.cuda(), vis.cuda())
    ref = T(g["out_" + tag])
    assert out.shape == ref.shape and out.is_contiguous()
    assert rel_err(out, ref) <= TOL_FP32, (rel_err(out, ref), rel_l2(out, ref))
    assert rel_l2(out, ref) <= TOL_FP32


def test_small_model_eval_and_train_mode_forward():
    g = golden("model_small.npz")
    cfg = small_cfg()
    m = build_model(cfg, act=nn.ELU())
    m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    m.eval()
    with torch.no_grad():
        assert rel_err(m(ir.cuda(), vis.cuda()), T(g["out_eval"])) <= TOL_FP32
    m.train()
    with torch.no_grad():
        out = m(ir.cuda(), vis.cuda())
    assert rel_err(out, T(g["out_train"])) <= TOL_FP32
    bn = m.final_layer[1]
    assert rel_err(bn.running_mean, T(g["bn_running_mean"])) <= 1e-4
    assert rel_err(bn.running_var, T(g["bn_running_var"])) <= 1e-4
    assert int(bn.num_batches_tracked) == 8   # synth state starts at 7


def test_batch_64_config_is_batch_independent_and_matches_oracle():
    """BASELINE config 2 shape (B=64, 256x256): samples are independent, so sample i of the batch
    must equal a batch-1 run bit for bit, and two samples are checked against the CPU oracle."""
    m = build_model().eval()
    sd = fo.synth_state_dict()
    m.load_state_dict(sd, strict=True)
    ir, vis = fo.synth_inputs(64, 256, 256)
    with torch.no_grad():
        out = m(ir.cuda(), vis.cuda())
        assert out.shape == (64, 1, 256, 256) and bool(torch.isfinite(out).all())
        for i in (0, 63):
            one = m(ir[i:i + 1].cuda(), vis[i:i + 1].cuda())
            assert torch.equal(one[0], out[i])
            ref = fo.model_forward(sd, ir[i:i + 1], vis[i:i + 1])
            assert rel_err(out[i:i + 1], ref) <= TOL_FP32


def test_high_res_1024_padding_and_mask_paths():
    """BASELINE config 4: 1024x1024 exercises pad2 = 6 at stage 0, pad1 = 1 at stages 2-3 and
    5,476-window shift masks."""
    m = build_model().eval()
    sd = fo.synth_state_dict()
    m.load_state_dict(sd, strict=True)
    ir, vis = fo.synth_inputs(1, 1024, 1024, seed=9)
    with torch.no_grad():
        out = m(ir.cuda(), vis.cuda())
        ref = fo.model_forward(sd, ir, vis)
    assert rel_err(out, ref) <= TOL_FP32


def test_stage_taps_against_oracle():
    """Per-stage intermediate tensors (encoder stage outputs) against the oracle's taps."""
    m = build_model().eval()
    sd = fo.synth_state_dict()
    m.load_state_dict(sd, strict=True)
    ir, vis = fo.synth_inputs(1, 100, 76, seed=2)
    taps = {}
    with torch.no_grad():
        fo.model_forward(sd, ir, vis, taps=taps)
        x, y = ir.cuda(), vis.cuda()
        for i, stage in enumerate(m.encoder_list):
            for mod in stage:
                x, y = mod(x=x, y=y)
            assert rel_err(x, taps[f"enc{i}.x"]) <= TOL_FP32, i
            assert rel_err(y, taps[f"enc{i}.y"]) <= TOL_FP32, i
            # re-synchronise with the oracle so errors do not compound across stages
            x, y = taps[f"enc{i}.x"].cuda(), taps[f"enc{i}.y"].cuda()
    for mp in m.modules():  # drain the recorders the encoder pass filled
        pass
    m.feature_shape_recorder.delete_all()
    m.padding_size_recorder.delete_all()


def test_identical_modalities_raise_instead_of_exit():
    """a005:111-118: a cross-attention block fed x == y prints and exit()s in the reference; the
    drop-in raises instead."""
    dropin()
    from a005_BasicBlock import BasicBlock
    blk = BasicBlock(in_out_dims=8, num_heads=2, dims_per_head=4, window_size=(7, 7), use_cyclic_shift=False,
                     use_dual_path=True, use_cross_attr=True, use_qkv_bias=True, attention_drop_ratio=0.0,
                     linear_after_att_drop_ratio=0.0, mlp_hidden_dims=16, mlp_activation_func=nn.ELU(),
                     mlp_drop_ratio=0.0).cuda().eval()
    x = torch.rand(1, 8, 7, 7, device="cuda")
    with pytest.raises(ValueError):
        with torch.no_grad():
            blk(x, x.clone())
