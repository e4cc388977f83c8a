"""Gradient parity: the CUDA backward kernels against the reference's own
autograd gradients stored in the golden fixtures, and against autograd over the oracle."""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import fusion_oracle as fo
from oracle.make_golden import small_cfg
from tests.util import build_model, dropin, golden, rel_err, wa_cases

pytestmark = pytest.mark.gpu
TOL_GRAD = 2e-4   # fp32 backward, relative to the largest entry of each gradient tensor


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_window_attention_gradients_golden():
    dropin()
    from a001_WindowAttention import WindowAttention
    g, cases = wa_cases()
    for c in cases:
        t = c["tag"]
        wa = WindowAttention(in_out_dims=c["c"], num_heads=c["nh"], dims_per_head=c["d"], window_size=(7, 7),
                             use_cyclic_shift=c["shifted"], use_cross_attention=c["cross"], use_qkv_bias=True,
                             attention_drop_ratio=0.0, linear_after_att_drop_ratio=0.0)
        wa.load_state_dict({k[len(t) + 3:]: T(g[k]) for k in g.files if k.startswith(t + "/p/")})
        wa = wa.cuda()
        q = T(g[t + "/q"]).cuda().requires_grad_(True)
        kv = T(g[t + "/kv"]).cuda().requires_grad_(True) if c["cross"] else q
        out = wa(q, kv, kv)
        (out * T(g[t + "/gout"]).cuda()).sum().backward()
        assert rel_err(q.grad, T(g[t + "/gq"])) <= TOL_GRAD, (t, "gq", rel_err(q.grad, T(g[t + "/gq"])))
        if c["cross"]:
            assert rel_err(kv.grad, T(g[t + "/gkv"])) <= TOL_GRAD, (t, "gkv")
        # k_for_heads.bias has an analytically ZERO gradient (softmax is invariant to a per-row shift of
        # the scores); both sides hold round-off there, so errors are measured against the largest
        # parameter gradient of the module, with each tensor's own maximum when that is larger.
        gscale = max(float(np.abs(g[f"{t}/g/{n}"]).max()) for n, _ in wa.named_parameters())
        for n, prm in wa.named_parameters():
            ref = T(g[f"{t}/g/{n}"])
            assert prm.grad is not None, (t, n)
            err = float((prm.grad.cpu() - ref).abs().max()) / max(float(ref.abs().max()), 0.05 * gscale)
            assert err <= TOL_GRAD, (t, n, err)


def _oracle_grads(fn, tensors):
    leaves = [t.clone().requires_grad_(True) for t in tensors]
    out = fn(*leaves)
    return out, leaves


def test_mlp_and_prenorm_gradients_vs_oracle_autograd():
    sw = dropin()
    g = torch.Generator().manual_seed(5)
    c, hid = 16, 40
    x = torch.randn(2, c, 9, 7, generator=g)
    w1, b1 = torch.randn(hid, c, 1, 1, generator=g) * 0.3, 0.1 * torch.randn(hid, generator=g)
    w2, b2 = torch.randn(c, hid, 1, 1, generator=g) * 0.3, 0.1 * torch.randn(c, generator=g)
    lg, lb = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    go = torch.randn(2, c, 9, 7, generator=g)
    ref_in = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2, lg, lb)]
    rx, rw1, rb1, rw2, rb2, rlg, rlb = ref_in
    ref = rx + torch.nn.functional.conv2d(torch.nn.functional.elu(torch.nn.functional.conv2d(fo.layer_norm_c(rx, rlg, rlb), rw1, rb1)), rw2, rb2)
    (ref * go).sum().backward()
    cu = [t.clone().cuda().requires_grad_(True) for t in (x, w1, b1, w2, b2, lg, lb)]
    cx, cw1, cb1, cw2, cb2, clg, clb = cu
    out = sw.ops.mlp(cx, w1=cw1, b1=cb1, w2=cw2, b2=cb2, ln=(clg, clb), residual=cx, precision="fp32")
    (out * go.cuda()).sum().backward()
    for a, b, n in zip(cu, ref_in, "x w1 b1 w2 b2 ln_g ln_b".split()):
        assert rel_err(a.grad, b.grad) <= TOL_GRAD, (n, rel_err(a.grad, b.grad))


@pytest.mark.parametrize("encoder", [True, False])
def test_patch_layer_gradients_vs_oracle_autograd(encoder):
    sw = dropin()
    g = torch.Generator().manual_seed(8)
    cin, cout = (6, 16) if encoder else (16, 6)
    kin, kout = (cin * 4, cout) if encoder else (cin, cout * 4)
    x = torch.randn(2, cin, 10, 12, generator=g)
    p = {"mlp_layer_x.weight": torch.randn(kout, kin, 1, 1, generator=g) * 0.3, "mlp_layer_x.bias": 0.1 * torch.randn(kout, generator=g),
         "layer_norm_x.weight": 1 + 0.2 * torch.randn(kout, generator=g), "layer_norm_x.bias": 0.1 * torch.randn(kout, generator=g)}
    keys = list(p)
    ref_leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rx = x.clone().requires_grad_(True)
    ref = fo.patch_layer(rx, ref_leaves, "", "x", encoder, (2, 2))
    go = torch.randn(ref.shape, generator=g)
    (ref * go).sum().backward()
    cu = {k: v.clone().cuda().requires_grad_(True) for k, v in p.items()}
    cx = x.clone().cuda().requires_grad_(True)
    out = sw.ops.patch_layer(cx, w=cu[keys[0]], b=cu[keys[1]], ln_gamma=cu[keys[2]], ln_beta=cu[keys[3]], encoder=encoder,
                             merging_size=(2, 2), out_dims=cout, precision="fp32")
    assert rel_err(out, ref) <= 1e-4
    (out * go.cuda()).sum().backward()
    assert rel_err(cx.grad, rx.grad) <= TOL_GRAD
    for k in keys:
        assert rel_err(cu[k].grad, ref_leaves[k].grad) <= TOL_GRAD, (k, rel_err(cu[k].grad, ref_leaves[k].grad))


def test_pad_crop_gradients():
    sw = dropin()
    x = torch.randn(2, 4, 9, 11, requires_grad=True)
    ref = torch.nn.functional.pad(x, (0, 3, 0, 5), mode="reflect")
    go = torch.randn(ref.shape)
    (ref * go).sum().backward()
    cx = x.detach().clone().cuda().requires_grad_(True)
    out = sw.ops.pad_reflect(cx, 5, 3)
    (out * go.cuda()).sum().backward()
    assert torch.allclose(cx.grad.cpu(), x.grad, atol=1e-6)
    y = torch.randn(2, 4, 9, 11, requires_grad=True)
    s = torch.randn(2, 4, 7, 10, requires_grad=True)
    ref = y[:, :, :7, :10] + s
    go = torch.randn(ref.shape)
    (ref * go).sum().backward()
    cy, cs = y.detach().clone().cuda().requires_grad_(True), s.detach().clone().cuda().requires_grad_(True)
    out = sw.ops.crop(cy, 2, 1, add=cs)
    (out * go.cuda()).sum().backward()
    assert torch.equal(cy.grad.cpu(), y.grad) and torch.equal(cs.grad.cpu(), s.grad)


@pytest.mark.parametrize("training", [False, True])
def test_final_head_gradients_vs_oracle_autograd(training):
    sw = dropin()
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(2, 1, 13, 17, generator=g), torch.randn(2, 1, 13, 17, generator=g)
    p = {"final_layer.0.weight": torch.randn(2, 2, 3, 3, generator=g) * 0.4, "final_layer.0.bias": 0.1 * torch.randn(2, generator=g),
         "final_layer.1.weight": 1 + 0.2 * torch.randn(2, generator=g), "final_layer.1.bias": 0.1 * torch.randn(2, generator=g),
         "final_layer.1.running_mean": 0.1 * torch.randn(2, generator=g), "final_layer.1.running_var": 0.5 + torch.rand(2, generator=g),
         "final_layer.3.weight": torch.randn(1, 2, 3, 3, generator=g) * 0.4, "final_layer.3.bias": 0.1 * torch.randn(1, generator=g)}
    learn = [k for k in p if "running" not in k]
    ref_p = {k: (v.clone().requires_grad_(True) if k in learn else v.clone()) for k, v in p.items()}
    rx, ry = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    ref = fo.final_head(rx, ry, ref_p, training=training)
    go = torch.randn(ref.shape, generator=g)
    (ref * go).sum().backward()
    cu = {k: (v.clone().cuda().requires_grad_(True) if k in learn else v.clone().cuda()) for k, v in p.items()}
    cx, cy = x.clone().cuda().requires_grad_(True), y.clone().cuda().requires_grad_(True)
    out = sw.ops.final_head(cx, cy, w1=cu["final_layer.0.weight"], b1=cu["final_layer.0.bias"], bn_gamma=cu["final_layer.1.weight"],
                            bn_beta=cu["final_layer.1.bias"], running_mean=cu["final_layer.1.running_mean"],
                            running_var=cu["final_layer.1.running_var"], w2=cu["final_layer.3.weight"], b2=cu["final_layer.3.bias"],
                            training=training)
    assert rel_err(out, ref) <= 1e-4
    (out * go.cuda()).sum().backward()
    assert rel_err(cx.grad, rx.grad) <= TOL_GRAD and rel_err(cy.grad, ry.grad) <= TOL_GRAD
    gmax = max(float(ref_p[k].grad.abs().max()) for k in learn)
    for k in learn:
        if training and k == "final_layer.0.bias":
            # batch-stat BatchNorm removes the conv bias: its gradient is analytically zero
            assert float(cu[k].grad.abs().max()) <= 1e-4 * gmax
            continue
        assert rel_err(cu[k].grad, ref_p[k].grad) <= TOL_GRAD, (k, rel_err(cu[k].grad, ref_p[k].grad))


def test_small_model_training_step_gradients_match_reference_autograd():
    """Whole model, train mode (batch-stat BatchNorm), every one of the parameter gradients."""
    g = golden("model_small.npz")
    cfg = small_cfg()
    m = build_model(cfg, act=nn.ELU()).train()
    m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    out = m(ir.cuda(), vis.cuda())
    assert rel_err(out, T(g["out_train"])) <= 1e-4
    (out * T(g["grad_weight"]).cuda()).sum().backward()
    seen, n, worst = set(), 0, (0.0, "")
    gscale = float(np.median([np.abs(g[k]).max() for k in g.files if k.startswith("grad::")]))
    for name, prm in m.named_parameters():
        if id(prm) in seen:
            continue
        seen.add(id(prm))
        ref = T(g["grad::" + name])
        assert prm.grad is not None, name
        if name.endswith("k_for_heads.bias") or name == "final_layer.0.bias":
            # analytically zero gradients (softmax shift invariance; batch-stat BatchNorm removes the
            # conv bias): both sides hold round-off only
            assert float(prm.grad.abs().max()) <= 1e-3 * gscale, name
            n += 1
            continue
        e = float((prm.grad.cpu() - ref).abs().max()) / max(float(ref.abs().max()), 0.05 * gscale)
        worst = max(worst, (e, name))
        n += 1
    assert n > 100
    assert worst[0] <= 1e-3, worst


def test_bf16_mode_gradients_track_the_fp32_ones():
    """SF_PREC_BF16 operators run their forward AND backward GEMMs on tcgen05 with bf16 operands (fp32 accumulation; the
    attention-core adjoint on fp16 mma.sync): the gradients of a whole train-mode model step must stay close to the exact
    fp32 backward: 2e-2 in relative L2 over all parameters, 2e-1 of its own scale for any single tensor -- the same
    bounds tests/test_gpu_train.py holds this path to against the REFERENCE's autograd fixtures."""
    sw = dropin()
    g = golden("model_small.npz")
    cfg = small_cfg()
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    grads = {}
    for prec in ("fp32", "bf16"):
        sw.set_default_precision(prec)
        try:
            m = build_model(cfg, act=nn.ELU()).train()
            m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
            out = m(ir.cuda(), vis.cuda())
            (out * T(g["grad_weight"]).cuda()).sum().backward()
            seen, gr = set(), {}
            for name, prm in m.named_parameters():
                if id(prm) not in seen:
                    seen.add(id(prm))
                    gr[name] = prm.grad.detach().cpu().clone()
            grads[prec] = gr
        finally:
            sw.set_default_precision("fp32")
    gscale = float(np.median([float(v.abs().max()) for v in grads["fp32"].values()]))
    worst, num, den = (0.0, ""), 0.0, 0.0
    for name, ref in grads["fp32"].items():
        d = grads["bf16"][name] - ref
        e = float(d.abs().max()) / max(float(ref.abs().max()), 0.05 * gscale)
        if not (name.endswith("k_for_heads.bias") or name == "final_layer.0.bias"):   # analytically zero: round-off on both sides
            worst = max(worst, (e, name))
        num += float((d.double() ** 2).sum())
        den += float((ref.double() ** 2).sum())
    rel_l2 = (num / den) ** 0.5
    print("bf16-mode gradient deviation: rel L2 over all parameters", rel_l2, "worst tensor (max-norm)", worst)
    # the deviation is that of the bf16 forward activations (1e-2 on the fused image) seen through the backward pass;
    # sums with cancellation (the 13x13 bias tables) sit highest
    assert rel_l2 <= 2e-2, rel_l2
    assert worst[0] <= 2e-1, worst


@pytest.mark.parametrize("c,nh,d", [(24, 8, 3), (48, 8, 6), (96, 8, 12), (192, 8, 24), (384, 8, 48)])
@pytest.mark.parametrize("shifted", [False, True])
@pytest.mark.parametrize("cross", [False, True])
def test_tensor_core_attention_backward_tracks_the_fp32_kernel(c, nh, d, shifted, cross):
    """SF_PREC_BF16 window attention runs its attention-core adjoint on mma.sync (fp16 operands, per-window power-of-two
    normalisation of dO) and every GEMM of its backward (q / k / v recompute, data gradients, weight gradients) on tcgen05
    with bf16 operands and fp32 accumulation, so the gradients must agree with the exact fp32 backward to operand-rounding
    level (2e-2 of each tensor's scale: 8-bit mantissas, and the softmax amplifies the rounding of q and k), also for
    upstream gradients as small as a mean-reduced loss produces (1e-7)."""
    sw = dropin()
    from a001_WindowAttention import WindowAttention
    gen = torch.Generator().manual_seed(c + 2 * int(shifted) + int(cross))
    wa = WindowAttention(in_out_dims=c, num_heads=nh, dims_per_head=d, window_size=(7, 7), use_cyclic_shift=shifted,
                         use_cross_attention=cross, use_qkv_bias=True, attention_drop_ratio=0.0,
                         linear_after_att_drop_ratio=0.0)
    with torch.no_grad():
        for name, prm in wa.named_parameters():
            # projection weights at the kaiming scale a016 initialises them with (scores of order one), table and biases small
            std = 0.3 if "table" in name else (c ** -0.5 if prm.dim() > 1 else 0.1)
            prm.copy_(torch.randn(prm.shape, generator=gen) * std)
    wa = wa.cuda()
    q0 = torch.randn(2, c, 21, 28, generator=gen)
    kv0 = torch.randn(2, c, 21, 28, generator=gen)
    gout = torch.randn(2, c, 21, 28, generator=gen) * 1e-7
    res = {}
    for prec in ("fp32", "bf16"):
        sw.set_default_precision(prec)
        try:
            wa.zero_grad(set_to_none=True)
            q = q0.cuda().requires_grad_(True)
            kv = kv0.cuda().requires_grad_(True) if cross else q
            out = wa(q, kv, kv)
            (out * gout.cuda()).sum().backward()
            r = {"gq": q.grad.cpu().clone()}
            if cross:
                r["gkv"] = kv.grad.cpu().clone()
            for n, prm in wa.named_parameters():
                r[n] = prm.grad.cpu().clone()
            res[prec] = r
        finally:
            sw.set_default_precision("fp32")
    gscale = max(float(v.abs().max()) for n, v in res["fp32"].items() if n not in ("gq", "gkv"))
    for n, ref in res["fp32"].items():
        floor = 0.05 * gscale if n not in ("gq", "gkv") else 0.0
        e = float((res["bf16"][n] - ref).abs().max()) / max(float(ref.abs().max()), floor)
        assert e <= 2e-2, (n, e)


def test_direct_parameter_gradient_accumulation_equals_autograd_accumulation():
    """ops.set_direct_param_grads(True) (what swinfuse.train.DataParallelTrainer switches on): the kernels add parameter
    gradients straight into pre-existing .grad buffers instead of returning tensors for autograd to add.  Same numbers
    (up to the order of the fp32 atomics), and gradients accumulate over two backward passes like autograd's do."""
    sw = dropin()
    g = golden("model_small.npz")
    cfg = small_cfg()
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    res = {}
    for direct in (False, True):
        m = build_model(cfg, act=nn.ELU()).train()
        m.load_state_dict(fo.synth_state_dict(cfg, seed=3), strict=True)
        params = {n: p for n, p in m.named_parameters()}
        if direct:
            for p in params.values():
                p.grad = torch.zeros_like(p)
        sw.ops.set_direct_param_grads(direct)
        try:
            for _ in range(2):
                out = m(ir.cuda(), vis.cuda())
                (out * T(g["grad_weight"]).cuda()).sum().backward()
        finally:
            sw.ops.set_direct_param_grads(False)
        res[direct] = {n: p.grad.detach().cpu().clone() for n, p in params.items()}
    gscale = float(np.median([float(v.abs().max()) for v in res[False].values()]))
    for n, ref in res[False].items():
        if n.endswith("k_for_heads.bias") or n == "final_layer.0.bias":
            continue   # analytically zero gradients: round-off of the atomics' order on both sides
        e = float((res[True][n] - ref).abs().max()) / max(float(ref.abs().max()), 0.05 * gscale)
        assert e <= 1e-4, (n, e)


@pytest.mark.parametrize("b,h,w,c,hid", [(2, 9, 7, 24, 96), (1, 16, 16, 24, 4), (3, 23, 11, 48, 192), (2, 19, 13, 96, 384),
                                         (1, 17, 15, 192, 768), (2, 14, 14, 384, 1536), (4, 70, 70, 48, 192)])
def test_tcgen05_mlp_backward_vs_torch_autograd(b, h, w, c, hid):
    """SF_PREC_BF16 MLP backward: forward recompute and dX through k_tc_gemm2 (transposed weight images), dW / db through
    k_tc_wgrad (both operands MN-major bf16 tiles, TMEM accumulation over all token tiles) -- against fp32 torch autograd
    of  x + W2 ELU(W1 LN(x) + b1) + b2  (a003:21-50 + a004:29-38).  Shapes: ragged token counts (tail rows of the last
    tile), 1..12 m-blocks and 1..6 n-blocks of the weight-gradient GEMM, hidden < 16.  bf16 operands: 2e-2 of each
    tensor's scale."""
    sw = dropin()
    g = torch.Generator().manual_seed(1000 * c + hid + h)
    x = torch.randn(b, c, h, w, generator=g)
    w1, b1 = torch.randn(hid, c, 1, 1, generator=g) * c ** -0.5, 0.1 * torch.randn(hid, generator=g)
    w2, b2 = torch.randn(c, hid, 1, 1, generator=g) * hid ** -0.5, 0.1 * torch.randn(c, generator=g)
    lg, lb = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    gout = torch.randn(b, c, h, w, generator=g)
    names = ["x", "lg", "lb", "w1", "b1", "w2", "b2"]

    def leaves():
        return [t.clone().cuda().requires_grad_(True) for t in (x, lg, lb, w1, b1, w2, b2)]

    ref = leaves()
    rx, rlg, rlb, rw1, rb1, rw2, rb2 = ref
    n = torch.nn.functional.layer_norm(rx.permute(0, 2, 3, 1), (c,), rlg, rlb, 1e-5).permute(0, 3, 1, 2)
    hdn = torch.nn.functional.elu(torch.nn.functional.conv2d(n, rw1, rb1))
    out_ref = rx + torch.nn.functional.conv2d(hdn, rw2, rb2)
    (out_ref * gout.cuda()).sum().backward()

    got = leaves()
    gx, glg, glb, gw1, gb1, gw2, gb2 = got
    out = sw.ops.mlp(gx, w1=gw1, b1=gb1, w2=gw2, b2=gb2, ln=(glg, glb), residual=gx, precision="bf16")
    assert rel_err(out, out_ref) <= 2e-2
    (out * gout.cuda()).sum().backward()
    for name, a, r in zip(names, got, ref):
        err = float((a.grad - r.grad).abs().max() / r.grad.abs().max())
        assert err <= 2e-2, (name, err)
