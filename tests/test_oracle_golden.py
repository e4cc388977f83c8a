"""Pin the oracle: the restatement must reproduce the fixtures generated from the real
reference (oracle/make_golden.py).  CPU only."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle.make_golden import small_cfg


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_state_dict_contract(golden_dir):
    with gzip.open(os.path.join(golden_dir, "state_dict_keys.json.gz"), "rt") as f:
        gold = json.load(f)
    spec = fo.state_dict_spec()
    assert len(spec) == 3139 == len(gold["rows"])
    assert [[k, list(s), c] for k, s, c in spec] == gold["rows"]
    assert len({c for _, _, c in spec}) == 1459
    assert gold["n_params"] == 33145973


def test_index_maps_match_reference_helpers(golden_dir):
    g = _load(golden_dir, "blocks_patch_pad.npz")
    np.testing.assert_array_equal(fo.relative_position_index((7, 7)), g["idx/relative_position_indices"])
    for h, w in [(14, 21), (7, 7), (35, 28)]:
        np.testing.assert_array_equal(fo.shift_mask(h, w, (7, 7)), g[f"idx/mask_{h}x{w}"])
        for shifted, key in ((False, "partition"), (True, "partition_shifted")):
            src = fo.window_token_source_index(h, w, (7, 7), shifted)
            np.testing.assert_array_equal(src[..., 0] * w + src[..., 1], g[f"idx/{key}_{h}x{w}"])


def test_prototype_relative_position_formula():
    # a001_prototype_unit_test/a002_relative_position_encoding.py:23-24:
    # (col//7 - row//7 + 6, col%7 - row%7 + 6) with row = query token, col = key token
    idx = fo.relative_position_index((7, 7))
    for row in (0, 5, 13, 48):
        for col in (0, 7, 20, 48):
            assert idx[0, row, col] == col // 7 - row // 7 + 6
            assert idx[1, row, col] == col % 7 - row % 7 + 6


def test_window_attention_cases(golden_dir):
    g = _load(golden_dir, "window_attention.npz")
    cases = json.loads(bytes(g["cases"]).decode())
    assert len(cases) == 12
    for c in cases:
        t = c["tag"]
        p = {k[len(t) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(t + "/p/")}
        q, kv = torch.from_numpy(g[t + "/q"]), torch.from_numpy(g[t + "/kv"])
        out = fo.window_attention(q, kv, p, "", c["nh"], c["d"], (7, 7), c["shifted"])
        np.testing.assert_array_equal(out.numpy(), g[t + "/out"])


def test_block_pair_patch_and_pad(golden_dir):
    g = _load(golden_dir, "blocks_patch_pad.npz")
    T = lambda k: torch.from_numpy(g[k])
    p = {k[len("pair/p/"):]: T(k) for k in g.files if k.startswith("pair/p/")}
    ox, oy = fo.self_and_cross_block_pair(T("pair/x"), T("pair/y"), p, "", 4, 4, (7, 7))
    np.testing.assert_array_equal(ox.numpy(), g["pair/ox"])
    np.testing.assert_array_equal(oy.numpy(), g["pair/oy"])
    pe = {k[len("enc/p/"):]: T(k) for k in g.files if k.startswith("enc/p/")}
    pd = {k[len("dec/p/"):]: T(k) for k in g.files if k.startswith("dec/p/")}
    ex = fo.patch_layer(T("enc/x"), pe, "", "x", True, (2, 2))
    ey = fo.patch_layer(T("enc/y"), pe, "", "y", True, (2, 2))
    np.testing.assert_array_equal(ex.numpy(), g["enc/ox"])
    np.testing.assert_array_equal(ey.numpy(), g["enc/oy"])
    np.testing.assert_array_equal(fo.patch_layer(ex, pd, "", "x", False, (2, 2)).numpy(), g["dec/ox"])
    np.testing.assert_array_equal(fo.patch_layer(ey, pd, "", "y", False, (2, 2)).numpy(), g["dec/oy"])
    px, pad = fo.pad_reflect(T("pad/x"), (7, 7))
    assert pad == (5, 3)
    np.testing.assert_array_equal(px.numpy(), g["pad/px"])
    np.testing.assert_array_equal(fo.crop(px, pad).numpy(), g["pad/cx"])
    # the explicit index map is the same function
    idx = fo.reflect_pad_index(9, 11, 5, 3)
    x = g["pad/x"]
    np.testing.assert_array_equal(x[:, :, idx[..., 0], idx[..., 1]], g["pad/px"])


@pytest.mark.parametrize("tag,shape", [("b2_64", (2, 64, 64)), ("65x97", (1, 65, 97)), ("256", (1, 256, 256))])
def test_default_model_forward(golden_dir, tag, shape):
    g = _load(golden_dir, "model_default.npz")
    sd = fo.synth_state_dict()
    ir, vis = fo.synth_inputs(*shape)
    with torch.no_grad():
        out = fo.model_forward(sd, ir, vis)
    # same ops in the same order as the reference -> bit identical on the same machine;
    # allow fp32 round-off across CPUs / BLAS thread counts
    np.testing.assert_allclose(out.numpy(), g["out_" + tag], rtol=0, atol=2e-5)


def test_small_model_eval_and_train_forward(golden_dir):
    g = _load(golden_dir, "model_small.npz")
    cfg = small_cfg()
    sd = fo.synth_state_dict(cfg, seed=3)
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    with torch.no_grad():
        np.testing.assert_allclose(fo.model_forward(sd, ir, vis, cfg).numpy(), g["out_eval"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(fo.model_forward(sd, ir, vis, cfg, training=True).numpy(), g["out_train"],
                                   rtol=0, atol=2e-5)


def test_small_model_gradients_via_autograd(golden_dir):
    """Backward oracle = autograd over the restated forward; pinned against the reference's
    own autograd gradients."""
    g = _load(golden_dir, "model_small.npz")
    cfg = small_cfg()
    sd = fo.synth_state_dict(cfg, seed=3)
    canon = {}
    for k, _, c in fo.state_dict_spec(cfg):
        canon[k] = c
    leaves = {}
    p = {}
    for k, v in sd.items():
        c = canon[k]
        if c not in leaves:
            leaves[c] = v.clone().requires_grad_(v.is_floating_point())
        p[k] = leaves[c]
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    out = fo.model_forward(p, ir, vis, cfg, training=True)
    (out * torch.from_numpy(g["grad_weight"])).sum().backward()
    n = 0
    for k in g.files:
        if not k.startswith("grad::"):
            continue
        name = k[len("grad::"):]
        got = leaves[canon[name]].grad
        ref = g[k]
        scale = max(1e-6, float(np.abs(ref).max()))
        assert float((got - torch.from_numpy(ref)).abs().max()) <= 2e-4 * scale + 1e-6, name
        n += 1
    assert n > 100
