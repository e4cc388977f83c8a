"""CPU-side checks of the boundary: the shared library loads and exports every symbol declared in
include/swinfuse.h, the ctypes mirror matches, the drop-in modules honour the reference's module /
state_dict contract, and nothing silently falls back to a CPU path."""
import gzip
import json
import os
import re

import pytest
import torch

from tests.util import GOLDEN, build_model, dropin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    sw = dropin()
    from swinfuse import _lib
    lib = sw.load()
    hdr = open(os.path.join(ROOT, "include", "swinfuse.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in swinfuse.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sf_abi_version() == 1


def test_struct_layouts_match_header_field_order():
    dropin()
    from swinfuse import _lib
    hdr = open(os.path.join(ROOT, "include", "swinfuse.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for cname, struct in (("sf_window_attn_params", _lib.WindowAttnParams), ("sf_mlp_params", _lib.MlpParams),
                          ("sf_patch_params", _lib.PatchParams), ("sf_head_params", _lib.HeadParams)):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", hdr).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(float|int|long long|size_t|void)\s*\*?", "", decl)
            names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
        mine = [n.rstrip("_") for n, _ in struct._fields_]
        assert names == mine, (cname, names, mine)


def test_state_dict_contract_of_dropin_model():
    m = build_model(device=None)
    with gzip.open(os.path.join(GOLDEN, "state_dict_keys.json.gz"), "rt") as f:
        gold = json.load(f)
    sd = m.state_dict()
    assert [[k, list(v.shape)] for k, v in sd.items()] == [[k, s] for k, s, _ in gold["rows"]]
    ptr = {}
    for k, v in sd.items():
        ptr.setdefault(v.data_ptr() if v.numel() else k, []).append(k)
    canon = {k: g[0] for g in ptr.values() for k in g}
    assert all(canon[k] == c for k, _, c in gold["rows"])
    assert sum(p.numel() for p in m.parameters()) == gold["n_params"]


def test_no_cpu_fallback():
    sw = dropin()
    m = build_model(device=None).eval()
    x = torch.rand(1, 1, 32, 32)
    with pytest.raises(sw.SwinFuseError, match="CUDA"):
        m(x, x + 0.1)


def test_unsupported_configurations_raise():
    sw = dropin()
    from torch import nn
    from a003_AutoPathMLP import AutoPathMLP
    with pytest.raises(sw.SwinFuseError):
        AutoPathMLP(8, 16, nn.ReLU(), True, 0.0)


def test_module_api_surface_matches_reference_names():
    dropin()
    import a001_WindowAttention, a002_AutoPathWinAtt, a006_PaddingOperation, a011_PatchOperation
    import a012_SelfAndCrossBlockPair, a013_ModelDefinition
    import inspect
    sig = inspect.signature(a001_WindowAttention.WindowAttention.__init__)
    assert list(sig.parameters)[1:] == ["in_out_dims", "num_heads", "dims_per_head", "window_size", "use_cyclic_shift",
                                        "use_cross_attention", "use_qkv_bias", "attention_drop_ratio",
                                        "linear_after_att_drop_ratio"]
    sig = inspect.signature(a013_ModelDefinition.MyModel.__init__)
    assert len(sig.parameters) - 1 == 14
    assert hasattr(a002_AutoPathWinAtt, "AutoPathWinAtt") and hasattr(a006_PaddingOperation, "MyPadding")
    assert hasattr(a011_PatchOperation, "PatchMergingAndLinearLayer")
    assert hasattr(a012_SelfAndCrossBlockPair, "SelfAndCrossBlockPair")
