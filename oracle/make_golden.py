"""Generate tests/golden/* by running the REAL reference  --  TEST INFRASTRUCTURE.

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden

Every fixture stores the reference's own outputs on deterministic inputs
(``fusion_oracle.synth_state_dict`` / ``synth_inputs``; module-level cases carry their
inputs and weights inline).  The GPU box has no /root/reference, so these files are what
pins both the oracle (tests/test_oracle_golden.py) and the CUDA path there.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np
import torch
from torch import nn

from oracle import fusion_oracle as fo
from oracle.ref_import import ReferenceModules, build_reference_model

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_CFG = dict(in_dims_list=[1, 8, 16], out_dims_list=[8, 16, 32], att_num_heads=4, att_dims_per_head_ratio=1 / 4)


def small_cfg() -> fo.FusionConfig:
    return fo.FusionConfig(**SMALL_CFG)


def _np(d):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def _rand_module_weights(mod: nn.Module, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        seen = set()
        for name, prm in mod.named_parameters():
            if id(prm) in seen:
                continue
            seen.add(id(prm))
            if "norm" in name:
                base = 1.0 if name.endswith("weight") else 0.0
                prm.copy_(base + 0.2 * torch.randn(prm.shape, generator=g))
            elif name.endswith("bias"):
                prm.copy_(0.1 * torch.randn(prm.shape, generator=g))
            elif name.endswith("relative_position_bias_table"):
                prm.copy_(torch.randn(prm.shape, generator=g))
            else:
                fan_in = int(np.prod(prm.shape[1:]))
                prm.copy_(torch.randn(prm.shape, generator=g) * (2.0 / fan_in) ** 0.5)


def gen_state_dict_keys(ref) -> None:
    m = build_reference_model(ref)
    sd = m.state_dict()
    ptr = {}
    for k, v in sd.items():
        ptr.setdefault(v.data_ptr() if v.numel() else k, []).append(k)
    canon = {k: g[0] for g in ptr.values() for k in g}
    rows = [[k, list(v.shape), canon[k]] for k, v in sd.items()]
    with gzip.open(os.path.join(GOLD, "state_dict_keys.json.gz"), "wt") as f:
        json.dump({"n_params": sum(p.numel() for p in m.parameters()), "rows": rows}, f)


def gen_model_outputs(ref) -> None:
    cfg = fo.FusionConfig()
    m = build_reference_model(ref, cfg).eval()
    m.load_state_dict(fo.synth_state_dict(cfg), strict=True)
    out = {}
    with torch.no_grad():
        for tag, (b, h, w) in {"256": (1, 256, 256), "65x97": (1, 65, 97), "b2_64": (2, 64, 64)}.items():
            ir, vis = fo.synth_inputs(b, h, w)
            out["out_" + tag] = m(ir, vis).numpy()
    np.savez_compressed(os.path.join(GOLD, "model_default.npz"), **out)

    # small config: eval + train-mode (batch-stat BatchNorm) forward, and gradients
    scfg = small_cfg()
    m = build_reference_model(ref, scfg)
    ssd = fo.synth_state_dict(scfg, seed=3)
    m.load_state_dict(ssd, strict=True)
    ir, vis = fo.synth_inputs(2, 37, 45, seed=5)
    out = {}
    m.eval()
    with torch.no_grad():
        out["out_eval"] = m(ir, vis).numpy()
    m.train()
    fused = m(ir, vis)
    out["out_train"] = fused.detach().numpy()
    gw = torch.rand(fused.shape, generator=torch.Generator().manual_seed(11))
    (fused * gw).sum().backward()
    out["grad_weight"] = gw.numpy()
    seen = set()
    for name, prm in m.named_parameters():  # named_parameters de-duplicates aliases
        if id(prm) in seen:
            continue
        seen.add(id(prm))
        out["grad::" + name] = prm.grad.numpy()
    out["bn_running_mean"] = m.final_layer[1].running_mean.numpy()
    out["bn_running_var"] = m.final_layer[1].running_var.numpy()
    np.savez_compressed(os.path.join(GOLD, "model_small.npz"), **out)


def gen_window_attention(ref) -> None:
    WA = ref.a001_WindowAttention.WindowAttention
    out = {}
    cases = []
    for ci, (c, nh, d, b, h, w) in enumerate([(16, 4, 4, 2, 14, 21), (24, 8, 3, 1, 21, 14), (12, 2, 8, 1, 7, 7)]):
        for shifted in (False, True):
            for cross in (False, True):
                tag = f"c{ci}_s{int(shifted)}_x{int(cross)}"
                wa = WA(in_out_dims=c, num_heads=nh, dims_per_head=d, window_size=(7, 7), use_cyclic_shift=shifted,
                        use_cross_attention=cross, use_qkv_bias=True, attention_drop_ratio=0.0,
                        linear_after_att_drop_ratio=0.0).eval()
                _rand_module_weights(wa, 100 + len(cases))
                g = torch.Generator().manual_seed(200 + len(cases))
                q = torch.randn(b, c, h, w, generator=g)
                kv = torch.randn(b, c, h, w, generator=g) if cross else q
                q.requires_grad_(True)
                if cross:
                    kv.requires_grad_(True)
                o = wa(q, kv, kv)
                go = torch.randn(o.shape, generator=g)
                (o * go).sum().backward()
                out[tag + "/q"], out[tag + "/kv"], out[tag + "/out"], out[tag + "/gout"] = \
                    q.detach().numpy(), kv.detach().numpy(), o.detach().numpy(), go.numpy()
                out[tag + "/gq"] = q.grad.numpy()
                if cross:
                    out[tag + "/gkv"] = kv.grad.numpy()
                for n, prm in wa.named_parameters():
                    out[f"{tag}/p/{n}"] = prm.detach().numpy()
                    out[f"{tag}/g/{n}"] = prm.grad.numpy()
                cases.append(dict(tag=tag, c=c, nh=nh, d=d, shifted=shifted, cross=cross))
    out["cases"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "window_attention.npz"), **out)


def gen_blocks_and_patch(ref) -> None:
    out = {}
    g = torch.Generator().manual_seed(7)
    # SelfAndCrossBlockPair (dual path): 4 BasicBlocks, a012:70-78
    pair = ref.a012_SelfAndCrossBlockPair.SelfAndCrossBlockPair(
        in_out_dims=16, num_heads=4, dims_per_head=4, window_size=(7, 7), use_dual_path=True, use_qkv_bias=True,
        attention_drop_ratio=0.0, linear_after_att_drop_ratio=0.0, mlp_hidden_dims=40,
        mlp_activation_func=nn.ELU(), mlp_drop_ratio=0.0).eval()
    _rand_module_weights(pair, 31)
    x, y = torch.randn(2, 16, 14, 14, generator=g), torch.randn(2, 16, 14, 14, generator=g)
    with torch.no_grad():
        ox, oy = pair(x, y)
    out.update({"pair/x": x, "pair/y": y, "pair/ox": ox, "pair/oy": oy})
    for k, v in pair.state_dict().items():
        out["pair/p/" + k] = v
    # PatchMergingAndLinearLayer encoder and decoder, a011:236-264
    Rec = ref.a010_StateRecorder.StateRecorder
    PM = ref.a011_PatchOperation.PatchMergingAndLinearLayer
    enc = PM(belongs_to_encoder=True, use_dual_path=True, in_dims=6, out_dims=16, patch_merging_size_recorder=Rec(),
             merging_or_unmerging_size=(2, 2), activation_func=nn.ELU()).eval()
    dec = PM(belongs_to_encoder=False, use_dual_path=True, in_dims=16, out_dims=6, patch_merging_size_recorder=Rec(),
             merging_or_unmerging_size=(2, 2), activation_func=nn.ELU()).eval()
    _rand_module_weights(enc, 41)
    _rand_module_weights(dec, 42)
    x, y = torch.randn(2, 6, 10, 12, generator=g), torch.randn(2, 6, 10, 12, generator=g)
    with torch.no_grad():
        ex, ey = enc(x, y)
        dx, dy = dec(ex, ey)
    out.update({"enc/x": x, "enc/y": y, "enc/ox": ex, "enc/oy": ey, "dec/ox": dx, "dec/oy": dy})
    for k, v in enc.state_dict().items():
        out["enc/p/" + k] = v
    for k, v in dec.state_dict().items():
        out["dec/p/" + k] = v
    # MyPadding, a006:167-187 (reflect pad bottom/right, LIFO crop)
    Pad = ref.a006_PaddingOperation.MyPadding
    r1, r2 = Rec(), Rec()
    pe = Pad(belongs_to_encoder=True, window_size=(7, 7), use_dual_path=True, feature_shape_recorder=r1,
             padding_size_recorder=r2).eval()
    pdm = Pad(belongs_to_encoder=False, window_size=(7, 7), use_dual_path=True, feature_shape_recorder=r1,
              padding_size_recorder=r2).eval()
    x, y = torch.randn(2, 3, 9, 11, generator=g), torch.randn(2, 3, 9, 11, generator=g)
    px, py = pe(x, y)
    cx, cy = pdm(px, py)
    out.update({"pad/x": x, "pad/y": y, "pad/px": px, "pad/py": py, "pad/cx": cx, "pad/cy": cy})
    # index-level known answers from the reference's own helpers
    WA = ref.a001_WindowAttention.WindowAttention
    wa = WA(8, 2, 4, (7, 7), True, False, True, 0.0, 0.0).eval()
    out["idx/relative_position_indices"] = wa.get_initial_relative_position_indices()
    for (h, w) in [(14, 21), (7, 7), (35, 28)]:
        wa.feature_shape_hw = (h, w)
        wa.initialize_mask_for_cyclic_shift()
        out[f"idx/mask_{h}x{w}"] = wa.mask_for_cyclic_shift
        img = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w)
        rolled = torch.roll(img, shifts=(-3, -3), dims=(2, 3))
        out[f"idx/partition_shifted_{h}x{w}"] = wa.rearrange_1(rolled).squeeze(-1).to(torch.int32)
        out[f"idx/partition_{h}x{w}"] = wa.rearrange_1(img).squeeze(-1).to(torch.int32)
    np.savez_compressed(os.path.join(GOLD, "blocks_patch_pad.npz"), **_np(out))


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    with ReferenceModules() as ref:
        gen_state_dict_keys(ref)
        gen_window_attention(ref)
        gen_blocks_and_patch(ref)
        gen_model_outputs(ref)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
