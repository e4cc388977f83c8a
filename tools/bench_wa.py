"""BASELINE configs[4]: isolated window-attention microbenchmark at the five model shapes (B=64, 256x256 input) x
{self, cross} x {plain, shifted}: device time of one fused operator call (LayerNorm + q|k|v GEMM + attention core +
projection + residual), per kernel, against the algorithmic FLOPs 8NC^2 + 196NC and bytes.
    python tools/bench_wa.py > profiles/<tag>_wa_microbench.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402

import swinfuse  # noqa: E402

STAGES = [(24, 133, 3), (48, 70, 6), (96, 35, 12), (192, 21, 24), (384, 14, 48)]
PEAK_TF, PEAK_GBS = 1389.5, 6551.0
if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_GBS = pk["bf16_tflops_sustained"], pk["hbm_gbs"]


def main():
    ops = swinfuse.ops
    b, nh, reps = 64, 8, 5
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    rows = []
    for (c, hp, d) in STAGES:
        x = r(b, c, hp, hp).contiguous(memory_format=torch.channels_last)
        y = r(b, c, hp, hp).contiguous(memory_format=torch.channels_last)
        ln = (1 + 0.1 * r(c), 0.1 * r(c))
        w = lambda: torch.nn.Parameter(r(nh * d, c) * c ** -0.5)
        P = dict(wq=w(), bq=0.1 * r(nh * d), wk=w(), bk=0.1 * r(nh * d), wv=w(), bv=0.1 * r(nh * d),
                 wo=torch.nn.Parameter(r(c, nh * d) * c ** -0.5), bo=0.1 * r(c), bias_table=r(13, 13))
        for cross in (False, True):
            for shift in (False, True):
                call = lambda: ops.window_attention(x, y if cross else None, num_heads=nh, head_dim=d, window_size=(7, 7), shift=shift,
                                                    ln_q=ln, ln_kv=ln, residual=x, precision="bf16", **P)
                with torch.no_grad():
                    for _ in range(2):
                        call()          # warm-up: packs the weights once (cached on the parameter objects)
                    torch.cuda.synchronize()
                    ops.profile_enable(True)
                    for _ in range(reps):
                        call()
                    torch.cuda.synchronize()
                    prof = ops.profile_summary()
                    ops.profile_enable(False)
                n_tok = b * hp * hp
                flops = 8.0 * n_tok * c * c + 196.0 * n_tok * c
                kern = {k: round(v["total_ms"] / reps * 1e3, 1) for k, v in prof.items() if not k.startswith("pack")}
                us = sum(kern.values())
                core = [v for k, v in kern.items() if k.startswith("attn_core")]
                rows.append({"C": c, "Hp": hp, "head_dim": d, "windows": b * (hp // 7) ** 2, "cross": cross, "shift": shift,
                             "us_per_call": round(us, 1), "tflops": round(flops / us / 1e6, 2),
                             "pct_of_bf16_tensor_peak": round(100 * flops / us / 1e6 / PEAK_TF, 2),
                             "attn_core_us": core[0] if core else None, "kernels_us": kern})
    print(json.dumps({"config": "BASELINE configs[4]: window attention microbench, B=64, 7x7 windows, 8 heads, bf16 path",
                      "peak_tflops_sustained": PEAK_TF, "peak_hbm_gbs": PEAK_GBS, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
