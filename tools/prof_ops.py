"""Run single operators at the BASELINE config-2 stage shapes (B=64, 256x256) for ncu captures.

    python tools/prof_ops.py wa 0 [shift] [cross]     # window attention at stage 0..4
    python tools/prof_ops.py mlp 0                     # MLP at stage 0..4
    python tools/prof_ops.py wa 0 shift bwd            # forward + backward of the operator at the training batch (B=32)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402

import swinfuse  # noqa: E402

STAGES = [(24, 133, 3, 96), (48, 70, 6, 192), (96, 35, 12, 384), (192, 21, 24, 768), (384, 14, 48, 1536)]


def main():
    what, stage = sys.argv[1], int(sys.argv[2])
    shift, cross = "shift" in sys.argv, "cross" in sys.argv
    prec = "fp32" if "fp32" in sys.argv else "bf16"
    reps = 3
    c, hp, d, hid = STAGES[stage]
    bwd = "bwd" in sys.argv
    b, nh = (32 if bwd else 64), 8
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = r(b, c, hp, hp).contiguous(memory_format=torch.channels_last)
    y = r(b, c, hp, hp).contiguous(memory_format=torch.channels_last)
    ops = swinfuse.ops
    ln = (1 + 0.1 * r(c), 0.1 * r(c))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    if bwd:
        x.requires_grad_(True)
    with torch.set_grad_enabled(bwd):
        for i in range(reps + 1):
            ev[i].record()
            if i == reps:
                break
            if what == "wa":
                w = lambda: r(nh * d, c) * c ** -0.5
                out = ops.window_attention(x, y if cross else None, wq=w(), bq=0.1 * r(nh * d), wk=w(), bk=0.1 * r(nh * d),
                                           wv=w(), bv=0.1 * r(nh * d), wo=r(c, nh * d) * c ** -0.5, bo=0.1 * r(c),
                                           bias_table=r(13, 13), num_heads=nh, head_dim=d, window_size=(7, 7), shift=shift,
                                           ln_q=ln, ln_kv=ln, residual=x, precision=prec)
            else:
                out = ops.mlp(x, w1=(r(hid, c, 1, 1) * c ** -0.5), b1=0.1 * r(hid), w2=(r(c, hid, 1, 1) * hid ** -0.5),
                              b2=0.1 * r(c), ln=ln, residual=x, precision=prec)
            if bwd:
                out.backward(torch.ones_like(out) * 1e-6)
                x.grad = None
    torch.cuda.synchronize()
    print(what, "stage", stage, "shift", shift, "cross", cross, prec, "ms per call:",
          [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(reps)], float(out.abs().mean()))


if __name__ == "__main__":
    main()
