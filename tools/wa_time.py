"""Device time of the window-attention operator at a model stage (B=64, 256x256 shapes), CUDA events, 20 calls.
    python tools/wa_time.py [stage ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402

import swinfuse  # noqa: E402

STAGES = [(24, 133, 3), (48, 70, 6), (96, 35, 12), (192, 21, 24), (384, 14, 48)]


def main():
    ops = swinfuse.ops
    stages = [int(a) for a in sys.argv[1:]] or [0, 1]
    b, nh, reps = 64, 8, 20
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    for st in stages:
        c, hp, d = STAGES[st]
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
                    for _ in range(3):
                        call()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(reps):
                        call()
                    e1.record()
                    torch.cuda.synchronize()
                print(f"stage {st} C={c} cross={int(cross)} shift={int(shift)}: {e0.elapsed_time(e1) / reps * 1e3:8.1f} us/call", flush=True)


if __name__ == "__main__":
    main()
