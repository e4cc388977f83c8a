"""Small invocations of every tcgen05 / mbarrier kernel for compute-sanitizer (memcheck / racecheck): fused window attention
(both flavours, self + cross, shifted), fused MLP, the streaming GEMMs (C >= 96), MLP and window-attention backward on
tcgen05 (k_tc_gemm2 with transposed images, k_tc_wgrad) and the HMMA attention adjoint.
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402

import swinfuse  # noqa: E402


def main():
    ops = swinfuse.ops
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    nh = 8
    for (c, d, b, hp, wp) in [(24, 3, 2, 21, 28), (48, 6, 1, 21, 14), (96, 12, 1, 14, 14), (192, 24, 1, 7, 14)]:
        for cross, shift in ((False, True), (True, True), (False, False)):
            x = r(b, c, hp, wp).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            y = r(b, c, hp, wp).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            ln = (torch.nn.Parameter(1 + 0.1 * r(c)), torch.nn.Parameter(0.1 * r(c)))
            w = lambda *s: torch.nn.Parameter(r(*s) * s[-1] ** -0.5)
            P = dict(wq=w(nh * d, c), bq=torch.nn.Parameter(0.1 * r(nh * d)), wk=w(nh * d, c), bk=torch.nn.Parameter(0.1 * r(nh * d)),
                     wv=w(nh * d, c), bv=torch.nn.Parameter(0.1 * r(nh * d)), wo=w(c, nh * d), bo=torch.nn.Parameter(0.1 * r(c)),
                     bias_table=torch.nn.Parameter(r(13, 13)))
            out = ops.window_attention(x, y if cross else None, num_heads=nh, head_dim=d, window_size=(7, 7), shift=shift, ln_q=ln, ln_kv=ln,
                                       residual=x, precision="bf16", **P)
            out.backward(torch.ones_like(out) * 1e-3)
            hid = 4 * c
            out = ops.mlp(x, w1=w(hid, c, 1, 1), b1=torch.nn.Parameter(0.1 * r(hid)), w2=torch.nn.Parameter(r(c, hid, 1, 1) * hid ** -0.5),
                          b2=torch.nn.Parameter(0.1 * r(c)), ln=ln, residual=x, precision="bf16")
            out.backward(torch.ones_like(out) * 1e-3)
            torch.cuda.synchronize()
            print("ok", c, d, "cross" if cross else "self", "shift" if shift else "plain", float(out.abs().mean()), flush=True)


if __name__ == "__main__":
    main()
