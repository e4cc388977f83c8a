"""A/B check of the two ln_to_tiled kernels (SWINFUSE_LN_TILED_BULK=0|1): MLP operator forward + backward, bit comparison.
    python tools/ab_ln_tiled.py run out.pt   |   python tools/ab_ln_tiled.py cmp a.pt b.pt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402


def run(path):
    import swinfuse
    ops = swinfuse.ops
    res = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    for c, hp in [(96, 7), (96, 35), (192, 7), (192, 21), (384, 7), (384, 14)]:
        for b in (1, 3):
            x = r(b, c, hp, hp).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            w1 = torch.nn.Parameter(r(4 * c, c, 1, 1) * c ** -0.5)
            w2 = torch.nn.Parameter(r(c, 4 * c, 1, 1) * (4 * c) ** -0.5)
            b1, b2 = torch.nn.Parameter(0.1 * r(4 * c)), torch.nn.Parameter(0.1 * r(c))
            lg, lb = torch.nn.Parameter(1 + 0.1 * r(c)), torch.nn.Parameter(0.1 * r(c))
            y = ops.mlp(x, w1=w1, b1=b1, w2=w2, b2=b2, ln=(lg, lb), residual=x, precision="bf16")
            (y * r(*y.shape)).sum().backward()
            key = f"c{c}_hp{hp}_b{b}"
            res[key + "_y"] = y.detach().cpu()
            for n, t in [("gx", x), ("gw1", w1), ("gw2", w2), ("gb1", b1), ("gb2", b2), ("glg", lg), ("glb", lb)]:
                res[key + "_" + n] = t.grad.detach().cpu()
    torch.save(res, path)


def cmp(a, b):
    A, B = torch.load(a), torch.load(b)
    for k in A:
        d = (A[k].double() - B[k].double()).abs().max().item()
        print(k, "equal" if torch.equal(A[k], B[k]) else f"DIFF max {d:.3e} / scale {A[k].abs().max().item():.3e}")


if __name__ == "__main__":
    run(sys.argv[2]) if sys.argv[1] == "run" else cmp(sys.argv[2], sys.argv[3])
