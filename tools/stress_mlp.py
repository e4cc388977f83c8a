"""Determinism / race stress for the fused MLP: same inputs many times, outputs must be bit-identical and correct."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch
import swinfuse
from oracle import fusion_oracle as fo
ops = swinfuse.ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for (m, c, hidden) in [(148 * 128 * 2 + 5, 48, 192), (148 * 128 * 3 + 77, 24, 96), (148 * 128 * 5 + 1, 64, 256), (1000, 48, 192), (64 * 70 * 70, 48, 192), (64 * 133 * 133, 24, 96)]:
    g = torch.Generator().manual_seed(m + c)
    x = torch.randn(1, c, 1, m, generator=g)
    w1, b1 = torch.randn(hidden, c, 1, 1, generator=g) * (2 / c) ** 0.5, 0.1 * torch.randn(hidden, generator=g)
    w2, b2 = torch.randn(c, hidden, 1, 1, generator=g) * (2 / hidden) ** 0.5, 0.1 * torch.randn(c, generator=g)
    lg, lb = 1 + 0.2 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    nx = fo.layer_norm_c(x, lg, lb)
    ref = x + torch.nn.functional.conv2d(torch.nn.functional.elu(torch.nn.functional.conv2d(nx, w1, b1)), w2, b2)
    xc = x.cuda()
    args = dict(w1=w1.cuda(), b1=b1.cuda(), w2=w2.cuda(), b2=b2.cuda(), ln=(lg.cuda(), lb.cuda()))
    first = None; bad = 0; worst = 0.0
    for i in range(reps):
        got = ops.mlp(xc, residual=xc, precision="bf16", **args)
        torch.cuda.synchronize()
        if first is None:
            first = got.clone()
        elif not torch.equal(got, first):
            bad += 1
            d = (got - first).abs()
            idx = d.flatten().argmax().item()
            if bad <= 3:
                nz = (d.flatten() > 0).nonzero().flatten()
                print("   mismatch run", i, "n_diff", nz.numel(), "first idx", nz[:4].tolist(), "max", float(d.max()), "row of max", idx % m, "ch", idx // m)
        e = float((got.cpu() - ref).abs().max() / ref.abs().max())
        worst = max(worst, e)
    print(f"M={m} C={c} H={hidden}: nondeterministic runs {bad}/{reps - 1}, worst rel err {worst:.3e}")
