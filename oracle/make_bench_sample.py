"""Writes tests/golden/bench_sample.npz -- TEST INFRASTRUCTURE (run here, committed with its output).

bench.py must prove parity before it times anything but may not execute the oracle on its
product arm.  This script runs the pinned CPU oracle (oracle/fusion_oracle.py, bit-identical to the reference:
tests/test_oracle_golden.py) ONCE on the first image pair of bench.py's rank-0 batch (seed 1000, 256x256, synthetic
state dict) and stores the fused image; bench.py compares its own result for that pair with the stored one.

    python -m oracle.make_bench_sample
"""
import os

import numpy as np
import torch

from oracle import fusion_oracle as fo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_sample.npz")


def bench_inputs(batch: int, size: int, rank: int = 0):
    """The host batches bench.py draws (CPU generator: identical on every machine with this torch build)."""
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    return torch.rand(batch, 1, size, size, generator=g), torch.rand(batch, 1, size, size, generator=g)


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for batch, size in ((64, 256), (32, 256), (1, 1024)):
        ir, vis = bench_inputs(batch, size)
        with torch.no_grad():
            ref = fo.model_forward(fo.synth_state_dict(), ir[:1], vis[:1])
        # 1024x1024: the bottom-right 256x256 corner (where the a006 reflect padding and the shift masks act) keeps the file small
        out[f"fused_b{batch}_s{size}"] = ref.numpy()[..., -256:, -256:].astype(np.float32)
        out[f"ir00_b{batch}_s{size}"] = ir[0, 0, 0, :8].numpy()   # guards against a different RNG stream
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()}, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
