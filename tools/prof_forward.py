"""One forward (B=64, 256x256, bf16) or one training step (B=32) of the drop-in model between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` captures of every kernel of the step at the bench shapes.
    python tools/prof_forward.py [train] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "swin-unet-image-fusion_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    train = "train" in sys.argv
    nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
    b = nums[0] if nums else (32 if train else 64)
    model, swinfuse = bench.build_model("bf16")
    swinfuse.ops.set_dual_streams(False)
    ir, vis = (t.cuda() for t in bench.host_inputs(b, 256, 0))
    if train:
        from swinfuse.loss_ops import FusionLoss
        from swinfuse.train import DataParallelTrainer
        model.train()
        tr = DataParallelTrainer(model, FusionLoss(clamp01=True).cuda(), lr=1e-2, use_graph=False)
        for _ in range(2):
            tr.step(ir, vis)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        tr.step(ir, vis)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        with torch.no_grad():
            for _ in range(2):
                model(ir, vis)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            model(ir, vis)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    print("profiled one", "training step" if train else "forward", "at B =", b, "launches", swinfuse.ops.launch_count())


if __name__ == "__main__":
    main()
