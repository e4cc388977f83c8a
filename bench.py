#!/usr/bin/env python
"""bench.py -- headline benchmark of the Swin-UNet fusion hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--batch 64] [--size 256] [--precision fp32|bf16] [--mode infer]

A "step" is one pass of the hot path (MyModel.forward, a013:209-230) over one batch of
synthetic image pairs.  At N=1 the workload is BASELINE.json configs[1]: inference, batch 64,
256x256 pairs, one B200.  For N>1 (launched by torch.distributed.run, one rank per GPU) every
rank runs the same per-GPU batch on its own pairs -- image pairs are independent, there is no
data-path collective -- and `value` is all pairs processed / max-over-ranks device time ("weak").

One JSON line is printed by rank 0 (see the task contract): value (inputs resident in HBM,
CUDA-graph replay), e2e (host pinned buffers -> H2D -> forward -> D2H through the drop-in
MyModel API), roofline of the dominant kernel measured live with CUDA events
(sf_profile_enable), cpu_baseline (the oracle port of the reference's CPU path on the host
cores), clocks sampled with nvidia-smi during the timed region, gpu_launches.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port,
oracle/fusion_oracle.py: same ops in the same order as the reference's PyTorch code) on the
host cores with all threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "swin-unet-image-fusion_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "fused image pairs/sec"
UNIT = "pairs/s"
GFLOP_PER_PAIR_256 = 28.120  # SURVEY.md appendix B (forward, MAC = 2 FLOP)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE configs[1] (default, headline); train: configs[2] training step")
    ap.add_argument("--batch", type=int, default=None, help="image pairs per GPU per step (64 infer / 32 train)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("SWINFUSE_BENCH_PRECISION", "auto"))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=4, help="pairs in the cpu_baseline sample (0 = skip)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load": drop idle samples far below the maximum seen
        load = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_forward_pairs_per_s(n_pairs: int, size: int, reps: int, warmup: int):
    from oracle import fusion_oracle as fo
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = fo.synth_state_dict()
    ir, vis = fo.synth_inputs(n_pairs, size, size)
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            fo.model_forward(sd, ir, vis)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    best, mean = min(times), sum(times) / len(times)
    return dict(best_s=best, mean_s=mean, threads=torch.get_num_threads(), pairs=n_pairs)


def cpu_train_pairs_per_s(n_pairs: int, size: int, reps: int, warmup: int):
    """One training step of the reference path on the host cores: oracle forward (train mode: batch-stat BatchNorm,
    a013:133) -> clamp (a016:153) -> a008 loss (dense restatement of the kornia formulation) -> autograd backward.
    No optimizer step (a few ms on the CPU).  SURVEY 8(d): B=4, nn.ELU() semantics."""
    from oracle import fusion_oracle as fo
    from oracle import kornia_restatement as kr
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "num_batches" not in k else v.clone())
          for k, v in fo.synth_state_dict().items()}
    ir, vis = fo.synth_inputs(n_pairs, size, size)
    ms, sobel = kr.MS_SSIMLoss(), kr.Sobel()
    leaves = [v for v in sd.values() if v.requires_grad]
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        fusion = torch.clamp(fo.model_forward(sd, ir, vis, training=True), 0, 1)
        loss = kr.total_loss(fusion, ir, vis, ms, sobel)
        torch.autograd.grad(loss, leaves, allow_unused=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return dict(best_s=min(times), mean_s=sum(times) / len(times), threads=torch.get_num_threads(), pairs=n_pairs)


def run_reference_train(args):
    n = max(1, min(args.cpu_sample if args.cpu_sample > 0 else 4, args.batch))
    r = cpu_train_pairs_per_s(n, args.size, reps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    value = n / r["mean_s"]
    sample = (f"{n} of {args.batch} pairs per step ({args.size}x{args.size}, fp32 oracle forward + dense a008 loss restatement + "
              f"autograd backward, torch CPU ops)")
    line = {"impl": "reference", "metric": "training image pairs/sec", "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 3)), "warmup": min(args.warmup, 1), "ms_per_step": r["mean_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"training step B={args.batch}/GPU {args.size}x{args.size} pairs (BASELINE configs[2])", "step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_reference(args, rank: int):
    if rank != 0:
        return
    if args.mode == "train":
        return run_reference_train(args)
    n = max(1, min(args.cpu_sample if args.cpu_sample > 0 else 4, args.batch))
    r = cpu_forward_pairs_per_s(n, args.size, reps=max(1, args.steps), warmup=min(args.warmup, 1))
    value = n / r["mean_s"]
    sample = f"{n} of {args.batch} pairs per step ({args.size}x{args.size}, fp32 eval forward, torch CPU ops)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": r["mean_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"inference B={args.batch}/GPU {args.size}x{args.size} pairs (BASELINE configs[1])",
                       "step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_model(precision: str):
    from torch import nn
    import swinfuse
    from oracle import fusion_oracle as fo  # synthetic weights only (deterministic state_dict)
    swinfuse.install_dropin()
    swinfuse.set_default_precision(precision)
    from a013_ModelDefinition import MyModel
    cfg = fo.FusionConfig()
    m = MyModel(window_size=cfg.window_size, merging_size=cfg.merging_size, in_dims_list=cfg.in_dims_list,
                out_dims_list=cfg.out_dims_list, att_num_heads=cfg.att_num_heads,
                att_dims_per_head_ratio=cfg.att_dims_per_head_ratio, attention_drop_ratio=0,
                linear_after_att_drop_ratio=0, mlp_hidden_dims_ratio=cfg.mlp_hidden_dims_ratio,
                mlp_activation_func=nn.ELU(inplace=True), mlp_drop_ratio=0, final_layer_att_dims_per_head_ratio=1,
                final_conv_layer_kernel_size=3, final_layer_mlp_hidden_dims_ratio=1).cuda().eval()
    m.load_state_dict(fo.synth_state_dict(cfg), strict=True)
    return m, swinfuse


def pick_precision(arg: str) -> str:
    if arg in ("fp32", "bf16"):
        return arg
    # auto: bf16 tensor-core path when the library has it, else fp32
    import swinfuse
    return "bf16" if getattr(swinfuse, "BF16_READY", False) else "fp32"


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    precision = pick_precision(args.precision)
    model, swinfuse = build_model(precision)
    ops = swinfuse.ops
    B, S = args.batch, args.size
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    h_ir = torch.rand(B, 1, S, S, generator=g).pin_memory()
    h_vis = torch.rand(B, 1, S, S, generator=g).pin_memory()
    h_out = torch.empty(B, 1, S, S).pin_memory()
    d_ir, d_vis = h_ir.cuda(non_blocking=True), h_vis.cuda(non_blocking=True)
    stream = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad(), torch.cuda.stream(stream):
        # ---- eager warm-up (also the first-call module checks) + launch count per forward ---------
        d_out = model(d_ir, d_vis)           # first call: one-time weight packing, module input checks
        ops.reset_launch_count()
        d_out = model(d_ir, d_vis)
        launches_per_fwd = ops.launch_count()   # steady-state kernels per forward (what the CUDA graph replays)
        for _ in range(max(0, args.warmup - 2)):
            d_out = model(d_ir, d_vis)
        stream.synchronize()

        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                g_out = model(d_ir, d_vis)
            for _ in range(2):
                graph.replay()
            stream.synchronize()

        def step():
            if graph is not None:
                graph.replay()
                return g_out
            return model(d_ir, d_vis)

        # ---- value: device-resident inputs, K steps, CUDA events on the launching stream ----------
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        stream.synchronize()
        barrier()
        ms_total = e0.elapsed_time(e1)

        # ---- e2e: pinned host buffers -> H2D -> forward (public MyModel API) -> D2H, every step ----
        def e2e_step():
            d_ir.copy_(h_ir, non_blocking=True)
            d_vis.copy_(h_vis, non_blocking=True)
            out = step()
            h_out.copy_(out, non_blocking=True)

        for _ in range(2):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        f1.record(stream)
        stream.synchronize()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        clocks = sampler.stop() if sampler else None

        # ---- attribution pass: per-kernel CUDA-event timing inside the library (eager, same stream) --
        prof = {}
        if rank == 0:
            ops.set_dual_streams(False)   # one kernel at a time: clean per-kernel durations for the roofline numbers
            ops.profile_enable(True)
            nprof = min(3, args.steps)
            for _ in range(nprof):
                model(d_ir, d_vis)
            stream.synchronize()
            prof = ops.profile_summary()
            ops.profile_enable(False)
            ops.set_dual_streams(True)
            for v in prof.values():
                v["steps"] = nprof

    # ---- max over ranks ------------------------------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    pk = peaks()
    pairs = B * world * args.steps
    value = pairs / (ms_total / 1e3)
    e2e = pairs / (ms_e2e / 1e3)

    # dominant kernel by summed device time
    roofline = None
    kernels = {}
    if prof:
        tot = sum(v["total_ms"] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
            kernels[k] = {"launches_per_step": v["launches"] // v["steps"], "ms_per_step": v["total_ms"] / v["steps"],
                          "share": v["total_ms"] / tot if tot else 0.0,
                          "tflops": v["flops"] / (v["total_ms"] * 1e-3) / 1e12 if v["total_ms"] else 0.0,
                          "gbs": v["bytes"] / (v["total_ms"] * 1e-3) / 1e9 if v["total_ms"] else 0.0}
        top, tv = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
        avg_s = tv["total_ms"] * 1e-3 / tv["launches"]
        ai = tv["flops"] / max(tv["bytes"], 1.0)
        ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        if ai >= ridge or "gemm" in top or "attn" in top or "mlp" in top:
            ach = tv["flops"] / tv["launches"] / avg_s / 1e12
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                        "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                        "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)"}
        else:
            ach = tv["bytes"] / tv["launches"] / avg_s / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"]}
        roofline["avg_launch_us"] = avg_s * 1e6
        roofline["share_of_step"] = tv["total_ms"] / tot if tot else None
        # the other roof, for reference: algorithmic HBM bytes of the same launches
        roofline["hbm_gbs_algorithmic"] = tv["bytes"] / tv["launches"] / avg_s / 1e9
        roofline["hbm_frac_algorithmic"] = roofline["hbm_gbs_algorithmic"] / pk["hbm_gbs"]
        # measured DRAM traffic per launch of this kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
        # `ncu --set full` capture, committed under profiles/): far above the algorithmic bytes = wasted re-reads
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if top in tj:
                roofline["traffic"] = tj[top]["dram_bytes_per_launch"]
                roofline["traffic_source"] = tj[top].get("source")
                roofline["algorithmic_bytes_per_launch"] = tv["bytes"] / tv["launches"]

    cpu = None
    if world == 1 and args.cpu_sample > 0:
        r = cpu_forward_pairs_per_s(args.cpu_sample, S, reps=2, warmup=1)
        cpu = {"value": r["pairs"] / r["mean_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{r['pairs']} of {B} pairs, {S}x{S}, fp32 eval forward of the oracle port (torch CPU ops, "
                         f"same op sequence as the reference), mean of 2 after 1 warm-up"}

    model_tf = value * GFLOP_PER_PAIR_256 * (S / 256.0) ** 2 / 1e3
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": f"inference B={B}/GPU {S}x{S} IR+visible pairs, default A000_CONFIG Swin-UNet "
                                   f"(BASELINE configs[1])", "global_batch": B * world, "precision": precision,
                       "launch": ("cuda-graph replay" if graph is not None else "eager") + ", IR / visible paths on two streams",
                       "l2": "per-step activation working set (>2 GB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "synthetic deterministic state_dict (oracle.synth_state_dict)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_fwd * args.steps,
            "model_tflops": model_tf, "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu}
    emit(line)


def run_train(args, rank: int, world: int, local_rank: int):
    """BASELINE configs[2]: training step = forward -> clamp -> loss -> backward -> gradient all-reduce -> Adam,
    batch 32 pairs per GPU, data parallel.  Forward in `precision`, backward kernels fp32."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    precision = pick_precision(args.precision)
    model, swinfuse = build_model(precision)
    model.train()
    from swinfuse.loss_ops import FusionLoss
    from swinfuse.train import DataParallelTrainer
    loss_fn = FusionLoss(clamp01=True).to(dev)   # a016:153's clamp is folded into the loss kernels
    trainer = DataParallelTrainer(model, loss_fn, lr=1e-2, use_graph=not args.no_graph)
    B, S = args.batch, args.size
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    ir = torch.rand(B, 1, S, S, generator=g).to(dev)
    vis = torch.rand(B, 1, S, S, generator=g).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ops = swinfuse.ops
    ops.reset_launch_count()
    loss0 = trainer.step(ir, vis)
    launches = ops.launch_count()
    for _ in range(max(2, args.warmup - 1)):   # the third step captures the CUDA graph of forward + loss + backward
        trainer.step(ir, vis)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = trainer.step(ir, vis)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    prof = {}
    if rank == 0:
        ops.profile_enable(True)
    trainer.use_graph = False   # attribution pass runs eagerly (CUDA-event brackets around every launch)
    trainer.step(ir, vis)   # every rank steps (the step contains the gradient all-reduce)
    torch.cuda.synchronize()
    if rank == 0:
        prof = ops.profile_summary()
        ops.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank != 0:
        return
    pairs = B * world * args.steps
    cpu = None
    if world == 1 and args.cpu_sample > 0:   # the reference path's training step on this box's host cores, bounded sample
        n = max(1, min(args.cpu_sample, B))
        r = cpu_train_pairs_per_s(n, S, reps=1, warmup=1)
        cpu = {"value": n / r["mean_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{n} of {B} pairs, {S}x{S}: fp32 oracle forward (train mode) + dense a008 loss restatement + autograd backward, "
                         f"torch CPU ops, 1 step after 1 warm-up"}
    kernels = {k: {"launches_per_step": v["launches"], "ms_per_step": v["total_ms"]} for k, v in
               sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])}
    line = {"metric": "training image pairs/sec", "value": pairs / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "steps_per_s": args.steps / (ms / 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision + " fwd / tf32 + fp16 tensor-core bwd, fp32 accumulation",
            "data": "synthetic",
            "config": {"workload": f"training step B={B}/GPU {S}x{S} pairs, data parallel (BASELINE configs[2])",
                       "global_batch": B * world, "loss": "a008 loss + gradient in libswinfuse kernels (sf_fusion_loss: separable MS-SSIM+L1, Sobel, intensity; clamp folded in)",
                       "optimizer": "Adam lr 1e-2, sf_adam_step over one flat buffer",
                       "collective": "one NCCL all-reduce of the flat fp32 gradient buffer" if world > 1 else "none",
                       "launch": "eager" if args.no_graph else "cuda-graph replay of zero-grad + forward + loss + backward; all-reduce + Adam eager"},
            "gpu_launches": launches * args.steps, "loss_first": float(loss0), "loss_last": float(loss), "clocks": clocks,
            "model_tflops": pairs / (ms / 1e3) * 3 * GFLOP_PER_PAIR_256 * (S / 256.0) ** 2 / 1e3, "kernels": kernels,
            "cpu_baseline": cpu}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # Libraries (NCCL's version banner, for one) write to fd 1: keep the real stdout for the JSON line only and
    # send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.batch is None:
        args.batch = 64 if args.mode == "infer" else 32
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.mode == "train":
            run_train(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
