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
MyModel API, copies on a second stream), roofline of the dominant kernel measured live with
CUDA events (sf_profile_enable; the bound is chosen by arithmetic intensity against the ridge
and both fractions are printed), cpu_baseline (the oracle port of the reference's CPU path on
the host cores), clocks sampled with nvidia-smi during the timed region, gpu_launches.  Before
anything is timed the fused image of the first pair is checked against a stored oracle sample
(tests/golden/bench_sample.npz, written by oracle/make_bench_sample.py): a wrong result aborts.

The same line carries `training` (BASELINE configs[2]: forward -> clamp -> a008 loss -> backward
-> gradient all-reduce over NCCL -> Adam, batch 32 pairs per GPU, data parallel: the only path
with a collective), `window_attention` (per-stage operator time against min(tensor, HBM) roof)
and, at N=1, `highres_1024` (configs[3]).  `--mode train` prints the training line alone.

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
GFLOP_PER_PAIR = {224: 13.443, 256: 28.120, 1024: 308.48}


def gflop_per_pair(size: int) -> float:
    return GFLOP_PER_PAIR.get(size, GFLOP_PER_PAIR_256 * (size / 256.0) ** 2)


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
    ap.add_argument("--train-steps", type=int, default=None, help="timed steps of the training block (default min(steps, 10); 0 = skip)")
    ap.add_argument("--train-batch", type=int, default=32, help="pairs per GPU per training step (BASELINE configs[2])")
    ap.add_argument("--no-highres", action="store_true", help="skip the 1024x1024 lines (configs[3])")
    ap.add_argument("--no-parity-check", action="store_true", help="do not compare with tests/golden/bench_sample.npz first")
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
    r = cpu_forward_pairs_per_s(n, args.size, reps=max(1, args.steps), warmup=max(0, args.warmup))
    value = n / r["mean_s"]
    sample = (f"{n} of {args.batch} pairs per step ({args.size}x{args.size}, fp32 eval forward of the oracle port of the reference's "
              f"CPU path -- same torch ops in the same order; the Python reference itself cannot travel to the GPU box)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": r["mean_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"inference B={args.batch}/GPU {args.size}x{args.size} IR+visible pairs, default A000_CONFIG Swin-UNet "
                                   f"(BASELINE configs[1])", "global_batch": args.batch * max(1, args.gpus), "precision": "fp32",
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


def host_inputs(batch: int, size: int, rank: int):
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    return torch.rand(batch, 1, size, size, generator=g), torch.rand(batch, 1, size, size, generator=g)


def parity_gate(fused_first_pair: torch.Tensor, h_ir: torch.Tensor, batch: int, size: int, precision: str, rank: int):
    """no number without parity.  The fused image of pair 0 of rank 0's batch against the stored output
    of the pinned CPU oracle for exactly that pair (tests/golden/bench_sample.npz).  Tolerances are north_star's:
    1e-4 relative for fp32, 2e-2 for bf16.  Returns the relative error (None when no sample is stored for this shape)."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "bench_sample.npz")
    key = f"fused_b{batch}_s{size}"
    if rank != 0 or not os.path.isfile(path):
        return None
    z = np.load(path)
    if key not in z.files:
        return None
    if not np.array_equal(z[f"ir00_b{batch}_s{size}"], h_ir[0, 0, 0, :8].numpy()):
        raise SystemExit("bench.py: the host RNG stream differs from the one the stored oracle sample was drawn with")
    ref = torch.from_numpy(z[key])
    got = fused_first_pair.detach().float().cpu()[..., -ref.shape[-2]:, -ref.shape[-1]:]
    err = float((got - ref).abs().max() / ref.abs().max())
    tol = 1e-4 if precision == "fp32" else 2e-2
    if not err <= tol:
        raise SystemExit(f"bench.py: parity check FAILED before timing: fused image of pair 0 differs from the oracle sample "
                         f"by {err:.3e} relative (tolerance {tol:g}, {precision}, B={batch}, {size}x{size})")
    return err


def roofline_block(prof: dict, pk: dict, steps: int):
    """(roofline of the dominant kernel, per-kernel table) from the library's per-launch CUDA-event profile.  The
    bound is chosen by arithmetic intensity against the ridge (achieved / min(compute, HBM) roof); both
    fractions are reported whichever binds."""
    if not prof:
        return None, {}
    tot = sum(v["total_ms"] for v in prof.values())
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
        sec = v["total_ms"] * 1e-3
        tf = v["flops"] / sec / 1e12 if sec else 0.0
        gbs = v["bytes"] / sec / 1e9 if sec else 0.0
        ai = v["flops"] / max(v["bytes"], 1.0)
        kernels[k] = {"launches_per_step": v["launches"] // steps, "ms_per_step": v["total_ms"] / steps,
                      "share": v["total_ms"] / tot if tot else 0.0, "tflops": tf, "gbs": gbs,
                      "bound": "tensor" if ai >= ridge else "hbm",
                      "frac": (tf / pk["bf16_tflops_sustained"]) if ai >= ridge else (gbs / pk["hbm_gbs"])}
    top, tv = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
    avg_s = tv["total_ms"] * 1e-3 / tv["launches"]
    ai = tv["flops"] / max(tv["bytes"], 1.0)
    tf = tv["flops"] / tv["launches"] / avg_s / 1e12
    gbs = tv["bytes"] / tv["launches"] / avg_s / 1e9
    if ai >= ridge:
        roofline = {"kernel": top, "bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": tf / pk["bf16_tflops_sustained"],
                    "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)"}
    else:
        roofline = {"kernel": top, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / pk["hbm_gbs"], "peak_source": pk["source"]}
    roofline.update({"traffic": None, "avg_launch_us": avg_s * 1e6, "share_of_step": tv["total_ms"] / tot if tot else None,
                     "arithmetic_intensity": ai, "ridge": ridge,
                     "tensor_tflops": tf, "tensor_frac": tf / pk["bf16_tflops_sustained"],
                     "hbm_gbs_algorithmic": gbs, "hbm_frac_algorithmic": gbs / pk["hbm_gbs"],
                     "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                     "algorithmic_flops_per_launch": tv["flops"] / tv["launches"]})
    # measured DRAM traffic per launch of this kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
    # `ncu --set full` capture, committed under profiles/): far above the algorithmic bytes = wasted re-reads
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if top in tj:
            roofline["traffic"] = tj[top]["dram_bytes_per_launch"]
            roofline["traffic_source"] = tj[top].get("source")
    return roofline, kernels


# window-attention operator (a001:448-474 + a004:29-38) per stage at 256x256: (C, tokens per image and path)
WA_STAGES_256 = ((24, 17689), (48, 4900), (96, 1225), (192, 441), (384, 196))
WA_KERNEL_PREFIXES = ("wa_fused", "tc_gemm_qkv", "tc_gemm_q_", "tc_gemm_kv", "attn_core", "tc_gemm_proj")


def window_attention_block(kernels: dict, batch: int, size: int, pk: dict):
    """Per stage: device time of ONE window-attention operator call (all of its kernels, from the attribution pass)
    against SURVEY 8(d)'s roof min(compute, HBM) for the ideal fusion: FLOPs = 8NC^2 + 196NC, bytes = fp32 rows read once
    (q source [+ k/v source in cross blocks]) and written once.  16 calls per stage and step: 8 self + 8 cross."""
    if size != 256 or not kernels:
        return None
    out = {}
    for c, n in WA_STAGES_256:
        ms = 0.0
        names = []
        for k, v in kernels.items():
            if k.endswith(f"_c{c}") and k.startswith(WA_KERNEL_PREFIXES):
                ms += v["ms_per_step"]
                names.append(k)
        pre = kernels.get(f"ln_to_tiled_c{c}")
        if pre:   # LayerNorm pre-pass of the wide stages: 24 of its 40 launches per step feed window attention (16 feed the MLP)
            ms += pre["ms_per_step"] * 24.0 / 40.0
            names.append(f"ln_to_tiled_c{c} (24/40)")
        if ms == 0.0:
            continue
        tok = float(batch) * n
        flops = 8.0 * tok * c * c + 196.0 * tok * c
        byts = 4.0 * tok * c * 2.5            # 8 self calls move 2 maps, 8 cross calls 3
        us = ms * 1e3 / 16.0
        t_tensor = flops / (pk["bf16_tflops_sustained"] * 1e12) * 1e6
        t_hbm = byts / (pk["hbm_gbs"] * 1e9) * 1e6
        out[f"c{c}"] = {"us_per_call": us, "kernels": names, "tflops": flops / (us * 1e-6) / 1e12,
                        "tensor_frac": t_tensor / us, "hbm_frac_ideal_fusion": t_hbm / us,
                        "bound": "tensor" if t_tensor >= t_hbm else "hbm", "frac_of_roof": max(t_tensor, t_hbm) / us}
    return out


def time_forward(model, ops, B, S, steps, warmup, use_graph, stream, rank, world, barrier, precision, check_parity):
    """value / e2e timing of the inference forward at one (B, S); returns a dict of raw measurements."""
    h_ir, h_vis = host_inputs(B, S, rank)
    h_ir, h_vis = h_ir.pin_memory(), h_vis.pin_memory()
    h_out = torch.empty(B, 1, S, S).pin_memory()
    res = {}
    with torch.no_grad(), torch.cuda.stream(stream):
        # uploads on the stream the forward runs on (torch streams do not synchronise with the default stream)
        d_ir, d_vis = h_ir.cuda(non_blocking=True), h_vis.cuda(non_blocking=True)
        # ---- eager warm-up (also the first-call module checks) + launch count per forward ---------
        d_out = model(d_ir, d_vis)           # first call: one-time weight packing, module input checks
        if check_parity:
            stream.synchronize()
            res["parity_rel_err"] = parity_gate(d_out[:1], h_ir, B, S, precision, rank)
        ops.reset_launch_count()
        d_out = model(d_ir, d_vis)
        res["launches_per_fwd"] = ops.launch_count()   # steady-state kernels per forward (what the CUDA graph replays)
        for _ in range(max(0, warmup - 2)):
            d_out = model(d_ir, d_vis)
        stream.synchronize()

        graph = None
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                g_out = model(d_ir, d_vis)
            for _ in range(2):
                graph.replay()
            stream.synchronize()
            if check_parity:   # the replayed graph computes the same image as the eager call that passed the gate
                assert torch.equal(g_out[:1], d_out[:1]) or float((g_out[:1] - d_out[:1]).abs().max()) <= 1e-6

        def step():
            if graph is not None:
                graph.replay()
                return g_out
            return model(d_ir, d_vis)

        # ---- value: device-resident inputs, K steps, CUDA events on the launching stream ----------
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        stream.synchronize()
        barrier()
        res["ms_total"] = e0.elapsed_time(e1)

        # ---- e2e: pinned host buffers -> H2D -> forward (public MyModel API) -> D2H, every step.  Copies run on a
        # second stream into / out of staging buffers, so step i+1's upload and step i-1's download overlap step i's
        # forward; the compute stream only pays two device-to-device copies.  All of it is inside the timed region.
        copy = torch.cuda.Stream()
        s_in = [(torch.empty_like(d_ir), torch.empty_like(d_vis)) for _ in range(2)]
        s_out = [torch.empty(B, 1, S, S, device=d_ir.device) for _ in range(2)]
        up = [torch.cuda.Event() for _ in range(2)]
        used = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        down = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            with torch.cuda.stream(copy):
                copy.wait_event(used[i & 1])             # the forward that read this staging pair has consumed it
                s_in[i & 1][0].copy_(h_ir, non_blocking=True)
                s_in[i & 1][1].copy_(h_vis, non_blocking=True)
                up[i & 1].record(copy)

        def e2e_run(n):
            for ev in used + down:
                ev.record(stream)
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                stream.wait_event(up[i & 1])
                d_ir.copy_(s_in[i & 1][0], non_blocking=True)
                d_vis.copy_(s_in[i & 1][1], non_blocking=True)
                used[i & 1].record(stream)
                out = step()
                stream.wait_event(down[i & 1])           # the download of step i-2 has left this staging buffer
                s_out[i & 1].copy_(out, non_blocking=True)
                done[i & 1].record(stream)
                with torch.cuda.stream(copy):
                    copy.wait_event(done[i & 1])
                    h_out.copy_(s_out[i & 1], non_blocking=True)
                    down[i & 1].record(copy)
            stream.wait_stream(copy)                      # the last result is on the host when the closing event fires

        e2e_run(2)
        stream.synchronize()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        e2e_run(steps)
        f1.record(stream)
        stream.synchronize()
        barrier()
        res["ms_e2e"] = f0.elapsed_time(f1)
        if check_parity and rank == 0:   # what arrived on the host is the image that passed the gate
            assert float((h_out[:1] - d_out[:1].cpu()).abs().max()) <= 1e-6
    res.update(graph=graph, d_ir=d_ir, d_vis=d_vis)
    return res


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    precision = pick_precision(args.precision)
    model, swinfuse = build_model(precision)
    ops = swinfuse.ops
    B, S = args.batch, args.size
    stream = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = time_forward(model, ops, B, S, args.steps, args.warmup, not args.no_graph, stream, rank, world, barrier, precision,
                     check_parity=not args.no_parity_check)
    clocks = sampler.stop() if sampler else None
    ms_total, ms_e2e, launches_per_fwd = r["ms_total"], r["ms_e2e"], r["launches_per_fwd"]
    parity_err = r.get("parity_rel_err")

    # ---- attribution pass: per-kernel CUDA-event timing inside the library (eager, same stream) --
    prof, nprof = {}, min(3, args.steps)
    if rank == 0:
        with torch.no_grad(), torch.cuda.stream(stream):
            ops.set_dual_streams(False)   # one kernel at a time: clean per-kernel durations for the roofline numbers
            ops.profile_enable(True)
            for _ in range(nprof):
                model(r["d_ir"], r["d_vis"])
            stream.synchronize()
            prof = ops.profile_summary()
            ops.profile_enable(False)
            ops.set_dual_streams(True)

    # ---- configs[3]: 1024x1024 inference, B = 1 and 4 (rank 0 of a single-GPU run only) ------------------
    highres = None
    if world == 1 and S == 256 and not args.no_highres:
        highres = {}
        for hb in (1, 4):
            hr = time_forward(model, ops, hb, 1024, max(3, min(args.steps, 10)), 3, not args.no_graph, stream, rank, world, barrier,
                              precision, check_parity=(hb == 1 and not args.no_parity_check))
            n = max(3, min(args.steps, 10))
            highres[f"b{hb}"] = {"pairs_per_s": hb * n / (hr["ms_total"] / 1e3), "ms_per_step": hr["ms_total"] / n,
                                 "e2e_pairs_per_s": hb * n / (hr["ms_e2e"] / 1e3),
                                 "model_tflops": hb * n / (hr["ms_total"] / 1e3) * gflop_per_pair(1024) / 1e3,
                                 "parity_rel_err": hr.get("parity_rel_err")}
            del hr
        torch.cuda.empty_cache()

    # ---- max over ranks ------------------------------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    # ---- configs[2]: the training step, same process, same ranks (the only path with a collective) -------
    training = None
    tsteps = min(args.steps, 10) if args.train_steps is None else args.train_steps
    if tsteps > 0 and S == 256:
        del r
        torch.cuda.empty_cache()
        targs = argparse.Namespace(**vars(args))
        targs.batch, targs.steps, targs.warmup = args.train_batch, tsteps, max(3, args.warmup)
        training = measure_training(targs, rank, world, local_rank, precision)
    if rank != 0:
        return
    pk = peaks()
    pairs = B * world * args.steps
    value = pairs / (ms_total / 1e3)
    e2e = pairs / (ms_e2e / 1e3)
    roofline, kernels = roofline_block(prof, pk, nprof)

    cpu = None
    if world == 1 and args.cpu_sample > 0:
        c = cpu_forward_pairs_per_s(args.cpu_sample, S, reps=2, warmup=1)
        cpu = {"value": c["pairs"] / c["mean_s"], "unit": UNIT, "cores": c["threads"], "kind": "port",
               "sample": f"{c['pairs']} of {B} pairs, {S}x{S}, fp32 eval forward of the oracle port (torch CPU ops, same op "
                         f"sequence as the reference; /root/reference itself does not exist on the GPU box), mean of 2 after 1 warm-up"}

    model_tf = value * gflop_per_pair(S) / 1e3
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": f"inference B={B}/GPU {S}x{S} IR+visible pairs, default A000_CONFIG Swin-UNet "
                                   f"(BASELINE configs[1])", "global_batch": B * world, "precision": precision,
                       "launch": ("cuda-graph replay" if not args.no_graph else "eager") + ", IR / visible paths on two streams",
                       "l2": "per-step activation working set (>2 GB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "synthetic deterministic state_dict (oracle.synth_state_dict)",
                       "parity_gate": "fused image of pair 0 vs stored oracle sample before timing (tests/golden/bench_sample.npz)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4,
                    "ms_per_step": ms_e2e / args.steps, "copies": "pinned host <-> device on a second stream, overlapping the forward"},
            "gpu_launches": launches_per_fwd * args.steps, "parity_rel_err": parity_err,
            "model_tflops": model_tf, "model_tensor_frac": model_tf / pk["bf16_tflops_sustained"], "clocks": clocks,
            "roofline": roofline, "window_attention": window_attention_block(kernels, B, S, pk), "kernels": kernels,
            "cpu_baseline": cpu, "training": training, "highres_1024": highres}
    emit(line)


def measure_training(args, rank: int, world: int, local_rank: int, precision: str):
    """BASELINE configs[2]: training step = forward -> clamp -> loss -> backward -> gradient all-reduce -> Adam,
    batch 32 pairs per GPU, data parallel.  Returns the `training` block (rank 0) / None (other ranks)."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    model, swinfuse = build_model(precision)
    model.train()
    from swinfuse.loss_ops import FusionLoss
    from swinfuse.train import DataParallelTrainer
    loss_fn = FusionLoss(clamp01=True).to(dev)   # a016:153's clamp is folded into the loss kernels
    trainer = DataParallelTrainer(model, loss_fn, lr=1e-2, use_graph=not args.no_graph)
    B, S = args.batch, args.size
    ir, vis = (t.to(dev) for t in host_inputs(B, S, rank))

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
    # the exchange step alone: the flat-gradient all-reduce, timed on the device (every rank takes part)
    ar_ms = 0.0
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        trainer.flat.all_reduce_grads(trainer.group, trainer.n_buckets)
        barrier()
        a0.record()
        for _ in range(5):
            trainer.flat.all_reduce_grads(trainer.group, trainer.n_buckets)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 5
    prof = {}
    if rank == 0:
        ops.profile_enable(True)
    trainer.use_graph = False   # attribution pass runs eagerly (CUDA-event brackets around every launch)
    trainer.step(ir, vis)   # every rank steps (the step contains the gradient all-reduce)
    torch.cuda.synchronize()
    if rank == 0:
        prof = ops.profile_summary()
        ops.profile_enable(False)
    ops.set_direct_param_grads(False)
    t = torch.tensor([ms, ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0]), float(t[1])
    if rank != 0:
        return None
    pk = peaks()
    pairs = B * world * args.steps
    cpu = None
    if world == 1 and args.cpu_sample > 0:   # the reference path's training step on this box's host cores, bounded sample
        n = max(1, min(args.cpu_sample, B))
        r = cpu_train_pairs_per_s(n, S, reps=1, warmup=1)
        cpu = {"value": n / r["mean_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{n} of {B} pairs, {S}x{S}: fp32 oracle forward (train mode) + dense a008 loss restatement + autograd backward, "
                         f"torch CPU ops, 1 step after 1 warm-up"}
    roofline, kernels = roofline_block(prof, pk, 1)
    nbytes = trainer.flat.numel * 4
    value = pairs / (ms / 1e3)
    model_tf = value * 3 * gflop_per_pair(S) / 1e3
    return {"metric": "training image pairs/sec", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "steps_per_s": args.steps / (ms / 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision + " fwd / tensor-core bwd, fp32 accumulation and master weights",
            "data": "synthetic",
            "config": {"workload": f"training step B={B}/GPU {S}x{S} pairs, data parallel (BASELINE configs[2])",
                       "global_batch": B * world, "loss": "a008 loss + gradient in libswinfuse kernels (sf_fusion_loss: separable MS-SSIM+L1, Sobel, intensity; clamp folded in)",
                       "optimizer": "Adam lr 1e-2, sf_adam_step over one flat buffer",
                       "collective": "one NCCL all-reduce of the flat fp32 gradient buffer" if world > 1 else "none",
                       "launch": "eager" if args.no_graph else "cuda-graph replay of zero-grad + weight re-pack + forward + loss + backward; all-reduce + Adam eager"},
            "allreduce_ms": ar_ms, "allreduce_bytes": nbytes if world > 1 else 0,
            "allreduce_busbw_gbs": (2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9) if ar_ms > 0 else None,
            "gpu_launches": launches * args.steps, "loss_first": float(loss0), "loss_last": float(loss), "clocks": clocks,
            "model_tflops": model_tf, "model_tensor_frac": model_tf / pk["bf16_tflops_sustained"],
            "roofline": roofline, "kernels": dict(list(kernels.items())[:24]), "cpu_baseline": cpu}


def run_train(args, rank: int, world: int, local_rank: int):
    precision = pick_precision(args.precision)
    line = measure_training(args, rank, world, local_rank, precision)
    if rank == 0:
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
