#!/bin/bash
# Multi-GPU lines of a round: tools/gpu_scale.sh <N> <tag>   (run through `gpurun --gpus N`; outputs under gpurun_out/)
# One process per GPU under torch.distributed.run; inference (BASELINE configs[1]) and the training step (configs[2]).
N=${1:-2}
tag=${2:-r1}
out=gpurun_out
mkdir -p $out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 10 --warmup 3 --cpu-sample 0 > $out/${tag}_infer_${N}gpu.json 2> $out/${tag}_infer_${N}gpu.err || tail -5 $out/${tag}_infer_${N}gpu.err
run 29512 --mode train --steps 6 --warmup 3 > $out/${tag}_train_${N}gpu.json 2> $out/${tag}_train_${N}gpu.err || tail -5 $out/${tag}_train_${N}gpu.err
python - <<PY
import json
for k in ("infer", "train"):
    try:
        d = json.load(open("$out/${tag}_%s_${N}gpu.json" % k))
        print(k, "$N gpus", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms/step", d.get("clocks"))
    except Exception as e:
        print(k, "failed", e)
PY
