#!/bin/bash
# One GPU-box pass that produces everything profiles/ needs for a round:
#   tools/gpu_round.sh <tag> [full]   (run through gpurun; outputs under gpurun_out/)
# 1. pytest -m gpu, smoke   2. bench (ours, reference arm)   3. ncu launch list of the same bench command
# 4. one `ncu --set full` capture of the dominant kernel (stage-0 window attention), exported as CSV.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
if [ "$2" = "full" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $out/${tag}_smoke.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || tail -5 $out/${tag}_bench.err
python tools/summarize_bench.py $out/${tag}_bench.json | cut -c1-150 | head -30
timeout 600 python bench.py --mode train --steps 6 --warmup 3 > $out/${tag}_train_1gpu.json 2> $out/${tag}_train_1gpu.err || tail -5 $out/${tag}_train_1gpu.err
python -c "import json,sys; d=json.load(open('$out/${tag}_train_1gpu.json')); print('train', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms/step', 'loss', d['loss_first'], '->', d['loss_last'])"
if [ "$2" = "full" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_reference_arm.json 2> $out/${tag}_reference_arm.err
  cat $out/${tag}_reference_arm.json | cut -c1-400
fi
# launch list of the same command (serialised, cold-cache: shares only)
# (capped at the first 3000 launches = the eager warm-up forwards plus the start of the graph replays: ncu serialises and
#  replays every launch, the whole default run would take tens of minutes of box time)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --cpu-sample 0 > $out/${tag}_ncu_bench.log 2>&1
python tools/ncu_launch_summary.py $out/${tag}_ncu_launches.csv > $out/${tag}_ncu_launch_summary.txt 2>&1; head -30 $out/${tag}_ncu_launch_summary.txt
# full capture of the dominant kernel on its own (stage-0 window attention, B=64)
K=${NCU_KERNEL:-k_attn_pack4}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o $out/${tag}_full_top \
    python tools/prof_ops.py wa 0 shift > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_full_top.ncu-rep --page raw --csv > $out/${tag}_full_top_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py $out/${tag}_full_top_raw.csv > $out/${tag}_full_top_summary.txt 2>&1; cat $out/${tag}_full_top_summary.txt
