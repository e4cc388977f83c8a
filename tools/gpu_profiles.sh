#!/bin/bash
# Evidence pass for profiles/ (run through gpurun; everything lands in gpurun_out/):
#   tools/gpu_profiles.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 600 python tools/bench_wa.py > $out/${tag}_wa_microbench.json 2> $out/${tag}_wa_microbench.err || tail -3 $out/${tag}_wa_microbench.err
# per-kernel counters of ONE forward at the bench shape (B=64) and ONE training step (B=32)
timeout 300 python tools/prof_forward.py > $out/${tag}_prof_forward_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $out/${tag}_ncu_forward_metrics.csv \
    python tools/prof_forward.py > $out/${tag}_ncu_forward.log 2>&1
timeout 300 python tools/prof_forward.py train > $out/${tag}_prof_train_plain.log 2>&1 &&
timeout 1500 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $out/${tag}_ncu_train_metrics.csv \
    python tools/prof_forward.py train > $out/${tag}_ncu_train.log 2>&1
# launch list of the bench command (serialised, cold-cache: shares only)
C="python bench.py --steps 2 --warmup 1 --cpu-sample 0 --train-steps 0 --no-highres"
timeout 300 $C > $out/${tag}_ncu_bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_ncu_launches.csv $C > $out/${tag}_ncu_bench.log 2>&1
ls -la $out | tail -12
