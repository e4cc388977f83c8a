"""Pick the judged metrics out of `ncu --page raw --csv` (one row per captured launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.avg.per_cycle_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    d = dict(zip(h, r))
    for k in KEYS:
        if k in d:
            print(f"{k:70s} {d[k]:>20s} {u[h.index(k)]}")
    stalls = sorted(((float(d[k]), k) for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
    print("stalls per issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in stalls[:12]))
    print()
