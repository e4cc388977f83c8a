"""Per-kernel summary of an `ncu --metrics ... --csv` log (long format: one line per launch and metric): launches, mean
duration, DRAM bytes and throughput %, SM / tensor / XU pipe utilisation, issue slots, warps, registers.
    python tools/ncu_metrics_summary.py gpurun_out/<tag>_ncu_forward_metrics.csv > profiles/<tag>_ncu_forward_summary.txt"""
import collections
import csv
import re
import sys

SHORT = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd_MB", "dram__bytes_write.sum": "wr_MB",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu%",
         "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
         "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block"}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iu, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    ids = collections.defaultdict(set)
    for r in rd:
        if len(r) <= iv or r[im] not in SHORT:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("sf::", "").replace("void ", "")
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= UNIT.get(r[iu], 1.0)
        per[name][SHORT[r[im]]].append(v)
        ids[name].add(r[iid])
    cols = ["us", "rd_MB", "wr_MB", "dram%", "sm%", "tensor%", "xu%", "issue%", "warps%", "regs", "grid", "block"]
    tot = sum(sum(m["us"]) for m in per.values())
    print(f"{'kernel':64s} {'n':>5s} {'sum_ms':>8s} {'share':>6s} " + " ".join(f"{c:>8s}" for c in cols) + "   (means per launch; cold-cache, serialised)")
    for name, m in sorted(per.items(), key=lambda kv: -sum(kv[1]["us"])):
        n = len(ids[name])
        s = sum(m["us"])
        mean = lambda k: (sum(m[k]) / len(m[k])) if m[k] else float("nan")
        print(f"{name[:64]:64s} {n:5d} {s / 1e3:8.3f} {s / tot:6.3f} " + " ".join(f"{mean(c):8.1f}" for c in cols))


if __name__ == "__main__":
    main()
