"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("sf::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[14].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {tot / 1e3:.3f} ms (serialised, cold-cache; shares only)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / tot:6.3f}  n={v[0]:5d}  {v[1] / 1e3:9.3f} ms  avg {v[1] / v[0]:9.1f} us  {k[:90]}")
