"""Summarise `ncu --page source --csv` output: hottest SASS lines by stall samples / executed."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
ia = h.index("Source"); isamp = h.index("Warp Stall Sampling (All Samples)"); iex = h.index("Instructions Executed")
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(I(r[isamp]) for r in data); totex = sum(I(r[iex]) for r in data)
print("total samples", tot, "total warp-inst", totex, "sass lines", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -I(r[isamp]))[:n]:
    print(f"{I(r[isamp]):7d} {100*I(r[isamp])/max(tot,1):5.1f}%  ex={I(r[iex]):9d}  {r[ia].strip()[:120]}")
print("--- top by executed")
for r in sorted(data, key=lambda r: -I(r[iex]))[:n//2]:
    print(f"ex={I(r[iex]):9d} {100*I(r[iex])/max(totex,1):5.1f}% samp={I(r[isamp]):6d} {r[ia].strip()[:110]}")
