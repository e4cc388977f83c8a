import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "dtype", d["dtype"], d["clocks"])
print("roofline", d["roofline"])
for k, v in d["kernels"].items():
    print(f"{k:28s} n={v['launches_per_step']:4d} ms={v['ms_per_step']:8.3f} share={v['share']:.3f} tflops={v['tflops']:8.1f} gbs={v['gbs']:8.1f}")
print("cpu", d.get("cpu_baseline"))
