"""Prints the interesting part of a bench.py JSON line (development aid)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
cb = d.get("cpu_baseline") or {}
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "cpu", round(cb.get("value", 0)), "parity", d.get("parity"), "launches", d.get("gpu_launches"))
print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("note", "peak_source")})
for k in d["kernels"]:
    if k["ms"] > 0.5:
        print("  ", k["kernel"], k["ms"])
