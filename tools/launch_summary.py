"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum) next to bench.py's CUDA-event stage table.
usage: python tools/launch_summary.py <launches.csv> <bench.json> <tag>  > profiles/<tag>_launch_summary.txt"""
import collections
import csv
import json
import sys

launches, bench_json, tag = sys.argv[1:4]
rows = [r for r in csv.reader(open(launches)) if len(r) > 8]
h = rows[0]
iK, iV = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[iK].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iV].replace(",", "")) / 1e6
d = json.loads(open(bench_json).read().strip().splitlines()[-1])
bench = collections.OrderedDict()
for k in d["kernels"]:
    bench["k_" + k["kernel"]] = bench.get("k_" + k["kernel"], 0.0) + k["ms"]
once = {k: v for k, v in agg.items() if k in ("k_densify_sa", "k_relayout")}
step = {k: v for k, v in agg.items() if k.startswith("k_") and k not in once}
other = {k: v for k, v in agg.items() if k not in once and k not in step}
tot = sum(v[1] for v in step.values()) + sum(v[1] for v in other.values())
btot = sum(bench.values())
print(f"# {tag} -- launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` under")
print(f"#   ncu --metrics gpu__time_duration.sum --clock-control none -c 600   (profiles/{tag}_launches.csv)")
print(f"# next to the per-kernel table of the un-profiled bench run (profiles/{tag}_bench.json: CUDA events on the engine's main stream).")
print("# ncu times are serialised and cold-cache: compare SHARES, not absolutes.  A bench interval also contains the host work that")
print("# precedes the launch.  Kernels launched on the auxiliary stream (wbody_seedB_heavy, wbody_chain_heavy) run BESIDE the next")
print("# main-stream kernel in the bench: their bench interval is only the part that neighbour did not cover (ncu serialises them).")
print(f"# bench: {d['ms_per_step']:.1f} ms/step device, value {d['value']:.0f} pairs/s, e2e {d['e2e']['value']:.0f} pairs/s, parity {d['parity']}")
print()
print("# index-open kernels (once per jnibwa_openIndex, not part of a step): " + "; ".join(f"{k} {v[1]:.1f} ms" for k, v in once.items()))
print()
print(f"{'kernel':<34}{'launches':>9}{'ncu ms':>10}{'ncu share':>10}{'bench ms/step':>14}{'bench share':>12}")
for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
    b = bench.get(k)
    print(f"{k:<34}{v[0]:>9}{v[1]:>10.2f}{100 * v[1] / tot:>9.1f}%" + (f"{b:>14.3f}{100 * b / btot:>11.1f}%" if b is not None else ""))
o, n = sum(v[1] for v in other.values()), sum(v[0] for v in other.values())
print(f"{'CUB scans / sorts / memsets':<34}{n:>9}{o:>10.2f}{100 * o / tot:>9.1f}%")
