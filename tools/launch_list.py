"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch log into a markdown table (developer aid):
    python tools/launch_list.py gpurun_out/launches.csv > profiles/launches_rN.md"""
import csv, sys, collections

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    k = r["Kernel Name"]
    a = tot.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
allus = sum(v[1] for v in tot.values())
ours = sum(v[1] for k, v in tot.items() if "icnv::" in k or k.startswith("icnv") or "icnv" in k)
print("| kernel | launches | total us | share of all | share of icnv kernels |")
print("|---|---|---|---|---|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    mine = "icnv" in k
    print(f"| `{k[:90]}` | {n} | {us:.1f} | {100 * us / allus:.1f}% | {100 * us / ours:.1f}% |" if mine else f"| `{k[:90]}` (torch: synthetic input / plumbing, untimed) | {n} | {us:.1f} | {100 * us / allus:.1f}% | |")
