"""Markdown table from an `ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv` launch list:
    python tools/launch_list.py gpurun_out/xx/launches.csv "title" > profiles/launches_rN.md"""
import csv, sys
from collections import OrderedDict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
lines = [ln for ln in open(path, errors="replace") if ln.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = lambda n: any(s in n for s in ("icnv::", "smooth_", "colsum_", "center_rows", "threshold_kernel", "dense_to_csr", "nnz_to_indptr",
                                      "build_bounds", "mean_from_sums", "reduce_partials", "count_rows", "col_table", "zrow_kernel",
                                      "probe_smem", "gene_values", "knn_", "fuzzy_", "community_", "gram_", "project_", "filter_csr"))
tot = sum(a[1] for a in agg.values())
tot_ours = sum(a[1] for n, a in agg.items() if ours(n))
print(f"# ncu launch list: {title}\n")
print(f"`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised; only the SHARES are comparable with the CUDA-event numbers of bench.py). {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms in total, {tot_ours / 1e3:.2f} ms in libicnv.so kernels.\n")
print("| kernel | launches | total us | share of all | share of libicnv kernels |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tag = "" if ours(n) else " (torch: synthetic input / plumbing)"
    share = f"{100 * a[1] / tot_ours:.1f}%" if ours(n) else ""
    print(f"| `{n[:100]}`{tag} | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {share} |")
