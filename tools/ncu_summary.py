"""Markdown summary of an `ncu --set full` report (first launch of every kernel), plus the DRAM traffic of the
smoothing kernel for bench.py's roofline.traffic:
    python tools/ncu_summary.py gpurun_out/step_r1.ncu-rep "title" > profiles/ncu_summary_rN.md"""
import csv, io, json, subprocess, sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full --clock-control none summaries: {title}\n")
print(f"Source report: `{Path(rep).name}` (kept in gpurun_out/, scratch). One launch per kernel; times are cold-cache and serialised.\n")
seen = set()
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if name in seen:
        continue
    seen.add(name)
    print(f"### `{name[:120]}`\n\n| metric | unit | value |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {units[i]} | {r[i]} |")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.2:
                st.append((v, h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    print("| warps stalled per issue (>= 0.2) | inst | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)) + " |\n")
    if "smooth_kernel" in name:
        def val(k):
            i = hdr.index(k)
            return float(r[i].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        traffic["smooth"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
if len(sys.argv) > 3 and "smooth" in traffic:
    p = Path(sys.argv[3])
    d = json.loads(p.read_text()) if p.exists() else {}
    d[sys.argv[4]] = traffic["smooth"]
    p.write_text(json.dumps(d))
