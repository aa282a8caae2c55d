"""Key metrics of every kernel in an .ncu-rep (developer aid): python tools/ncu_metrics.py file.ncu-rep"""
import csv, subprocess, sys, io

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "sm__inst_executed.sum.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if name in seen and "--all" not in sys.argv:
        continue
    seen.add(name)
    print("=====", name[:110])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:78s} {r[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.1:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
