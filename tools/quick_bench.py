"""Developer timing of the individual kernels (not the driver bench). Usage: python tools/quick_bench.py [N]"""
import sys, json, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
import numpy as np, torch
import infercnvpy_b200 as cnv
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
G = 20000
dev = torch.device("cuda", 0)
var = cnv.datasets.synthetic_var(G, seed=0)
Xd = cnv.datasets.device_counts(N, G, dev, seed=1000)
peak = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"] if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0

WINDOWS = [int(w) for w in os.environ.get("QB_WINDOWS", "100,250").split(",")]
REPS = int(os.environ.get("QB_REPS", "5"))

def timeit(fn, reps=None, warm=2):
    reps = reps or REPS
    warm = min(warm, reps)
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if os.environ.get("QB_TRACE"):
        print("trace", [round(t, 3) for t in ts])
    return min(ts), float(np.median(ts))

# box sanity: plain device copy bandwidth (read + write), to spot a slow box
_a = torch.empty(1 << 28, dtype=torch.float32, device=dev); _b = torch.empty_like(_a)
_t = timeit(lambda: _b.copy_(_a), reps=5)
print(json.dumps(dict(box_copy_GBs=2 * _a.numel() * 4 / _t[0] / 1e6)))
del _a, _b
for window in WINDOWS:
    layout = build_layout(var, window, 10)
    with DevicePlan(layout, dev) as plan:
        info = plan.launch_info()
        K = plan.K
        t_cs = timeit(lambda: plan.colsum(Xd))
        sums, counts = plan.colsum(Xd)
        ref = plan.mean_from_sums(sums, counts)
        plan.set_reference(ref)
        tmp = torch.empty((N, plan.tmp_width()), dtype=torch.float64, device=dev)
        out = torch.empty((N, K), dtype=torch.float32, device=dev)
        stats = torch.empty((N, 2), dtype=torch.float64, device=dev)
        t_sm = timeit(lambda: plan.smooth(Xd, 3.0, tmp=tmp))
        t_ce = timeit(lambda: plan.center(tmp, out=out, row_stats=stats))
        t_th = timeit(lambda: plan.threshold(out, stats, 5000, 1.5))
        if os.environ.get("QB_GENEVALS"):
            thr, _, _ = plan.threshold(out, stats, 5000, 1.5)
            ng = min(N, 10000)
            gv = torch.empty((ng, G), dtype=torch.float64, device=dev)
            t_gv = timeit(lambda: plan.gene_values(tmp[:ng], 5000, thr, out=gv))
            print(json.dumps(dict(window=window, gene_values_rows=ng, gene_values_ms=t_gv, gene_values_write_GBs=ng * G * 8 / t_gv[0] / 1e6)))
            del gv
        by = N * (4 * G + 4 * K)
        print(json.dumps(dict(window=window, N=N, K=K, launch=info,
              colsum_ms=t_cs, colsum_GBs=N*G*4/t_cs[0]/1e6,
              smooth_ms=t_sm, smooth_GBs=by/t_sm[0]/1e6, smooth_frac=by/t_sm[0]/1e6/peak, cells_per_s=N/t_sm[0]*1e3,
              center_ms=t_ce, center_GBs=N*(plan.tmp_width()*8+K*4)/t_ce[0]/1e6, thr_ms=t_th, thr_GBs=2*N*K*4/t_th[0]/1e6)))
