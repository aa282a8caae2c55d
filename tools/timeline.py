"""Per-row stage durations of the smoothing kernel from clock64 stamps (developer aid)."""
import sys, json, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import infercnvpy_b200 as cnv
from infercnvpy_b200 import _lib
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
window = int(sys.argv[2]) if len(sys.argv) > 2 else 100
G = 20000; dev = torch.device("cuda", 0)
var = cnv.datasets.synthetic_var(G, seed=0)
Xd = cnv.datasets.device_counts(N, G, dev, seed=1000)
layout = build_layout(var, window, 10)
lib = _lib.load()
with DevicePlan(layout, dev) as plan:
    info = plan.launch_info(); grid = info["ctas_per_sm"] * info["n_sm"]; R = 48
    s, c = plan.colsum(Xd); plan.set_reference(plan.mean_from_sums(s, c))
    plan.smooth(Xd, 3.0); torch.cuda.synchronize()
    buf = torch.zeros((grid, R, 16), dtype=torch.int64, device=dev)
    lib.icnv_debug_set_timeline(buf.data_ptr(), R)
    plan.smooth(Xd, 3.0); torch.cuda.synchronize()
    lib.icnv_debug_set_timeline(None, 0)
t = buf.cpu().numpy().astype(np.float64)
ROWS = info.get('rows', 1)
rows = slice(4, min(R, N // (grid * ROWS) - 2))   # steady state (iterations; an iteration stages ROWS rows)
T = t[:, rows, :]
def d(a, b): x = (T[..., b] - T[..., a]).ravel(); x = x[(T[..., a].ravel() > 0) & (T[..., b].ravel() > 0)]; return float(np.median(x)), float(np.mean(x))
names = {(0,1): "wait row (mbarrier)", (1,2): "own phase-2 blocks", (2,3): "wait BAR_A", (3,13): "after BAR_A", (13,15): "phase 3 reads+FMA (last row)", (15,4): "hand-over + stores (last row)", (13,4): "phase 3 + stores (all staged rows)",
         (4,7): "BAR_B", (0,7): "whole iteration"}
period = np.diff(t[:, 4:rows.stop, 0], axis=1).ravel()
print(json.dumps(dict(N=N, window=window, grid=grid, rows_per_iteration=ROWS, period_median=float(np.median(period)), period_mean=float(period.mean()))))
for k, v in names.items():
    m, a = d(*k); print(f"{v:32s} median {m:9.0f}  mean {a:9.0f} cycles")
