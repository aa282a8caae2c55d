"""One CSR-input smoothing + column-sum launch at a given size (developer timing / ncu target):
    python tools/csr_one.py [N] [window]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import infercnvpy_b200 as cnv
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
var = cnv.datasets.synthetic_var(20000, seed=0)
Xd = cnv.datasets.device_counts(N, 20000, dev, seed=1000)
csr = Xd.to_sparse_csr()
t = (csr.crow_indices().to(torch.int64), csr.col_indices().to(torch.int32), csr.values())
del Xd, csr
with DevicePlan(build_layout(var, w, 10), dev) as plan:
    s, c = plan.colsum(t)
    plan.set_reference(plan.mean_from_sums(s, c))
    for _ in range(3):
        tmp = plan.smooth(t, 3.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tmp = plan.smooth(t, 3.0); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    by = 8.0 * t[2].numel() + 4.0 * N + 4.0 * plan.K * N
    print(f"window {w} rows {N} nnz {t[2].numel()}: smooth_csr {ms:.3f} ms = {by / ms / 1e6:.0f} GB/s algorithmic; launch {plan.launch_info()}")
    a.record(); s, c = plan.colsum(t); b.record(); torch.cuda.synchronize()
    print(f"colsum_csr {a.elapsed_time(b):.3f} ms")
