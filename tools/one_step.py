"""One launch of every kernel of the tl.infercnv step at the bench size, for `ncu --set full` (developer aid):
    ncu --set full --clock-control none --import-source on -k regex:icnv -o gpurun_out/step python tools/one_step.py [N] [window]
Also exercises the per-gene layer kernel on a slice (calculate_gene_values=True)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import infercnvpy_b200 as cnv
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
window = int(sys.argv[2]) if len(sys.argv) > 2 else 100
G = 20000
dev = torch.device("cuda", 0)
var = cnv.datasets.synthetic_var(G, seed=0)
Xd = cnv.datasets.device_counts(N, G, dev, seed=1000)
with DevicePlan(build_layout(var, window, 10), dev) as plan:
    sums, counts = plan.colsum(Xd)
    plan.set_reference(plan.mean_from_sums(sums, counts))
    tmp = plan.smooth(Xd, 3.0)
    out, stats = plan.center(tmp)
    thr, row_abs, row_nnz = plan.threshold(out, stats, 5000, 1.5)
    indptr, indices, data = plan.to_csr(out, row_nnz)
    pre, stats = plan.center(tmp)
    plan.filter_to_csr(pre, stats, 5000, 1.5)  # the fused path tl.infercnv takes (count + scan + filtering compaction)
    gv = plan.gene_values(tmp[:10000], 5000, thr)
    torch.cuda.synchronize()
    print("ok", plan.launch_info(), int(indptr[-1]), float(torch.nan_to_num(gv).abs().sum()))
