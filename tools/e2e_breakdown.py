"""Where the end-to-end time of cnv.tl.infercnv goes (developer aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import infercnvpy_b200 as cnv
N, G = 100000, 20000
dev = torch.device("cuda", 0)
host = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
Xd = cnv.datasets.device_counts(N, G, dev, seed=1000)
host.copy_(Xd); torch.cuda.synchronize()
arr = host.numpy()
print("from_numpy pinned:", torch.from_numpy(arr).is_pinned())
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
dst = torch.empty_like(Xd)
print("H2D 8GB pinned one shot  s:", t(lambda: dst.copy_(host, non_blocking=True)))
print("H2D 8GB pinned 256MB slabs s:", t(lambda: [dst[i:i+3355].copy_(host[i:i+3355], non_blocking=True) for i in range(0, N, 3355)]))
var = cnv.datasets.synthetic_var(G, seed=0)
adata = cnv.AnnData(arr, var=var)
import cProfile, pstats
cnv.tl.infercnv(adata, inplace=False)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); chr_pos, res, _ = cnv.tl.infercnv(adata, inplace=False); torch.cuda.synchronize(); t1 = time.perf_counter()
pr.disable()
print("infercnv total s:", t1 - t0, "nnz", res.nnz)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
