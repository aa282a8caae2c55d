"""End-to-end tail behind the PCIe upload (developer aid): python tools/e2e_tail.py [N]
Times cnv.tl.infercnv on N pinned cells for several host copy thread counts and splits the call into upload wait / D2H + assembly."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import infercnvpy_b200 as cnv
from infercnvpy_b200.tl import _infercnv as mod

N, G = int(sys.argv[1]) if len(sys.argv) > 1 else 250000, 20000
dev = torch.device("cuda", 0)
host = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
for a in range(0, N, 50000):
    host[a:a + 50000].copy_(cnv.datasets.device_counts(min(50000, N - a), G, dev, seed=1000 + a))
torch.cuda.synchronize()
adata = cnv.AnnData(host.numpy(), var=cnv.datasets.synthetic_var(G, seed=0))
orig = mod._host_csr
acc = {}
def timed_host_csr(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    acc["host_csr"] = time.perf_counter() - t0
    return r
mod._host_csr = timed_host_csr
for thr in (16, 8, 4, 2):
    os.environ["ICNV_COPY_THREADS"] = str(thr)
    cnv.tl.infercnv(adata, inplace=False)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cnv.tl.infercnv(adata, inplace=False)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"threads {thr:2d}: infercnv {min(ts):.4f} s (median {sorted(ts)[1]:.4f}), of which D2H + host assembly {acc['host_csr']:.4f} s")
