"""Small end-to-end calls for compute-sanitizer (memcheck / racecheck), developer aid:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Covers: row-pair kernel, single-row kernel (CSR input), window 250, runtime-weight tier, direct kernel in parts,
wide-row centring, per-gene layer, ITH correlation."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, pandas as pd, scipy.sparse as sp
import infercnvpy_b200 as cnv

def run(g, n, **kw):
    var = cnv.datasets.synthetic_var(g, seed=2, with_extras=True)
    X = cnv.datasets.synthetic_counts(n, g, seed=g + n)
    for container in (np.asarray, sp.csr_matrix):
        a = cnv.AnnData(container(X), var=var)
        chr_pos, res, pg = cnv.tl.infercnv(a, inplace=False, chunksize=16, **kw)
        print(g, n, kw, container.__name__, res.shape, res.nnz, None if pg is None else float(np.nansum(np.abs(pg))))

run(2400, 37)                                             # tier 0, row pairs (odd row count) / single row for CSR
run(2400, 20, window_size=250)                            # C partials
run(2400, 20, window_size=50, calculate_gene_values=True) # tier 1 + per-gene layer
run(3000, 9, window_size=37, step=7)                      # direct kernel
run(6000, 5, window_size=20, step=1)                      # wide rows
rng = np.random.default_rng(0)
a = cnv.AnnData(rng.normal(size=(150, 40)), obs=pd.DataFrame({"g": rng.integers(0, 3, 150).astype(str)}, index=[str(i) for i in range(150)]))
print(cnv.tl.ithgex(a, "g", inplace=False))
