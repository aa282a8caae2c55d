"""Small end-to-end calls for compute-sanitizer (memcheck / racecheck), developer aid:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Covers: row-pair kernel, CSR input (delta kernel with its producer / consumer barrier protocol and shared-memory atomics;
staged-row kernel for reference categories), window 250, runtime-weight tier, direct kernel in parts, wide-row centring,
per-gene layer, fused filter / CSR compaction, ITH correlation, kNN (tcgen05) with fuzzy graph, Leiden sweeps, UMAP, t-SNE."""
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
# reference categories: bounded centring -> staged-row CSR kernel, per-category column sums
var = cnv.datasets.synthetic_var(2400, seed=2)
X = cnv.datasets.synthetic_counts(41, 2400, seed=5)
obs = pd.DataFrame({"ct": np.array(["a", "b", "c"])[np.arange(41) % 3]}, index=[str(i) for i in range(41)])
for container in (np.asarray, sp.csr_matrix):
    a = cnv.AnnData(container(X), obs=obs.copy(), var=var)
    _, res, _ = cnv.tl.infercnv(a, reference_key="ct", reference_cat=["a", "b"], inplace=False, chunksize=16)
    print("categories", container.__name__, res.nnz)
# pca -> neighbors (tcgen05 kNN) -> leiden -> umap -> tsne -> cnv_score on a small clone data set
var = cnv.datasets.synthetic_var(3000, seed=0)
X, clone = cnv.datasets.synthetic_counts_with_cnv(300, var, seed=7)
a = cnv.AnnData(X, obs=pd.DataFrame({"clone": [f"k{c}" for c in clone]}, index=[f"c{i}" for i in range(300)]), var=var)
cnv.tl.infercnv(a, reference_key="clone", reference_cat="k0", chunksize=100)
cnv.tl.pca(a, n_comps=10)
cnv.pp.neighbors(a)
cnv.tl.leiden(a)
cnv.tl.umap(a, maxiter=20)
cnv.tl.tsne(a, n_iter=20, perplexity=10)
cnv.tl.cnv_score(a)
print("workflow", a.obs["cnv_leiden"].nunique(), a.obsm["X_cnv_umap"].shape, a.obsm["X_cnv_tsne"].shape)
rng = np.random.default_rng(0)
a = cnv.AnnData(rng.normal(size=(150, 40)), obs=pd.DataFrame({"g": rng.integers(0, 3, 150).astype(str)}, index=[str(i) for i in range(150)]))
print(cnv.tl.ithgex(a, "g", inplace=False))
