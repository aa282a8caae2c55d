"""Synthetic inputs for tests and the bench.

The reference ships one bundled dataset (``oligodendroglioma.h5ad``,
``/root/reference/src/infercnvpy/datasets/__init__.py:13-19``) and downloads
another; neither is readable in this image (no h5py, no network), so every
test and bench line runs on the seeded synthetic spec below.  The spec is
fixed so numbers are comparable across sessions:

* ``G`` genes named ``g0..g{G-1}`` that are **not** position-sorted in memory
  (like maynard2020_3k), drawn over chr1..chr22 with human-like proportions;
* values ``log1p(Poisson(0.3))`` as float32 (density ~26 %).
"""

from __future__ import annotations

import numpy as np
import pandas as pd

CHROM_NAMES = [f"chr{i}" for i in range(1, 23)]
# relative gene counts per autosome
CHROM_PROPS = np.array(
    [2050, 1300, 1080, 750, 880, 1050, 920, 680, 780, 730, 1300, 1030, 320, 650, 600, 860, 1180, 270, 1470, 540, 230, 440],
    dtype=np.float64,
)


def synthetic_var(n_genes: int = 20000, seed: int = 0, *, with_extras: bool = False) -> pd.DataFrame:
    """``var`` frame with ``chromosome``, ``start``, ``end`` columns.

    ``with_extras=True`` (tests only) re-labels a few genes as chrX / chrY /
    chrM / a non-``chr`` contig / NaN so the masking rules of
    ``_infercnv.py:104-108`` and ``:327`` are exercised.
    """
    rng = np.random.default_rng(seed)
    chrom = rng.choice(CHROM_NAMES, size=n_genes, p=CHROM_PROPS / CHROM_PROPS.sum()).astype(object)
    start = rng.integers(0, 2 * 10**8, size=n_genes)
    if with_extras:
        pick = rng.permutation(n_genes)
        k = max(1, n_genes // 40)
        chrom[pick[0 * k : 1 * k]] = "chrX"
        chrom[pick[1 * k : 2 * k]] = "chrY"
        chrom[pick[2 * k : 3 * k]] = "chrM"
        chrom[pick[3 * k : 4 * k]] = "GL000219.1"
        chrom[pick[4 * k : 5 * k]] = np.nan
    var = pd.DataFrame(
        {"chromosome": chrom, "start": start, "end": start + 1000},
        index=pd.Index([f"g{i}" for i in range(n_genes)], name=None),
    )
    return var


def synthetic_counts(n_cells: int, n_genes: int, seed: int = 1000, lam: float = 0.3) -> np.ndarray:
    """Dense float32 ``log1p(Poisson(lam))`` matrix (host, numpy)."""
    rng = np.random.default_rng(seed)
    return np.log1p(rng.poisson(lam, size=(n_cells, n_genes))).astype(np.float32)


def synthetic_counts_with_cnv(n_cells: int, var: pd.DataFrame, seed: int = 7, lam: float = 0.6) -> tuple[np.ndarray, np.ndarray]:
    """Counts with planted chromosome-arm gains/losses in three "clones".

    Returns ``(X float32, clone_id int)``.  Clone 0 is "normal".  Used by the
    workflow tests (pca -> neighbors -> leiden -> cnv_score) where i.i.d. noise
    would have no structure to find.
    """
    rng = np.random.default_rng(seed)
    n_genes = var.shape[0]
    clone = rng.integers(0, 3, size=n_cells)
    rate = np.full((3, n_genes), lam)
    chrom = var["chromosome"].to_numpy()
    rate[1, chrom == "chr1"] *= 1.8
    rate[1, chrom == "chr7"] *= 0.45
    rate[2, chrom == "chr3"] *= 0.4
    rate[2, chrom == "chr11"] *= 2.0
    rate[2, chrom == "chr17"] *= 1.7
    X = np.log1p(rng.poisson(rate[clone])).astype(np.float32)
    return X, clone


def device_counts(n_cells: int, n_genes: int, device, seed: int = 1000, lam: float = 0.3):
    """``log1p(Poisson(lam))`` float32 generated directly in HBM (bench only).

    torch is used purely as an allocator / RNG here; the values never touch
    the host.
    """
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((n_cells, n_genes), dtype=torch.float32, device=device)
    # generate in slabs so the temporary rate tensor stays small
    slab = max(1, (1 << 28) // max(1, n_genes))
    for r0 in range(0, n_cells, slab):
        r1 = min(n_cells, r0 + slab)
        rates = torch.full((r1 - r0, n_genes), lam, dtype=torch.float32, device=device)
        out[r0:r1] = torch.log1p(torch.poisson(rates, generator=gen))
    return out
