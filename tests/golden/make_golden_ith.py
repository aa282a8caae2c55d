"""Golden values of the ITH scores from the UNMODIFIED reference (``tl/_scores.py:77-221``), build container only:

    python tests/golden/make_golden_ith.py      # rewrites tests/golden/ith_scores.npz

Inputs are regenerated from seeds by ``ith_case()`` (shared with the tests)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pandas as pd
import scipy.sparse as sp

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def ith_case():
    """-> (expression [n, g] float32 dense, X_cnv [n, k] float64 CSR with ~85 % zeros, labels [n])."""
    rng = np.random.default_rng(77)
    n, g, k = 260, 300, 120
    labels = rng.choice(np.array(["p1", "p2", "p3", "solo"]), size=n, p=[0.5, 0.3, 0.195, 0.005])
    labels[0] = "solo"  # exactly one cell in this group: skipped by the reference (:135)
    labels[1:][labels[1:] == "solo"] = "p1"
    base = {lab: rng.normal(size=g) for lab in np.unique(labels)}
    expr = np.stack([base[lab] * rng.uniform(0.2, 1.0) + rng.normal(size=g) for lab in labels]).astype(np.float32)
    expr = np.log1p(np.abs(expr)).astype(np.float32)
    dense = rng.normal(scale=0.1, size=(n, k))
    dense[rng.random((n, k)) < 0.85] = 0.0
    dense[5] = dense[6]  # two identical cells
    return expr, sp.csr_matrix(dense), labels


def main():
    from oracle import ref_loader

    if not ref_loader.available():
        raise SystemExit("reference not present; goldens can only be regenerated in the build container")
    _, ref_sco = ref_loader.load()
    expr, x_cnv, labels = ith_case()
    obs = pd.DataFrame({"patient": labels}, index=[f"c{i}" for i in range(len(labels))])
    adata = ref_loader.MiniAnnData(expr, obs=obs, obsm={"X_cnv": x_cnv})
    gex = ref_sco.ithgex(adata, "patient", inplace=False)
    cna = ref_sco.ithcna(adata, "patient", inplace=False)
    keys = sorted(gex)
    assert sorted(cna) == keys
    np.savez_compressed(
        HERE / "ith_scores.npz", keys=np.array(keys), ithgex=np.array([gex[k] for k in keys]), ithcna=np.array([cna[k] for k in keys])
    )
    print({k: (float(gex[k]), float(cna[k])) for k in keys})


if __name__ == "__main__":
    main()
