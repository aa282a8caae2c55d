"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Needs ``/root/reference`` (absent on the GPU box, where the committed ``.npz``
files are used instead).  Inputs are re-generated from seeds by
``tests/golden/cases.py``; each fixture stores a checksum of the input so a
drifting RNG stream is caught instead of silently comparing different data.

What is stored per case: the reference profile that the reference computed
(``_get_reference``), ``chr_pos`` and the CSR float64 result of the public
``infercnv()`` (``/root/reference/src/infercnvpy/tl/_infercnv.py:18-161``), and
for some cases ``cnv_score`` (``tl/_scores.py:14-74``) or the per-gene layer of
``calculate_gene_values=True``.  ``python tests/golden/make_golden.py NAME ...`` regenerates only those cases.
"""

from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader  # noqa: E402
from tests.golden.cases import CASES, build_case  # noqa: E402


def checksum(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(case):
    ref_inf, ref_sco = ref_loader.load()
    X, var, obs, kw = build_case(case)
    Xin = sp.csr_matrix(X) if case.get("container") == "csr" else X
    adata = ref_loader.MiniAnnData(Xin, obs=obs, var=var)
    profile = ref_inf._get_reference(
        adata, kw.get("reference_key"), kw.get("reference_cat"), kw.get("reference"), None
    )
    profile = np.asarray(profile)
    chr_pos, res, per_gene = ref_inf.infercnv(
        adata, inplace=False, n_jobs=1, calculate_gene_values=bool(case.get("gene_values")), **kw
    )
    res = sp.csr_matrix(res)
    out = {
        "x_sha256": np.array(checksum(X)),
        "profile": profile,
        "chr_names": np.array(list(chr_pos.keys())),
        "chr_offsets": np.array([int(v) for v in chr_pos.values()], dtype=np.int64),
        "data": res.data,
        "indices": res.indices.astype(np.int32),
        "indptr": res.indptr.astype(np.int64),
        "shape": np.array(res.shape, dtype=np.int64),
    }
    if case.get("gene_values"):
        out["per_gene"] = np.asarray(per_gene, dtype=np.float64)  # [n, n_var], NaN = not covered / masked
    if case.get("score_labels") is not None:
        rng = np.random.default_rng(case["score_labels"])
        labels = rng.integers(0, 4, size=X.shape[0]).astype(str)
        adata.obsm["X_cnv"] = res
        adata.obs = adata.obs.assign(grp=labels)
        score = ref_sco.cnv_score(adata, "grp", inplace=False)
        out["score_labels"] = labels
        out["score_keys"] = np.array(list(score.keys()))
        out["score_vals"] = np.array([float(v) for v in score.values()])
    return out


def main():
    if not ref_loader.available():
        raise SystemExit("reference not present; goldens can only be regenerated in the build container")
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case["name"] not in only:
            continue
        out = run_reference(case)
        path = HERE / f"{case['name']}.npz"
        np.savez_compressed(path, **out)
        print(f"{case['name']:>22}: shape={tuple(out['shape'])} nnz={out['data'].size} -> {path.name} ({path.stat().st_size/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
