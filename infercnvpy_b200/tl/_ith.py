"""``cnv.tl.ithcna`` / ``cnv.tl.ithgex`` — intratumoral-heterogeneity scores
(reference: ``/root/reference/src/infercnvpy/tl/_scores.py:77-221``).

Per group of cells: Pearson correlation of every pair of cells (``np.corrcoef`` over rows, ``:136,209``), score =
inter-quartile range of ALL entries of that cells x cells matrix (``np.percentile(pcorr, [75, 25])``, ``:141,214``).
The correlation matrix is built by ``icnv_row_corrcoef_f64`` (fp64 tiles on the GPU); the two quartiles are read off a
device sort of its entries with numpy's linear-interpolation rule.  Groups do not shard: every rank scores the cells
it holds ("replicas only").
"""

from __future__ import annotations

import math
from collections.abc import Mapping

import numpy as np
import scipy.sparse as sp

from .. import _lib


def _np_linear_quantile(sorted_flat, n_total: int, q: float) -> float:
    """numpy's default ('linear') percentile on an ascending device vector: virtual index ``q * (N - 1)`` computed like
    ``numpy.lib._function_base_impl._compute_virtual_index`` (alpha = beta = 1), value by numpy's ``_lerp``."""
    vi = n_total * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0
    lo = int(math.floor(vi))
    lo = min(max(lo, 0), n_total - 1)
    hi = min(lo + 1, n_total - 1)
    g = vi - math.floor(vi)
    a = float(sorted_flat[lo].item())
    b = float(sorted_flat[hi].item())
    diff = b - a
    val = a + diff * g
    if g >= 0.5:
        val = b - diff * (1.0 - g)
    return val


def _group_iqr(block: np.ndarray, device) -> float:
    """IQR of the cell-cell correlation matrix of one group (dense float64 block [n_g, K])."""
    import torch

    lib = _lib.load()
    n, K = block.shape
    free, _ = torch.cuda.mem_get_info(device)
    need = 3 * 8 * n * n + 8 * n * K
    if need > free:
        raise _lib.IcnvError(
            f"group of {n} cells needs {need / 2**30:.1f} GiB for its correlation matrix and the sort; "
            f"{free / 2**30:.1f} GiB are free on {device}"
        )
    X = torch.from_numpy(np.ascontiguousarray(block, dtype=np.float64)).to(device)
    corr = torch.empty((n, n), dtype=torch.float64, device=device)
    work = torch.empty((2 * n,), dtype=torch.float64, device=device)
    _lib.check(
        lib.icnv_row_corrcoef_f64(_lib.ptr(X), n, X.stride(0), K, _lib.ptr(corr), corr.stride(0), _lib.ptr(work), _lib.stream_handle(device)),
        "icnv_row_corrcoef_f64",
    )
    if bool(torch.isnan(corr).any().item()):  # np.percentile of an array with a NaN is NaN
        return float("nan")
    flat = torch.sort(corr.reshape(-1)).values
    del corr
    q75 = _np_linear_quantile(flat, n * n, 0.75)
    q25 = _np_linear_quantile(flat, n * n, 0.25)
    return q75 - q25


def _ith(adata, groupby: str, get_block) -> dict:
    import torch

    if not torch.cuda.is_available():
        raise _lib.IcnvError("infercnvpy_b200 ITH scores need a CUDA device; there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device())
    labels = adata.obs[groupby]
    res = {}
    for group in labels.unique():
        rows = np.flatnonzero(np.asarray(labels == group))
        X = get_block(rows)
        if sp.issparse(X):
            X = X.toarray()
        X = np.asarray(X)
        if X.shape[0] <= 1:  # _scores.py:135,208
            continue
        res[group] = _group_iqr(X, device)
    return res


def _store(adata, groupby, scores, key_added):
    obs_vals = np.empty(adata.shape[0])
    for group in adata.obs[groupby].unique():
        obs_vals[np.asarray(adata.obs[groupby] == group)] = scores[group]  # KeyError for 1-cell groups, like :144-146
    adata.obs[key_added] = obs_vals


def ithgex(
    adata,
    groupby: str,
    *,
    use_raw: bool | None = None,
    layer: str | None = None,
    inplace: bool = True,
    key_added: str = "ithgex",
) -> Mapping[str, float] | None:
    """ITHGEX diversity score from gene expression (``_scores.py:77-149``).  Same parameters and return value."""
    is_layer = layer is not None
    if use_raw and is_layer:  # _util.py:15-18
        raise ValueError(
            f"Cannot use expression from both layer and raw. You provided:'use_raw={use_raw}' and 'layer={layer}'"
        )
    mtx = adata.layers[layer] if is_layer else (adata.raw.X if use_raw else adata.X)
    if sp.issparse(mtx):
        mtx = mtx.tocsr()
    scores = _ith(adata, groupby, lambda rows: mtx[rows])
    if inplace:
        _store(adata, groupby, scores, key_added)
    else:
        return scores


def ithcna(
    adata,
    groupby: str,
    *,
    use_rep: str = "X_cnv",
    key_added: str = "ithcna",
    inplace: bool = True,
) -> Mapping[str, float] | None:
    """ITHCNA diversity score from the CNV matrix ``adata.obsm[use_rep]`` (``_scores.py:152-221``)."""
    mtx = adata.obsm[use_rep]
    if sp.issparse(mtx):
        mtx = mtx.tocsr()
    scores = _ith(adata, groupby, lambda rows: mtx[rows])
    if inplace:
        _store(adata, groupby, scores, key_added)
    else:
        return scores
