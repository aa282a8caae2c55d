"""``cnv.tl.ithcna`` / ``cnv.tl.ithgex`` — intratumoral-heterogeneity scores
(reference: ``/root/reference/src/infercnvpy/tl/_scores.py:77-221``).

Per group of cells: Pearson correlation of every pair of cells (``np.corrcoef`` over rows, ``:136,209``), score =
inter-quartile range of ALL entries of that cells x cells matrix (``np.percentile(pcorr, [75, 25])``, ``:141,214``).
The correlation matrix is built by ``icnv_row_corrcoef_f64`` (fp64 tiles on the GPU); the two quartiles are exact
order statistics (numpy's linear-interpolation rule) found by a sample-bracketed selection instead of a sort of the
``n_g^2`` entries, so a group is limited by its correlation matrix alone (~130k cells on a 180 GB GPU).  Groups do not shard: every rank scores the cells
it holds ("replicas only").
"""

from __future__ import annotations

import math
from collections.abc import Mapping

import numpy as np
import scipy.sparse as sp

from .. import _lib


def _lerp_np(a: float, b: float, g: float) -> float:
    """numpy's ``_lerp`` (``numpy/lib/_function_base_impl.py``): ``a + (b - a) * g``, evaluated from ``b`` when ``g >= 0.5``."""
    diff = b - a
    val = a + diff * g
    if g >= 0.5:
        val = b - diff * (1.0 - g)
    return val


def _virtual_index(n_total: int, q: float):
    """numpy's default ('linear') percentile: virtual index ``q * (N - 1)`` computed like ``_compute_virtual_index``
    (alpha = beta = 1) -> (lower order statistic, upper order statistic, interpolation weight)."""
    vi = n_total * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0
    lo = min(max(int(math.floor(vi)), 0), n_total - 1)
    hi = min(lo + 1, n_total - 1)
    return lo, hi, vi - math.floor(vi)


def _order_statistics(flat, ranks, panel: int = 1 << 26):
    """Exact order statistics ``sorted(flat)[r] for r in ranks`` of a device float64 vector WITHOUT sorting it (a sort
    needs two more copies of the n_g^2 correlations).  A strided sample brackets every rank; one counting pass gives the
    number of values below the bracket, one gathering pass its members (a few thousand), which are sorted exactly.  A
    bracket that misses its rank (cannot happen for a consistent sample, but ties can crowd it) is widened and retried."""
    import torch

    n = flat.numel()
    stride = max(1, n // (1 << 20))
    sample = torch.sort(flat[::stride]).values
    m = sample.numel()
    out = {}
    for r in sorted(set(ranks)):
        pos = r / max(1, n - 1) * (m - 1)
        margin = max(8.0, 6.0 * math.sqrt(m))  # ~6 sigma of the sample rank
        for _attempt in range(8):
            i0, i1 = int(max(0, math.floor(pos - margin))), int(min(m - 1, math.ceil(pos + margin)))
            a = float(sample[i0].item()) if i0 > 0 else -math.inf
            b = float(sample[i1].item()) if i1 < m - 1 else math.inf
            below = 0
            parts = []
            for p0 in range(0, n, panel):
                x = flat[p0 : p0 + panel]
                below += int((x < a).sum().item())
                parts.append(x[(x >= a) & (x <= b)])
            cand = torch.cat(parts)
            k = r - below
            if 0 <= k < cand.numel() and cand.numel() <= (1 << 27):
                out[r] = float(torch.sort(cand).values[k].item())
                break
            margin *= 4.0
        else:  # pragma: no cover
            raise _lib.IcnvError("ITH: quantile bracket did not converge")
    return [out[r] for r in ranks]


def _group_iqr(block: np.ndarray, device) -> float:
    """IQR of the cell-cell correlation matrix of one group (dense float64 block [n_g, K])."""
    import torch

    lib = _lib.load()
    n, K = block.shape
    free, _ = torch.cuda.mem_get_info(device)
    need = 8 * n * n + 8 * n * K + (3 << 30)  # the matrix, the block, selection scratch (no sort of the n_g^2 entries)
    if need > free:
        raise _lib.IcnvError(
            f"group of {n} cells needs {need / 2**30:.1f} GiB for its correlation matrix; "
            f"{free / 2**30:.1f} GiB are free on {device}"
        )
    X = torch.from_numpy(np.ascontiguousarray(block, dtype=np.float64)).to(device)
    corr = torch.empty((n, n), dtype=torch.float64, device=device)
    work = torch.empty((2 * n,), dtype=torch.float64, device=device)
    _lib.check(
        lib.icnv_row_corrcoef_f64(_lib.ptr(X), n, X.stride(0), K, _lib.ptr(corr), corr.stride(0), _lib.ptr(work), _lib.stream_handle(device)),
        "icnv_row_corrcoef_f64",
    )
    del X
    flat = corr.reshape(-1)
    panel = 1 << 26
    for p0 in range(0, flat.numel(), panel):  # np.percentile of an array with a NaN is NaN
        if bool(torch.isnan(flat[p0 : p0 + panel]).any().item()):
            return float("nan")
    N = n * n
    if N <= (1 << 22):  # small groups: a plain sort is cheaper than the selection passes
        srt = torch.sort(flat).values
        pick = lambda r: float(srt[r].item())  # noqa: E731
        (l75, h75, g75), (l25, h25, g25) = _virtual_index(N, 0.75), _virtual_index(N, 0.25)
        return _lerp_np(pick(l75), pick(h75), g75) - _lerp_np(pick(l25), pick(h25), g25)
    (l75, h75, g75), (l25, h25, g25) = _virtual_index(N, 0.75), _virtual_index(N, 0.25)
    a75, b75, a25, b25 = _order_statistics(flat, [l75, h75, l25, h25])
    return _lerp_np(a75, b75, g75) - _lerp_np(a25, b25, g25)


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
