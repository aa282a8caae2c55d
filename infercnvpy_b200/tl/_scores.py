"""``cnv.tl.cnv_score`` — reference: ``/root/reference/src/infercnvpy/tl/_scores.py:14-74``."""

from __future__ import annotations

import warnings
from collections.abc import Mapping
from typing import Any

import numpy as np
import pandas as pd
import scipy.sparse as sp

from .. import _lib
from .._engine import global_label_order, label_scores


def cnv_score(
    adata,
    groupby: str = "cnv_leiden",
    *,
    use_rep: str = "cnv",
    key_added: str = "cnv_score",
    inplace: bool = True,
    obs_key=None,
) -> Mapping[Any, np.number] | None:
    """Assign each group the mean of the absolute CNV values of its cells (GPU).

    Same parameters / return as the reference (``_scores.py:14-74``): per group
    ``mean(abs(X_cnv[group rows, :]))`` over ALL entries (zeros included, ``:66``).  Under an
    initialised ``torch.distributed`` group every rank passes its row shard; the label list is the union over ranks
    (order of first appearance in rank order) and the per-group sums / row counts are reduced over all ranks.
    """
    import torch

    if obs_key is not None:
        warnings.warn(
            "The obs_key argument has been renamed to `groupby` for consistency with "
            "other functions and will be removed in the future. ",
            category=FutureWarning,
            stacklevel=2,
        )
        groupby = obs_key

    if groupby not in adata.obs.columns and groupby == "cnv_leiden":
        raise ValueError("`cnv_leiden` not found in `adata.obs`. Did you run `tl.leiden`?")

    X = adata.obsm[f"X_{use_rep}"]
    groups = adata.obs[groupby]
    # sharded call: every rank must index the same label list (a rank-local pd.unique would misalign the all-reduce)
    clusters = global_label_order(pd.unique(groups))
    codes = pd.Categorical(groups, categories=clusters).codes.astype(np.int32)
    n, K = X.shape

    if not torch.cuda.is_available():
        raise _lib.IcnvError("infercnvpy_b200.tl.cnv_score needs a CUDA device; there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    stream = _lib.stream_handle(device)
    row_abs = torch.empty((n,), dtype=torch.float64, device=device)
    if sp.issparse(X):
        Xc = X.tocsr()
        is64 = Xc.dtype != np.float32
        data = torch.from_numpy(np.ascontiguousarray(Xc.data, dtype=np.float64 if is64 else np.float32)).to(device)
        indptr = torch.from_numpy(Xc.indptr.astype(np.int64)).to(device)
        _lib.check(lib.icnv_rowabs_csr(_lib.ptr(indptr), _lib.ptr(data), int(is64), n, _lib.ptr(row_abs), stream), "icnv_rowabs_csr")
    else:
        Xd = np.asarray(X)
        is64 = Xd.dtype != np.float32
        Xt = torch.from_numpy(np.ascontiguousarray(Xd, dtype=np.float64 if is64 else np.float32)).to(device)
        _lib.check(lib.icnv_rowabs_dense(_lib.ptr(Xt), int(is64), n, K, K, _lib.ptr(row_abs), stream), "icnv_rowabs_dense")
    labels = torch.from_numpy(codes).to(device)
    score = label_scores(lib, row_abs, labels, len(clusters), K, device).cpu().numpy()
    cluster_score = {c: score[i] for i, c in enumerate(clusters)}

    if inplace:
        adata.obs[key_added] = score[codes]
    else:
        return cluster_score
