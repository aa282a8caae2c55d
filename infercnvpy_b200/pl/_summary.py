"""Per-group column means of ``X_cnv`` on the GPU — reference: ``pl/_chromosome_heatmap.py:149-158``
(``np.mean(adata.obsm["X_cnv"][adata.obs[groupby].values == group, :], axis=0)`` per group)."""

from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp

from .. import _lib
from .._engine import allreduce_sums, global_label_order


def group_means(adata, groupby: str = "cnv_leiden", *, use_rep: str = "cnv"):
    """``(groups, means [n_groups, K] float64)``: column means of ``adata.obsm["X_" + use_rep]`` over the cells of every
    group of ``adata.obs[groupby]``, groups in order of first appearance (``Series.unique()``, ``:144``).

    One pass over the matrix: the per-category column-sum kernel of the hot path (``icnv_colsum_csr_f32`` /
    ``icnv_colsum_dense_f32``, fp64 accumulation, deterministic) with the group code as the category; under an
    initialised ``torch.distributed`` group the sums and cell counts of the row shards are all-reduced."""
    import torch

    if not torch.cuda.is_available():
        raise _lib.IcnvError("infercnvpy_b200.pl.group_means needs a CUDA device; there is no CPU fallback")
    X = adata.obsm[f"X_{use_rep}"]
    labels = adata.obs[groupby]
    groups = global_label_order(pd.unique(labels))
    codes = pd.Categorical(labels, categories=groups).codes.astype(np.int32)
    n, K = X.shape
    n_groups = len(groups)
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    stream = _lib.stream_handle(device)
    sums = torch.zeros((n_groups, K), dtype=torch.float64, device=device)
    counts = torch.zeros((n_groups,), dtype=torch.int64, device=device)
    row_cat = torch.from_numpy(codes).to(device)
    if n:
        if sp.issparse(X):
            Xc = X.tocsr()
            if not Xc.has_canonical_format:
                Xc = Xc.copy()
                Xc.sum_duplicates()
            # X_cnv holds float32-rounded values (tl.infercnv): the cast is exact for it, ~6e-8 relative otherwise
            data = torch.from_numpy(np.ascontiguousarray(Xc.data, dtype=np.float32)).to(device)
            indptr = torch.from_numpy(Xc.indptr.astype(np.int64)).to(device)
            indices = torch.from_numpy(Xc.indices.astype(np.int32)).to(device)
            rc = lib.icnv_colsum_csr_f32(_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), n, K, _lib.ptr(row_cat), n_groups,
                                         _lib.ptr(sums), _lib.ptr(counts), stream)
        else:
            Xd = torch.from_numpy(np.ascontiguousarray(np.asarray(X), dtype=np.float32)).to(device)
            rc = lib.icnv_colsum_dense_f32(_lib.ptr(Xd), n, K, K, _lib.ptr(row_cat), n_groups, _lib.ptr(sums), _lib.ptr(counts), stream)
        _lib.check(rc, "icnv_colsum")
    sums, counts = allreduce_sums(sums, counts)
    means = (sums / counts.to(torch.float64)[:, None]).cpu().numpy()
    return list(groups), means


def chromosome_heatmap_summary(adata, *, groupby: str = "cnv_leiden", use_rep: str = "cnv", **kwargs):
    """The data of the reference's summary heatmap (``pl/_chromosome_heatmap.py:90-189``) without the rendering:
    ``dict(groups, mean [n_groups, K], chr_pos, var_group_positions, var_group_labels, vmin, vmax)`` — one row per group
    (the reference repeats every row ten times for scanpy's heatmap, ``:137-156``), chromosome blocks from
    ``adata.uns[use_rep]["chr_pos"]`` (``:158-171``), colour range like ``:161-166``."""
    if groupby == "cnv_leiden" and "cnv_leiden" not in adata.obs.columns:
        raise ValueError("'cnv_leiden' is not in `adata.obs`. Did you run `tl.leiden()`?")
    groups, means = group_means(adata, groupby, use_rep=use_rep)
    chr_pos_dict = dict(sorted(adata.uns[use_rep]["chr_pos"].items(), key=lambda x: x[1]))
    chr_pos = [int(v) for v in chr_pos_dict.values()]
    vmin = kwargs.pop("vmin", None)
    vmax = kwargs.pop("vmax", None)
    return dict(
        groups=groups,
        mean=means,
        chr_pos=chr_pos_dict,
        var_group_positions=list(zip(chr_pos, chr_pos[1:] + [means.shape[1]])),
        var_group_labels=list(chr_pos_dict.keys()),
        vmin=float(np.min(means)) if vmin is None else vmin,
        vmax=float(np.max(means)) if vmax is None else vmax,
    )
