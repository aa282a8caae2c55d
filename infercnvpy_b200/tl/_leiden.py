"""``cnv.tl.leiden`` — reference: ``/root/reference/src/infercnvpy/tl/__init__.py:13-30``, a one-call wrapper around
``scanpy.tl.leiden(neighbors_key="cnv_neighbors", key_added="cnv_leiden")`` -> ``leidenalg.find_partition(
RBConfigurationVertexPartition, weights, resolution 1, seed 0, until convergence)``.

leidenalg is sequential, randomised and order dependent; label-for-label equality with a parallel implementation is
not attainable and the reference's tests assert nothing about the clustering: parity unpinned.
Here: multilevel optimisation of the same quality function (RB-configuration modularity, resolution ``gamma``) —
synchronous local-moving sweeps on the device (``icnv_louvain_sweep``) + graph aggregation, repeated until no node
moves at any level.  The Leiden refinement step (guaranteeing well-connected communities) is not implemented yet.
Labels are strings, numbered by decreasing cluster size like leidenalg's.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp

from .. import _lib


def _csr_to_device(A, device):
    import torch

    A = A.tocsr()
    return (
        torch.from_numpy(A.indptr.astype(np.int64)).to(device),
        torch.from_numpy(A.indices.astype(np.int32)).to(device),
        torch.from_numpy(np.ascontiguousarray(A.data, dtype=np.float32)).to(device),
    )


def modularity_device(indptr, indices, w, labels, gamma: float = 1.0) -> float:
    """RB-configuration quality / 2m of a labelling (device tensors) — used by tests and for reporting."""
    import torch

    n = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=w.device), (indptr[1:] - indptr[:-1]))
    k = torch.zeros(n, dtype=torch.float64, device=w.device).scatter_add_(0, rows, w.double())
    two_m = k.sum()
    same = labels[rows] == labels[indices.long()]
    inside = w.double()[same].sum()
    nl = int(labels.max().item()) + 1
    ctot = torch.zeros(nl, dtype=torch.float64, device=w.device).scatter_add_(0, labels.long(), k)
    return float((inside - gamma * (ctot * ctot).sum() / two_m) / two_m)


def louvain_device(indptr, indices, w, gamma: float = 1.0, max_levels: int = 20, max_sweeps: int = 200):
    """Multilevel local moving + aggregation; returns int64 labels (device) numbered arbitrarily."""
    import torch

    lib = _lib.load()
    device = w.device
    stream = _lib.stream_handle(device)
    n0 = indptr.numel() - 1
    labels = torch.arange(n0, device=device, dtype=torch.int64)  # community of every original node
    for _level in range(max_levels):
        n = indptr.numel() - 1
        kdeg = torch.empty(n, dtype=torch.float64, device=device)
        _lib.check(lib.icnv_weighted_degree(_lib.ptr(indptr), _lib.ptr(w), n, _lib.ptr(kdeg), stream), "icnv_weighted_degree")
        two_m = float(kdeg.sum())
        if two_m <= 0:
            break
        comm = torch.arange(n, device=device, dtype=torch.int32)
        comm_new = torch.empty_like(comm)
        ctot = torch.empty(n, dtype=torch.float64, device=device)
        n_moved = torch.zeros(1, dtype=torch.int32, device=device)
        moved_any = False
        quiet = 0
        for sweep in range(max_sweeps):
            _lib.check(
                lib.icnv_louvain_sweep(_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(w), _lib.ptr(kdeg), _lib.ptr(comm), _lib.ptr(ctot),
                                       n, two_m, float(gamma), sweep, _lib.ptr(comm_new), _lib.ptr(n_moved), stream),
                "icnv_louvain_sweep",
            )
            comm, comm_new = comm_new, comm
            m = int(n_moved.item())
            if m > 0:
                moved_any = True
                quiet = 0
            else:
                quiet += 1
                if quiet >= 2:  # both halves of the checkerboard had their turn
                    break
        if not moved_any:
            break
        # ---- aggregate: relabel communities 0..nc-1 and contract the graph
        uniq, inv = torch.unique(comm.long(), return_inverse=True)
        nc = uniq.numel()
        labels = inv[labels]
        if nc == n:
            break
        rows = torch.repeat_interleave(torch.arange(n, device=device), (indptr[1:] - indptr[:-1]))
        cr, cc = inv[rows], inv[indices.long()]
        key = cr * nc + cc
        ukey, kinv = torch.unique(key, return_inverse=True)
        wsum = torch.zeros(ukey.numel(), dtype=torch.float32, device=device).scatter_add_(0, kinv, w)
        nr = ukey // nc
        counts = torch.bincount(nr, minlength=nc)
        indptr = torch.zeros(nc + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(counts, 0)
        indices = (ukey % nc).to(torch.int32)
        w = wsum
    return labels


def leiden(
    adata,
    neighbors_key: str = "cnv_neighbors",
    key_added: str = "cnv_leiden",
    inplace: bool = True,
    **kwargs,
):
    """Cluster the CNV neighbourhood graph by modularity optimisation (GPU).

    Same parameters / keys as the reference (``tl/__init__.py:13-30``); ``resolution`` (default 1) may be passed as
    keyword.  Writes ``adata.obs[key_added]`` (categorical of strings, "0" = largest cluster); with
    ``inplace=False`` the labels are returned as a ``pandas.Categorical`` instead.
    """
    import torch

    from ._pca import _device

    resolution = float(kwargs.pop("resolution", 1.0))
    if neighbors_key not in adata.uns:
        raise KeyError(f"No neighbors graph under {neighbors_key!r}. Did you run `pp.neighbors`?")
    ckey = adata.uns[neighbors_key].get("connectivities_key", f"{neighbors_key}_connectivities")
    A = adata.obsp[ckey]
    device = _device()
    indptr, indices, w = _csr_to_device(sp.csr_matrix(A), device)
    labels = louvain_device(indptr, indices, w, gamma=resolution)
    # number clusters by decreasing size (leidenalg convention)
    counts = torch.bincount(labels)
    order = torch.argsort(counts, descending=True, stable=True)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=device)
    lab = rank[labels].cpu().numpy()
    cats = [str(i) for i in range(int(lab.max()) + 1)] if lab.size else []
    result = pd.Categorical([str(i) for i in lab], categories=cats)
    if inplace:
        adata.obs[key_added] = result
        adata.uns[key_added] = {"params": {"resolution": resolution, "random_state": 0, "n_iterations": -1}}
    else:
        return result
