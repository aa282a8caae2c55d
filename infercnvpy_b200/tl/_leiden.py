"""``cnv.tl.leiden`` — reference: ``/root/reference/src/infercnvpy/tl/__init__.py:13-30``, a one-call wrapper around
``scanpy.tl.leiden(neighbors_key="cnv_neighbors", key_added="cnv_leiden")`` -> ``leidenalg.find_partition(
RBConfigurationVertexPartition, weights, resolution 1, seed 0, until convergence)``.

leidenalg is sequential, randomised and order dependent; label-for-label equality with a parallel implementation is
not attainable and the reference's tests assert nothing about the clustering: parity unpinned.
Here: the Leiden scheme on the same quality function (RB-configuration modularity, resolution ``gamma``) — synchronous
local-moving sweeps, refinement sweeps inside every community (sub-communities are connected by construction) and
aggregation of the refined partition, all on the device (``icnv_community_sweep``), repeated until a level changes
nothing.  Deterministic (no random visiting order).  Labels are strings, numbered by decreasing cluster size like
leidenalg's; the tests compare quality and ARI with a sequential CPU restatement of Leiden (``oracle/leiden_oracle.py``).
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


def _sweeps(lib, graph, kdeg, two_m, comm, bound, gamma, work, stats, stream, max_sweeps, sweep0=0, tol_nodes: int = 0):
    """Synchronous sweeps of ``icnv_community_sweep`` until two consecutive ones (both halves of the checkerboard) move
    at most ``tol_nodes`` nodes (0: nothing).  Returns ``(assignment, any node moved, sweeps used)``."""
    import torch

    indptr, indices, w = graph
    n = indptr.numel() - 1
    comm_new = torch.empty_like(comm)
    moved_any, quiet, used = False, 0, 0
    for sweep in range(max_sweeps):
        _lib.check(
            lib.icnv_community_sweep(_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(w), _lib.ptr(kdeg), _lib.ptr(comm), _lib.ptr(bound),
                                     n, two_m, float(gamma), sweep0 + sweep, _lib.ptr(work), _lib.ptr(comm_new), _lib.ptr(stats), stream),
            "icnv_community_sweep",
        )
        comm, comm_new = comm_new, comm
        used += 1
        m, _, overflow = (int(v) for v in stats.tolist())
        if overflow:
            raise _lib.IcnvError("icnv_community_sweep: a node has more distinct neighbouring communities than the hash table holds")
        moved_any = moved_any or m > 0
        if m > tol_nodes:
            quiet = 0
        else:
            quiet += 1
            if quiet >= 2:
                break
    return comm, moved_any, used


def leiden_device(indptr, indices, w, gamma: float = 1.0, max_levels: int = 32, max_sweeps: int = 200, refine: bool = True,
                  move_tol: float = 1e-3):
    """Leiden scheme on a symmetric CSR graph (device tensors): local moving, refinement inside every community,
    aggregation of the refined sub-communities (which start the next level in their community); repeated until a level
    changes nothing.  ``refine=False`` gives plain multilevel Louvain.  Returns int64 labels (device), numbered arbitrarily.

    ``move_tol``: the local-moving phase of a level stops when two consecutive sweeps move at most ``move_tol * n`` nodes
    (rounded down: exact convergence below 1000 nodes).  Under the synchronous update a fraction of a percent of the nodes of
    a 1M-node kNN graph keeps flipping between two nearly equivalent communities: with ``move_tol = 0`` the first level runs
    into ``max_sweeps`` (200 sweeps, 443 of 573 ms), with 1e-3 it stops after 87 (337 ms in total) at the same quality
    (0.916666, all 12 planted clones; ``tools/leiden_profile.py``)."""
    import torch

    lib = _lib.load()
    device = w.device
    stream = _lib.stream_handle(device)
    n0 = indptr.numel() - 1
    node_of = torch.arange(n0, device=device, dtype=torch.int64)   # aggregated node every original node sits in
    comm = torch.arange(n0, device=device, dtype=torch.int32)       # community of every aggregated node
    stats = torch.zeros(3, dtype=torch.int32, device=device)
    sweep0 = 0
    for _level in range(max_levels):
        n = indptr.numel() - 1
        graph = (indptr, indices, w)
        kdeg = torch.empty(n, dtype=torch.float64, device=device)
        _lib.check(lib.icnv_weighted_degree(_lib.ptr(indptr), _lib.ptr(w), n, _lib.ptr(kdeg), stream), "icnv_weighted_degree")
        two_m = float(kdeg.sum())
        if two_m <= 0:
            break
        work = torch.empty(int(lib.icnv_community_sweep_work_bytes(n)), dtype=torch.uint8, device=device)
        # ---- 1. local moving (from the communities inherited from the previous level)
        comm, moved, used = _sweeps(lib, graph, kdeg, two_m, comm, None, gamma, work, stats, stream, max_sweeps, sweep0,
                                    tol_nodes=int(move_tol * n))
        sweep0 += used
        # ---- 2. refinement: singletons merge inside their community
        if refine:
            sub = torch.arange(n, device=device, dtype=torch.int32)
            sub, _, used = _sweeps(lib, graph, kdeg, two_m, sub, comm, gamma, work, stats, stream, max_sweeps, sweep0)
            sweep0 += used
        else:
            sub = comm
        uniq, inv = torch.unique(sub.long(), return_inverse=True)
        nc = uniq.numel()
        if nc == n and not moved:
            break  # nothing moved and nothing merged: converged
        # ---- 3. aggregate by sub-community; every new node starts in its members' community
        first = torch.full((nc,), n, dtype=torch.int64, device=device).scatter_reduce_(0, inv, torch.arange(n, device=device), "amin")
        new_comm_raw = comm.long()[first]
        _, new_comm = torch.unique(new_comm_raw, return_inverse=True)
        node_of = inv[node_of]
        if nc == n:
            comm = new_comm.to(torch.int32)
            if not refine:
                break
            continue
        rows = torch.repeat_interleave(torch.arange(n, device=device), (indptr[1:] - indptr[:-1]))
        cr, cc = inv[rows], inv[indices.long()]
        key = cr * nc + cc
        ukey, kinv = torch.unique(key, return_inverse=True)
        wsum = torch.zeros(ukey.numel(), dtype=torch.float64, device=device).scatter_add_(0, kinv, w.double()).float()
        nr = ukey // nc
        counts = torch.bincount(nr, minlength=nc)
        indptr = torch.zeros(nc + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(counts, 0)
        indices = (ukey % nc).to(torch.int32)
        w = wsum
        comm = new_comm.to(torch.int32)
    return comm.long()[node_of]


def louvain_device(indptr, indices, w, gamma: float = 1.0, max_levels: int = 32, max_sweeps: int = 200):
    """Multilevel local moving + aggregation without the refinement step (kept for comparison in the tests)."""
    return leiden_device(indptr, indices, w, gamma, max_levels, max_sweeps, refine=False)


def leiden(
    adata,
    neighbors_key: str = "cnv_neighbors",
    key_added: str = "cnv_leiden",
    inplace: bool = True,
    **kwargs,
):
    """Cluster the CNV neighbourhood graph by modularity optimisation (GPU).

    Same parameters / keys as the reference (``tl/__init__.py:13-30``); ``resolution`` (default 1) may be passed as
    keyword.  Writes ``adata.obs[key_added]`` (categorical of strings, "0" = largest cluster) and ``adata.uns[key_added]``;
    with ``inplace=False`` a COPY of ``adata`` carrying both is returned and ``adata`` is left alone — the reference passes
    ``copy=not inplace`` to ``scanpy.tl.leiden`` (``tl/__init__.py:28``), which returns the annotated copy.
    """
    import torch

    from ._pca import _device

    from ..pp._neighbors import _unsupported_kwargs, allgather_rows

    resolution = float(kwargs.pop("resolution", 1.0))
    _unsupported_kwargs("leiden", kwargs, {"random_state": None, "n_iterations": (-1,), "directed": None, "use_weights": (True,),
                                           "flavor": ("leidenalg",), "restrict_to": (None,), "adjacency": (None,),
                                           "partition_type": (None,), "obsp": (None,), "copy": (False,)})
    if neighbors_key not in adata.uns:
        raise KeyError(f"No neighbors graph under {neighbors_key!r}. Did you run `pp.neighbors`?")
    ckey = adata.uns[neighbors_key].get("connectivities_key", f"{neighbors_key}_connectivities")
    A = sp.csr_matrix(adata.obsp[ckey])
    device = _device()
    shard = adata.uns[neighbors_key].get("shard")
    if shard is not None:
        # row shard of the global graph: gather the COO triples of all ranks, cluster the whole graph on every rank
        # (deterministic -> identical labels) and keep this rank's rows
        coo = A.tocoo()
        r = torch.from_numpy(coo.row.astype(np.int64) + int(shard["row0"])).to(device)
        c = torch.from_numpy(coo.col.astype(np.int64)).to(device)
        v = torch.from_numpy(coo.data.astype(np.float32)).to(device)
        r, _ = allgather_rows(r)
        c, _ = allgather_rows(c)
        v, _ = allgather_rows(v)
        n_total = int(shard["n_total"])
        order = torch.argsort(r * n_total + c)
        counts = torch.bincount(r, minlength=n_total)
        indptr = torch.zeros(n_total + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(counts, 0)
        indices, w = c[order].to(torch.int32), v[order]
    else:
        indptr, indices, w = _csr_to_device(A, device)
    labels = leiden_device(indptr, indices, w, gamma=resolution)
    # number clusters by decreasing size (leidenalg convention)
    counts = torch.bincount(labels)
    order = torch.argsort(counts, descending=True, stable=True)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=device)
    lab = rank[labels]
    if shard is not None:
        lab = lab[int(shard["row0"]) : int(shard["row0"]) + adata.shape[0]]
    lab = lab.cpu().numpy()
    cats = [str(i) for i in range(int(order.numel()))]
    result = pd.Categorical([str(i) for i in lab], categories=cats)
    target = adata if inplace else adata.copy()
    target.obs[key_added] = result
    target.uns[key_added] = {"params": {"resolution": resolution, "random_state": 0, "n_iterations": -1}}
    if not inplace:
        return target
