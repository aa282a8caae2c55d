"""``cnv.pp.neighbors`` — reference: ``/root/reference/src/infercnvpy/pp/__init__.py:8-43``, a one-call wrapper around
``scanpy.pp.neighbors(use_rep="X_cnv_pca", key_added="cnv_neighbors")`` (euclidean kNN, ``method="umap"``:
umap-learn's ``fuzzy_simplicial_set`` with ``set_op_mix_ratio=1``, ``local_connectivity=1``).

Here: exact kNN on the device (``icnv_knn_f32``; scanpy switches to approximate pynndescent above 4096 cells), per-row
sigma/rho bisection + membership strengths (``icnv_fuzzy_rows``, umap-learn's published algorithm), fuzzy union
``A + A^T - A * A^T``.  Parity unpinned by the reference; the tests compare with scikit-learn kNN
and a numpy restatement of umap-learn's routine.
"""

from __future__ import annotations

import logging

import numpy as np
import scipy.sparse as sp

from .. import _lib

log = logging.getLogger("infercnvpy_b200")


def knn_device(P, k: int, q0: int = 0, nq: int | None = None):
    """Exact euclidean kNN of rows ``[q0, q0+nq)`` of ``P`` (device float32 ``[n, d]``, ``d <= 64``) against all rows:
    tensor-core distance GEMM (tcgen05, 3xTF32) + exact re-rank (``csrc/icnv_knn.cu``).
    Returns ``(idx [nq, k] int32, dist [nq, k] float32)``; column 0 is the query itself."""
    import torch

    lib = _lib.load()
    n, d = P.shape
    nq = n - q0 if nq is None else nq
    assert P.dtype == torch.float32 and P.stride(1) == 1
    idx = torch.empty((nq, k), dtype=torch.int32, device=P.device)
    d2 = torch.empty((nq, k), dtype=torch.float32, device=P.device)
    ws = torch.empty((int(lib.icnv_knn_workspace_bytes(n, nq)),), dtype=torch.uint8, device=P.device)
    _lib.check(
        lib.icnv_knn_f32(_lib.ptr(P), n, d, P.stride(0), q0, nq, k, _lib.ptr(idx), _lib.ptr(d2), k, _lib.ptr(ws), _lib.stream_handle(P.device)),
        "icnv_knn_f32",
    )
    return idx, d2.clamp_min(0).sqrt()


def fuzzy_graph_device(idx, dist, n_total: int, row0: int = 0):
    """umap-learn ``fuzzy_simplicial_set`` on device kNN lists -> symmetric connectivities as COO
    ``(rows, cols, vals)`` device tensors (both directions present)."""
    import torch

    from ..tl._pca import _allreduce

    lib = _lib.load()
    n, k = dist.shape
    device = dist.device
    tot = torch.stack([dist.double().sum(), torch.tensor(float(dist.numel()), dtype=torch.float64, device=device)])
    _allreduce(tot)
    mean_all = float(tot[0] / tot[1])
    vals = torch.empty((n, k), dtype=torch.float32, device=device)
    sigma = torch.empty((n,), dtype=torch.float32, device=device)
    rho = torch.empty((n,), dtype=torch.float32, device=device)
    _lib.check(
        lib.icnv_fuzzy_rows(_lib.ptr(dist), _lib.ptr(idx), n, k, row0, mean_all, _lib.ptr(vals), _lib.ptr(sigma), _lib.ptr(rho),
                            _lib.stream_handle(device)),
        "icnv_fuzzy_rows",
    )
    rows = (torch.arange(n, device=device, dtype=torch.int64) + row0)[:, None].expand(n, k).reshape(-1)
    cols = idx.reshape(-1).to(torch.int64)
    v = vals.reshape(-1)
    keep = v > 0
    return rows[keep], cols[keep], v[keep]


def symmetrize(rows, cols, vals, n_total: int):
    """Fuzzy set union ``A + A^T - A * A^T`` (set_op_mix_ratio = 1) of a directed COO graph; device tensors."""
    import torch

    lo, hi = torch.minimum(rows, cols), torch.maximum(rows, cols)
    key = lo * n_total + hi
    uniq, inv = torch.unique(key, return_inverse=True)
    a = torch.zeros(uniq.numel(), dtype=torch.float32, device=vals.device)
    b = torch.zeros_like(a)
    fwd = rows < cols
    a.scatter_(0, inv[fwd], vals[fwd])       # value of the (lo -> hi) direction, 0 if absent
    b.scatter_(0, inv[~fwd], vals[~fwd])     # value of the (hi -> lo) direction
    p = a + b - a * b
    ulo, uhi = uniq // n_total, uniq % n_total
    off = ulo != uhi
    r = torch.cat([ulo[off], uhi[off]])
    c = torch.cat([uhi[off], ulo[off]])
    w = torch.cat([p[off], p[off]])
    return r, c, w


def allgather_rows(t):
    """Concatenate the row shards ``t`` (``[n_local, ...]``, same trailing shape on every rank) of all ranks in rank order.
    Returns ``(full, row0)`` with ``row0`` the global index of this rank's first row.  Identity without a process group.
    NCCL gathers padded device blocks over NVLink; a gloo group (CPU tests) goes through host tensors."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t, 0
    world, rank = dist.get_world_size(), dist.get_rank()
    nccl = dist.get_backend() == "nccl"
    dev = t.device if nccl else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=dev))
    sizes = [int(x.item()) for x in sizes]
    n_max = max(sizes)
    src = t if nccl else t.cpu()
    pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[: t.shape[0]] = src
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad)
    full = torch.cat([b[:n] for b, n in zip(blocks, sizes)], dim=0).to(t.device)
    return full, sum(sizes[:rank])


def neighbors_device(P_local, n_neighbors: int):
    """kNN + fuzzy graph for this rank's rows of the PCA coordinates (device float32 ``[n_local, d]``).  Under an
    initialised process group the coordinates of all ranks are all-gathered (N x d floats: 200 MB at 1M cells x 50), every
    rank searches the neighbours of ITS rows among all N points, and the directed kNN lists are all-gathered again so that
    every rank holds the whole symmetric graph (<= 2 k N edges) for the clustering.
    Returns ``dict(idx, dist`` (local rows, global ids)``, row0, n_total, coo=(rows, cols, vals))``."""
    P_all, row0 = allgather_rows(P_local.contiguous())
    n_total, n_local = P_all.shape[0], P_local.shape[0]
    idx, dist = knn_device(P_all, n_neighbors, q0=row0, nq=n_local)
    rows, cols, vals = fuzzy_graph_device(idx, dist, n_total, row0=row0)
    rows_all, _ = allgather_rows(rows)
    cols_all, _ = allgather_rows(cols)
    vals_all, _ = allgather_rows(vals)
    r, c, w = symmetrize(rows_all, cols_all, vals_all, n_total)
    return dict(idx=idx, dist=dist, row0=row0, n_total=n_total, coo=(r, c, w))


def _unsupported_kwargs(fn: str, kwargs: dict, accepted: dict):
    """The reference wrappers forward ``**kwargs`` to scanpy; here only the listed values are implemented.  Anything else
    would silently give a different result than the caller asked for -> TypeError."""
    for key, val in kwargs.items():
        if key not in accepted:
            raise TypeError(f"{fn}() got an unsupported keyword argument {key!r} (supported: {sorted(accepted)})")
        ok = accepted[key]
        if ok is not None and val not in ok:
            raise TypeError(f"{fn}(): {key}={val!r} is not supported (supported: {ok})")


def neighbors(
    adata,
    use_rep: str = "cnv_pca",
    key_added: str = "cnv_neighbors",
    inplace: bool = True,
    **kwargs,
):
    """Compute the neighborhood graph on the PCA of the CNV matrix (GPU).

    Same parameters and keys as the reference (``pp/__init__.py:8-43``): distances in
    ``.obsp[key_added + "_distances"]``, connectivities in ``.obsp[key_added + "_connectivities"]``, parameters in
    ``.uns[key_added]``.  ``n_neighbors`` (default 15, counts the cell itself like scanpy) may be passed as keyword; other
    scanpy keywords are accepted only with the values implemented here (euclidean, method "umap") and raise ``TypeError``
    otherwise.  With ``inplace=False`` the two matrices are returned instead.  Under an initialised ``torch.distributed``
    group every rank passes its row shard and receives ITS rows of the global matrices (``[n_local, n_total]``).
    """
    import torch

    from ..tl._pca import _device, pca

    if f"X_{use_rep}" not in adata.obsm and use_rep == "cnv_pca":
        log.warning("X_cnv_pca not found in adata.obsm. Computing PCA with default parameters")
        pca(adata)
    n_neighbors = int(kwargs.pop("n_neighbors", 15))
    _unsupported_kwargs("neighbors", kwargs, {"metric": ("euclidean",), "method": ("umap",), "knn": (True,), "n_pcs": (None,),
                                              "random_state": None, "transformer": (None,), "copy": (False,)})
    P_host = np.ascontiguousarray(np.asarray(adata.obsm[f"X_{use_rep}"]), dtype=np.float32)
    n, d = P_host.shape
    if d > 64:
        raise ValueError("at most 64 dimensions are supported for the neighbour search")
    device = _device()
    if not 2 <= n_neighbors <= 20:
        raise ValueError("n_neighbors must be in [2, min(20, n_obs)]")
    g = neighbors_device(torch.from_numpy(P_host).to(device), n_neighbors)
    if n_neighbors > g["n_total"]:
        raise ValueError("n_neighbors must be in [2, min(20, n_obs)]")
    idx, dist, row0, n_total = g["idx"], g["dist"], g["row0"], g["n_total"]
    r, c, w = g["coo"]
    # this rank's rows of the (global) symmetric connectivity matrix: [n, n_total] (square when there is one rank)
    mine = (r >= row0) & (r < row0 + n)
    conn = sp.csr_matrix(
        (w[mine].cpu().numpy(), ((r[mine] - row0).cpu().numpy(), c[mine].cpu().numpy())), shape=(n, n_total), dtype=np.float32
    )
    # distances: the n_neighbors - 1 true neighbours of every cell (scanpy drops the cell itself)
    ii = np.repeat(np.arange(n), n_neighbors - 1)
    distances = sp.csr_matrix(
        (dist[:, 1:].reshape(-1).cpu().numpy().astype(np.float64), (ii, idx[:, 1:].reshape(-1).cpu().numpy())), shape=(n, n_total)
    )
    if inplace:
        adata.obsp[f"{key_added}_distances"] = distances
        adata.obsp[f"{key_added}_connectivities"] = conn
        adata.uns[key_added] = {
            "connectivities_key": f"{key_added}_connectivities",
            "distances_key": f"{key_added}_distances",
            "params": {"n_neighbors": n_neighbors, "method": "umap", "metric": "euclidean", "use_rep": f"X_{use_rep}"},
        }
        if n_total != n:
            adata.uns[key_added]["shard"] = {"row0": int(row0), "n_total": int(n_total)}
    else:
        return distances, conn
