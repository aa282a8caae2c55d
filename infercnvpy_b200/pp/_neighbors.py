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
    tensor-core distance GEMM (tcgen05, 3xTF32) + exact re-rank (``csrc/icnv_knn.cu``).  ``q0`` must be a multiple of 128.
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
    ``.uns[key_added]``.  ``n_neighbors`` (default 15, counts the cell itself like scanpy) may be passed as keyword.
    With ``inplace=False`` the two matrices are returned instead.
    """
    import torch

    from ..tl._pca import _device, pca

    if f"X_{use_rep}" not in adata.obsm and use_rep == "cnv_pca":
        log.warning("X_cnv_pca not found in adata.obsm. Computing PCA with default parameters")
        pca(adata)
    n_neighbors = int(kwargs.pop("n_neighbors", 15))
    P_host = np.ascontiguousarray(np.asarray(adata.obsm[f"X_{use_rep}"]), dtype=np.float32)
    n, d = P_host.shape
    if not 2 <= n_neighbors <= min(20, n):
        raise ValueError("n_neighbors must be in [2, min(20, n_obs)]")
    if d > 64:
        raise ValueError("at most 64 dimensions are supported for the neighbour search")
    device = _device()
    P = torch.from_numpy(P_host).to(device)
    idx, dist = knn_device(P, n_neighbors)
    rows, cols, vals = fuzzy_graph_device(idx, dist, n)
    r, c, w = symmetrize(rows, cols, vals, n)
    conn = sp.csr_matrix((w.cpu().numpy(), (r.cpu().numpy(), c.cpu().numpy())), shape=(n, n), dtype=np.float32)
    # distances: the n_neighbors - 1 true neighbours of every cell (scanpy drops the cell itself)
    ii = np.repeat(np.arange(n), n_neighbors - 1)
    distances = sp.csr_matrix(
        (dist[:, 1:].reshape(-1).cpu().numpy().astype(np.float64), (ii, idx[:, 1:].reshape(-1).cpu().numpy())), shape=(n, n)
    )
    if inplace:
        adata.obsp[f"{key_added}_distances"] = distances
        adata.obsp[f"{key_added}_connectivities"] = conn
        adata.uns[key_added] = {
            "connectivities_key": f"{key_added}_connectivities",
            "distances_key": f"{key_added}_distances",
            "params": {"n_neighbors": n_neighbors, "method": "umap", "metric": "euclidean", "use_rep": f"X_{use_rep}"},
        }
    else:
        return distances, conn
