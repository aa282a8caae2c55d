"""``cnv.tl.pca`` — reference: ``/root/reference/src/infercnvpy/tl/__init__.py:33-75`` (a one-call wrapper around
``scanpy.tl.pca(X_cnv, svd_solver="arpack", zero_center=False)`` -> scikit-learn ``TruncatedSVD(algorithm="arpack")``).

Here: Gram matrix ``X^T X`` (fp64, ``icnv_gram_f32``; one all-reduce of ``K x K`` when rows are sharded), symmetric
eigen-decomposition of the ``K x K`` matrix (cuSOLVER through ``torch.linalg.eigh`` — a plain library call on a
``K <= ~10^4`` matrix), scikit-learn's ``svd_flip`` sign convention, projection ``X V`` (``icnv_project_f32``).
``X V`` equals TruncatedSVD's ``U * Sigma`` up to solver tolerance.  Parity unpinned by the reference.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .. import _lib


def _device():
    import torch

    if not torch.cuda.is_available():
        raise _lib.IcnvError("infercnvpy_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _allreduce_sum(t):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(t)
        else:
            c = t.cpu()
            dist.all_reduce(c)
            t.copy_(c)
    return t


_allreduce = _allreduce_sum  # name used by pp/_neighbors.py


def matrix_to_device_dense(X, device):
    """``obsm`` matrix (dense / CSR / CSC, any float dtype) -> device float32 ``[n, K]``."""
    import torch

    lib = _lib.load()
    n, K = X.shape
    if sp.issparse(X):
        Xc = X.tocsr()
        is64 = Xc.dtype != np.float32
        data = torch.from_numpy(np.ascontiguousarray(Xc.data, dtype=np.float64 if is64 else np.float32)).to(device)
        indptr = torch.from_numpy(Xc.indptr.astype(np.int64)).to(device)
        indices = torch.from_numpy(Xc.indices.astype(np.int32)).to(device)
        dense = torch.empty((n, K), dtype=torch.float32, device=device)
        _lib.check(
            lib.icnv_csr_to_dense_f32(_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), int(is64), n, K, _lib.ptr(dense), K,
                                      _lib.stream_handle(device)),
            "icnv_csr_to_dense_f32",
        )
        return dense
    if isinstance(X, torch.Tensor):
        return X.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(X), dtype=np.float32)).to(device)


def pca_device(Xd, n_comps: int, zero_center: bool = False, n_total: int | None = None, reduce: bool = True):
    """Device-level PCA: ``Xd`` float32 ``[n, K]`` (this rank's rows) -> ``(Y [n, n_comps] float32, V [K, n_comps] f64,
    singular values)``.  ``reduce=False`` keeps the decomposition local even inside a process group (tests)."""
    import torch

    lib = _lib.load()
    device = Xd.device
    n, K = Xd.shape
    stream = _lib.stream_handle(device)
    C = torch.empty((K, K), dtype=torch.float64, device=device)
    _lib.check(lib.icnv_gram_f32(_lib.ptr(Xd), n, Xd.stride(0), K, _lib.ptr(C), stream), "icnv_gram_f32")
    _allreduce = _allreduce_sum if reduce else (lambda t: t)
    _allreduce(C)
    mu = None
    if zero_center:
        nt = torch.tensor([float(n)], dtype=torch.float64, device=device)
        _allreduce(nt)
        # column means from the column-sum kernel of the hot path
        sums = torch.empty((1, K), dtype=torch.float64, device=device)
        counts = torch.empty((1,), dtype=torch.int64, device=device)
        _lib.check(lib.icnv_colsum_dense_f32(_lib.ptr(Xd), n, Xd.stride(0), K, None, 1, _lib.ptr(sums), _lib.ptr(counts), stream))
        _allreduce(sums)
        mu = (sums[0] / nt).contiguous()
        C = C - nt * torch.outer(mu, mu)
    evals, evecs = torch.linalg.eigh(C)
    V = evecs[:, -n_comps:].flip(1).contiguous()          # descending singular value order
    sv = evals[-n_comps:].flip(0).clamp_min(0).sqrt()
    # sklearn.utils.extmath.svd_flip(u, v, u_based_decision=False): largest |entry| of every right vector positive
    idx = V.abs().argmax(dim=0)
    signs = torch.sign(V[idx, torch.arange(n_comps, device=device)])
    signs[signs == 0] = 1
    V = (V * signs[None, :]).contiguous()
    Y = torch.empty((n, n_comps), dtype=torch.float32, device=device)
    _lib.check(
        lib.icnv_project_f32(_lib.ptr(Xd), n, Xd.stride(0), K, _lib.ptr(V), n_comps, _lib.ptr(mu), _lib.ptr(Y), stream),
        "icnv_project_f32",
    )
    return Y, V, sv


def pca(
    adata,
    svd_solver: str = "arpack",
    zero_center: bool = False,
    inplace: bool = True,
    use_rep: str = "cnv",
    key_added: str = "cnv_pca",
    **kwargs,
) -> np.ndarray | None:
    """Compute the PCA on the result of :func:`infercnvpy_b200.tl.infercnv` (GPU).

    Same parameters / keys / errors as the reference (``tl/__init__.py:33-75``); ``svd_solver`` is accepted for
    compatibility (the decomposition is always the exact symmetric eigenproblem of the Gram matrix);
    ``n_comps`` (scanpy keyword, default ``min(50, min(shape) - 1)``) may be passed through ``**kwargs``; other scanpy
    keywords raise ``TypeError`` unless they carry the value implemented here.  Under an initialised ``torch.distributed``
    group every rank passes its row shard: the ``K x K`` Gram matrix is all-reduced and every rank projects its rows.
    """
    if f"X_{use_rep}" not in adata.obsm:
        raise KeyError(f"X_{use_rep} is not in adata.obsm. Did you run `tl.infercnv`?")
    X = adata.obsm[f"X_{use_rep}"]
    n, K = X.shape
    from ..pp._neighbors import _unsupported_kwargs

    n_comps = kwargs.pop("n_comps", None)
    _unsupported_kwargs("pca", kwargs, {"random_state": None, "dtype": ("float32",), "chunked": (False,), "mask_var": (None,),
                                        "use_highly_variable": (None, False), "return_info": (False,), "copy": (False,)})
    if n_comps is None:
        n_comps = min(50, min(n, K) - 1)
    if not 1 <= n_comps <= min(64, K):
        raise ValueError(f"n_comps must be in [1, {min(64, K)}]")
    device = _device()
    Xd = matrix_to_device_dense(X, device)
    Y, _, _ = pca_device(Xd, int(n_comps), zero_center=bool(zero_center))
    pca_res = Y.cpu().numpy()
    if inplace:
        adata.obsm[f"X_{key_added}"] = pca_res
    else:
        return pca_res
