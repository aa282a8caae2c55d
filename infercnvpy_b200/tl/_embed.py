"""``cnv.tl.umap`` / ``cnv.tl.tsne`` — reference: ``/root/reference/src/infercnvpy/tl/__init__.py:78-144``, thin wrappers
around ``scanpy.tl.umap(neighbors_key="cnv_neighbors")`` (umap-learn's ``simplicial_set_embedding``) and
``scanpy.tl.tsne(use_rep="X_cnv_pca")`` (scikit-learn's ``TSNE``).  Both optimisers are randomised upstream and the
reference's tests assert nothing about them: parity unpinned.  Here the published algorithms run on the
device (``csrc/icnv_embed.cu``); the tests check neighbourhood preservation (scikit-learn's ``trustworthiness``) and
that planted clones stay separated.
"""

from __future__ import annotations

import logging

import numpy as np
import scipy.sparse as sp

from .. import _lib
from ._pca import _device

log = logging.getLogger("infercnvpy_b200")


def find_ab_params(spread: float = 1.0, min_dist: float = 0.5):
    """umap-learn's ``find_ab_params``: least-squares fit of ``1 / (1 + a x^(2b))`` to the target membership curve
    (1 below ``min_dist``, ``exp(-(x - min_dist) / spread)`` above)."""
    from scipy.optimize import curve_fit

    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xv = np.linspace(0, spread * 3, 300)
    yv = np.where(xv < min_dist, 1.0, np.exp(-(xv - min_dist) / spread))
    (a, b), _ = curve_fit(curve, xv, yv)
    return float(a), float(b)


def _spectral_init(rows, cols, vals, n: int, device, dim: int = 2):
    """umap-learn's spectral layout for graphs small enough for a dense eigen-decomposition: eigenvectors 1..dim of the
    symmetric normalised Laplacian (cuSOLVER through torch: plumbing on an n x n matrix, n <= 8192)."""
    import torch

    A = torch.zeros((n, n), dtype=torch.float32, device=device)
    A.index_put_((rows, cols), vals)
    deg = A.sum(dim=1).clamp_min(1e-12)
    dinv = deg.rsqrt()
    L = torch.eye(n, device=device) - dinv[:, None] * A * dinv[None, :]
    w, v = torch.linalg.eigh(L.double())
    return v[:, 1 : dim + 1].float().contiguous()


def umap_device(rows, cols, vals, n: int, *, init, n_epochs: int | None = None, min_dist: float = 0.5, spread: float = 1.0,
                gamma: float = 1.0, negative_sample_rate: int = 5, alpha: float = 1.0, seed: int = 0):
    """umap-learn's ``simplicial_set_embedding`` + ``optimize_layout_euclidean`` on a symmetric COO graph (device tensors,
    both directions present).  ``init``: device float32 ``[n, 2]``.  Returns the embedding (device float32 ``[n, 2]``)."""
    import torch

    lib = _lib.load()
    device = vals.device
    if n_epochs is None:
        n_epochs = 500 if n <= 10000 else 200
    a, b = find_ab_params(spread, min_dist)
    keep = vals >= vals.max() / float(n_epochs)  # edges that would be sampled less than once are dropped
    head = rows[keep].to(torch.int32).contiguous()
    tail = cols[keep].to(torch.int32).contiguous()
    w = vals[keep].float()
    eps = (w.max() / w).contiguous()
    next_sample = eps.clone()
    next_negative = (eps / float(negative_sample_rate)).contiguous()
    # 10 * (init - min) / (max - min) per coordinate, after umap-learn's scaling to max |x| = 10 plus 1e-4 noise
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    emb = init.float() * (10.0 / init.abs().max().clamp_min(1e-12)) + 1e-4 * torch.randn(init.shape, generator=g, device=device)
    lo, hi = emb.min(dim=0).values, emb.max(dim=0).values
    emb = (10.0 * (emb - lo) / (hi - lo).clamp_min(1e-12)).contiguous()
    _lib.check(
        lib.icnv_umap_epochs(_lib.ptr(head), _lib.ptr(tail), head.numel(), _lib.ptr(emb), n, _lib.ptr(eps), _lib.ptr(next_sample),
                             _lib.ptr(next_negative), a, b, gamma, alpha, n_epochs, 0, n_epochs, negative_sample_rate, seed & 0xFFFFFFFF,
                             _lib.stream_handle(device)),
        "icnv_umap_epochs",
    )
    return emb


def umap(
    adata,
    neighbors_key: str = "cnv_neighbors",
    key_added: str = "cnv_umap",
    inplace: bool = True,
    **kwargs,
):
    """Compute the UMAP on the CNV neighbourhood graph (GPU).

    Same parameters / keys as the reference (``tl/__init__.py:78-109``): reads the connectivities stored by
    ``cnv.pp.neighbors``, writes ``adata.obsm["X_" + key_added]`` or returns the ``[n_obs, 2]`` array.  Supported scanpy
    keywords: ``min_dist`` (0.5), ``spread`` (1.0), ``maxiter`` / ``n_epochs``, ``alpha`` (1.0), ``gamma`` (1.0),
    ``negative_sample_rate`` (5), ``random_state`` (0), ``init_pos`` ("spectral" | "pca" | "random" | array).  The spectral
    initialisation is used up to 8192 cells; above that the first two CNV principal components take its place.
    """
    import torch

    from ..pp._neighbors import _unsupported_kwargs

    opts = dict(min_dist=0.5, spread=1.0, n_epochs=None, alpha=1.0, gamma=1.0, negative_sample_rate=5, random_state=0, init_pos="spectral")
    if "maxiter" in kwargs:
        kwargs["n_epochs"] = kwargs.pop("maxiter")
    for k in list(kwargs):
        if k in opts:
            opts[k] = kwargs.pop(k)
    _unsupported_kwargs("umap", kwargs, {"n_components": (2,), "copy": (False,), "method": ("umap",), "a": (None,), "b": (None,)})
    if neighbors_key not in adata.uns:
        raise KeyError(f"No neighbors graph under {neighbors_key!r}. Did you run `pp.neighbors`?")
    if "shard" in adata.uns[neighbors_key]:
        raise NotImplementedError("tl.umap works on the whole graph: gather the row shards first")
    ckey = adata.uns[neighbors_key].get("connectivities_key", f"{neighbors_key}_connectivities")
    A = sp.coo_matrix(adata.obsp[ckey])
    n = A.shape[0]
    device = _device()
    rows = torch.from_numpy(A.row.astype(np.int64)).to(device)
    cols = torch.from_numpy(A.col.astype(np.int64)).to(device)
    vals = torch.from_numpy(A.data.astype(np.float32)).to(device)
    seed = int(opts["random_state"] or 0)
    init_pos = opts["init_pos"]
    if isinstance(init_pos, str) and init_pos == "spectral" and n > 8192:
        init_pos = "pca"
    if isinstance(init_pos, str):
        if init_pos == "spectral":
            init = _spectral_init(rows, cols, vals, n, device)
        elif init_pos == "pca" and "X_cnv_pca" in adata.obsm:
            init = torch.from_numpy(np.ascontiguousarray(np.asarray(adata.obsm["X_cnv_pca"])[:, :2], dtype=np.float32)).to(device)
        else:
            g = torch.Generator(device=device)
            g.manual_seed(seed)
            init = torch.rand((n, 2), generator=g, device=device) * 20.0 - 10.0
    else:
        init = torch.from_numpy(np.ascontiguousarray(np.asarray(init_pos), dtype=np.float32)).to(device)
        if tuple(init.shape) != (n, 2):
            raise ValueError("init_pos must have shape (n_obs, 2)")
    emb = umap_device(rows, cols, vals, n, init=init, n_epochs=opts["n_epochs"], min_dist=float(opts["min_dist"]),
                      spread=float(opts["spread"]), gamma=float(opts["gamma"]), negative_sample_rate=int(opts["negative_sample_rate"]),
                      alpha=float(opts["alpha"]), seed=seed)
    res = emb.cpu().numpy()
    if inplace:
        adata.obsm[f"X_{key_added}"] = res
    else:
        return res


def tsne_device(Y, *, perplexity: float = 30.0, early_exaggeration: float = 12.0, learning_rate: float = 1000.0, n_iter: int = 1000,
                seed: int = 0):
    """Exact t-SNE (scikit-learn ``TSNE(method="exact")``'s schedule: 250 exaggerated iterations at momentum 0.5, then
    momentum 0.8) of the rows of ``Y`` (device float32 ``[n, d <= 64]``), PCA-style initialisation from its first two
    columns scaled to a standard deviation of 1e-4.  Returns device float32 ``[n, 2]``."""
    import torch

    lib = _lib.load()
    n, d = Y.shape
    if n > 32768:
        raise ValueError("exact t-SNE holds the n x n affinities in HBM: at most 32768 cells (use tl.umap beyond)")
    device = Y.device
    Yc = Y.contiguous()
    P = torch.empty((n, n), dtype=torch.float32, device=device)
    st = _lib.stream_handle(device)
    _lib.check(lib.icnv_tsne_affinities(_lib.ptr(Yc), n, d, Yc.stride(0), float(perplexity), _lib.ptr(P), st), "icnv_tsne_affinities")
    emb = Yc[:, :2].clone() if d >= 2 else torch.zeros((n, 2), dtype=torch.float32, device=device)
    emb = emb - emb.mean(dim=0)
    emb = (emb / emb[:, 0].std().clamp_min(1e-12) * 1e-4).contiguous()
    vel = torch.zeros_like(emb)
    gains = torch.ones_like(emb)
    work = torch.empty((int(lib.icnv_tsne_work_floats(n)),), dtype=torch.float32, device=device)
    n_early = min(250, n_iter)
    for iters, exag, mom in ((n_early, float(early_exaggeration), 0.5), (n_iter - n_early, 1.0, 0.8)):
        if iters > 0:
            _lib.check(
                lib.icnv_tsne_iterations(_lib.ptr(P), _lib.ptr(emb), _lib.ptr(vel), _lib.ptr(gains), _lib.ptr(work), n, iters, exag, mom,
                                         float(learning_rate), st),
                "icnv_tsne_iterations",
            )
    return emb


def tsne(
    adata,
    use_rep: str = "cnv_pca",
    key_added: str = "cnv_tsne",
    inplace: bool = True,
    **kwargs,
):
    """Compute the t-SNE on the PCA of the CNV matrix (GPU, exact gradient).

    Same parameters / keys as the reference (``tl/__init__.py:112-144``; like it, the PCA is computed first when
    ``X_cnv_pca`` is missing).  Supported scanpy keywords: ``perplexity`` (30), ``early_exaggeration`` (12),
    ``learning_rate`` (1000), ``random_state`` (0), ``n_iter`` (1000).  Exact t-SNE: at most 32768 cells.
    """
    import torch

    from ..pp._neighbors import _unsupported_kwargs
    from ._pca import pca

    if f"X_{use_rep}" not in adata.obsm and use_rep == "cnv_pca":
        log.warning("X_cnv_pca not found in adata.obsm. Computing PCA with default parameters")
        pca(adata)
    opts = dict(perplexity=30.0, early_exaggeration=12.0, learning_rate=1000.0, random_state=0, n_iter=1000)
    for k in list(kwargs):
        if k in opts:
            opts[k] = kwargs.pop(k)
    _unsupported_kwargs("tsne", kwargs, {"n_pcs": (None,), "copy": (False,), "metric": ("euclidean",), "use_fast_tsne": (False,), "n_jobs": None})
    # the reference always embeds X_cnv_pca (tl/__init__.py:139), whatever use_rep says
    Y = np.ascontiguousarray(np.asarray(adata.obsm["X_cnv_pca"]), dtype=np.float32)
    if Y.shape[1] > 64:
        Y = Y[:, :64]
    n = Y.shape[0]
    perplexity = float(opts["perplexity"])
    if perplexity >= n:
        raise ValueError("perplexity must be less than n_samples")
    device = _device()
    emb = tsne_device(torch.from_numpy(Y).to(device), perplexity=perplexity, early_exaggeration=float(opts["early_exaggeration"]),
                      learning_rate=float(opts["learning_rate"]), n_iter=int(opts["n_iter"]), seed=int(opts["random_state"] or 0))
    res = emb.cpu().numpy()
    if inplace:
        adata.obsm[f"X_{key_added}"] = res
    else:
        return res
