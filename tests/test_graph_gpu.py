"""pca / neighbors / leiden / whole workflow on the GPU against scikit-learn / numpy restatements.
Parity for these steps is unpinned by the reference: tolerances below are ours."""

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

import infercnvpy_b200 as cnv
from oracle import graph_oracle as gor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clones():
    """3000 cells x 6000 genes with three planted clones, run through tl.infercnv."""
    var = cnv.datasets.synthetic_var(6000, seed=0)
    X, clone = cnv.datasets.synthetic_counts_with_cnv(3000, var, seed=7)
    obs = pd.DataFrame({"clone": [f"k{c}" for c in clone]}, index=[f"c{i}" for i in range(3000)])
    adata = cnv.AnnData(X, obs=obs, var=var)
    cnv.tl.infercnv(adata, reference_key="clone", reference_cat="k0", window_size=100, chunksize=1000)
    return adata, clone


def test_pca_matches_truncated_svd(clones):
    from sklearn.decomposition import TruncatedSVD

    adata, _ = clones
    X = adata.obsm["X_cnv"]
    got = cnv.tl.pca(adata, inplace=False, n_comps=20)
    assert got.dtype == np.float32 and got.shape == (3000, 20)
    want = TruncatedSVD(n_components=20, algorithm="arpack", random_state=0).fit_transform(X)
    sv = np.linalg.norm(want, axis=0)
    # leading, well separated components agree column by column (sign fixed by svd_flip)
    for c in range(8):
        err = np.linalg.norm(got[:, c] - want[:, c]) / sv[c]
        assert err < 5e-3, (c, err)
    # all singular values agree
    np.testing.assert_allclose(np.linalg.norm(got.astype(np.float64), axis=0), sv, rtol=2e-4)
    # keys / errors of the reference wrapper (tl/__init__.py:63-73)
    cnv.tl.pca(adata)
    assert adata.obsm["X_cnv_pca"].shape == (3000, 50)
    with pytest.raises(KeyError, match="X_nope"):
        cnv.tl.pca(adata, use_rep="nope")
    # zero-centred variant equals sklearn PCA scores up to sign
    from sklearn.decomposition import PCA

    got_c = cnv.tl.pca(adata, inplace=False, n_comps=5, zero_center=True)
    want_c = PCA(n_components=5, svd_solver="full").fit_transform(X.toarray())
    for c in range(3):
        e = min(np.linalg.norm(got_c[:, c] - want_c[:, c]), np.linalg.norm(got_c[:, c] + want_c[:, c])) / np.linalg.norm(want_c[:, c])
        assert e < 5e-3


@pytest.mark.parametrize("n,d,k", [(3000, 50, 15), (1000, 7, 15), (500, 50, 31), (40, 3, 2)])
def test_knn_is_exact_on_random_points(n, d, k):
    import torch
    from infercnvpy_b200.pp._neighbors import knn_device

    P = np.random.default_rng(n + d).normal(size=(n, d)).astype(np.float32)
    idx, dist = knn_device(torch.from_numpy(P).cuda(), k)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    D = ((P[:, None, :].astype(np.float64) - P[None, :, :].astype(np.float64)) ** 2).sum(-1)
    want = np.argsort(D, axis=1, kind="stable")[:, :k]
    assert (idx[:, 0] == np.arange(n)).all()
    assert all(set(idx[i]) == set(want[i]) for i in range(n))
    np.testing.assert_allclose(dist**2, np.sort(D, axis=1)[:, :k], rtol=1e-5, atol=1e-6)


def test_knn_exact_and_fuzzy_graph(clones):
    from sklearn.neighbors import NearestNeighbors

    adata, _ = clones
    if "X_cnv_pca" not in adata.obsm:
        cnv.tl.pca(adata)
    P = adata.obsm["X_cnv_pca"]
    dist, conn = cnv.pp.neighbors(adata, inplace=False)
    n = P.shape[0]
    assert dist.shape == (n, n) and conn.shape == (n, n)
    # float64 brute force (scikit-learn's float32 brute force expands |x|^2 + |y|^2 - 2xy and mis-orders near-ties)
    nn = NearestNeighbors(n_neighbors=15, algorithm="brute").fit(P.astype(np.float64))
    wd, wi = nn.kneighbors(P.astype(np.float64))
    # same neighbour distances row by row (cells whose CNV profile is entirely below the noise filter have
    # identical PCA coordinates, so neighbour *identities* are only defined up to ties)
    got_d = np.sort(np.asarray(dist.todense()), axis=1)[:, -14:]
    nz_rows = wd[:, 1] > 0
    np.testing.assert_allclose(got_d[nz_rows], wd[nz_rows, 1:], rtol=2e-4, atol=1e-5)
    # connectivities: symmetric, in (0, 1], equal to the umap-learn restatement on the same kNN lists
    assert abs(conn - conn.T).max() < 1e-6
    assert conn.data.min() > 0 and conn.data.max() <= 1.0 + 1e-6
    import torch
    from infercnvpy_b200.pp._neighbors import knn_device

    ki, kd = knn_device(torch.from_numpy(P).cuda(), 15)
    want = gor.fuzzy_simplicial_set(ki.cpu().numpy().astype(np.int64), kd.cpu().numpy())
    diff = abs(conn - want)
    assert diff.max() < 5e-4, diff.max()
    assert (conn != 0).sum() == (want != 0).sum()
    # inplace keys (pp/__init__.py:43 via scanpy)
    cnv.pp.neighbors(adata)
    assert {"cnv_neighbors_distances", "cnv_neighbors_connectivities"} <= set(adata.obsp)
    assert adata.uns["cnv_neighbors"]["params"]["n_neighbors"] == 15


def test_leiden_and_workflow(clones):
    """/root/reference/tests/test_tools.py:206-218 (test_workflow: no assertions there) + quality checks of ours."""
    import networkx as nx
    import torch
    from sklearn.metrics import adjusted_rand_score

    from infercnvpy_b200.tl._leiden import _csr_to_device, modularity_device

    adata, clone = clones
    if "cnv_neighbors" not in adata.uns:
        cnv.tl.pca(adata)
        cnv.pp.neighbors(adata)
    with pytest.raises(ValueError, match="cnv_leiden"):
        cnv.tl.cnv_score(adata)
    cnv.tl.leiden(adata)
    lab = adata.obs["cnv_leiden"]
    assert str(lab.dtype) == "category" and list(lab.cat.categories) == [str(i) for i in range(len(lab.cat.categories))]
    sizes = lab.value_counts()[list(lab.cat.categories)].values
    assert all(sizes[i] >= sizes[i + 1] for i in range(len(sizes) - 1))  # "0" is the largest cluster
    # the planted clones are recovered: every cluster is (almost) pure
    codes = lab.cat.codes.values
    purity = sum(np.bincount(clone[codes == c]).max() for c in np.unique(codes)) / len(codes)
    assert purity > 0.97
    # quality: RB-configuration modularity not worse than networkx's Louvain on the same graph
    A = adata.obsp["cnv_neighbors_connectivities"].tocsr()
    dev = torch.device("cuda", 0)
    indptr, indices, w = _csr_to_device(A, dev)
    q_ours = modularity_device(indptr, indices, w, torch.from_numpy(codes.astype(np.int64)).to(dev))
    G = nx.from_scipy_sparse_array(A)
    comms = nx.community.louvain_communities(G, weight="weight", resolution=1.0, seed=0)
    nxlab = np.zeros(A.shape[0], dtype=np.int64)
    for k, cset in enumerate(comms):
        nxlab[list(cset)] = k
    q_nx = modularity_device(indptr, indices, w, torch.from_numpy(nxlab).to(dev))
    assert q_ours >= q_nx - 0.02, (q_ours, q_nx)
    assert adjusted_rand_score(nxlab, codes) > 0.5
    # cnv_score on the clusters: the altered clones score higher than the normal one
    cnv.tl.cnv_score(adata)
    score = adata.obs.groupby("clone", observed=True)["cnv_score"].mean()
    assert score["k0"] < score["k1"] and score["k0"] < score["k2"]
    res = cnv.tl.leiden(adata, inplace=False, resolution=0.5)
    assert len(res.categories) <= len(lab.cat.categories)
