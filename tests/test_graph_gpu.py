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
    # against the EXACT float64 SVD of the same matrix (ARPACK is itself approximate): singular values to float32
    # accuracy, the 20-dimensional score subspace to 1e-5 (largest principal angle), every component up to its spectral gap
    Xd = X.toarray().astype(np.float64)
    U, S, Vt = np.linalg.svd(Xd, full_matrices=False)
    g64 = got.astype(np.float64)
    np.testing.assert_allclose(np.linalg.norm(g64, axis=0), S[:20], rtol=5e-6)
    Qg, _ = np.linalg.qr(g64)
    cosines = np.linalg.svd(Qg.T @ U[:, :20], compute_uv=False)
    assert 1.0 - cosines.min() < 1e-9, cosines.min()
    for c in range(20):
        exact = U[:, c] * S[c]
        exact = exact * np.sign(Vt[c, np.abs(Vt[c]).argmax()])  # svd_flip convention: largest |loading| positive
        gap = min(S[c - 1] - S[c] if c else np.inf, S[c] - S[c + 1])
        err = np.linalg.norm(g64[:, c] - exact) / S[c]
        assert err < 3e-6 * max(1.0, S[c] / gap), (c, err, S[c] / gap)
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


@pytest.mark.parametrize("n,d,k", [(3000, 50, 15), (1000, 7, 15), (500, 50, 20), (40, 3, 2), (129, 64, 5)])
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


def test_knn_tensor_core_path_at_scale_and_sharded():
    """20 011 clustered points x 50 dims (not a multiple of the 128-point tile): the tcgen05 distance GEMM + exact re-rank
    against a float64 brute force on the device; then the same lists from two query shards, which
    is how ranks split the queries after the all-gather of the coordinates."""
    import torch
    from infercnvpy_b200.pp._neighbors import knn_device

    rng = np.random.default_rng(5)
    n, d, k = 20011, 50, 15
    centers = rng.normal(size=(12, d)) * 3.0
    P = (centers[rng.integers(0, 12, size=n)] + rng.normal(size=(n, d))).astype(np.float32)
    Pd = torch.from_numpy(P).cuda()
    idx, dist = knn_device(Pd, k)
    want_i = torch.empty((n, k), dtype=torch.int64, device="cuda")
    want_d = torch.empty((n, k), dtype=torch.float64, device="cuda")
    P64 = Pd.double()
    for a in range(0, n, 2048):
        D = torch.cdist(P64[a : a + 2048], P64).pow(2)
        v, i = torch.sort(D, dim=1, stable=True)
        want_i[a : a + 2048], want_d[a : a + 2048] = i[:, :k], v[:, :k]
    assert bool((idx[:, 0].long() == torch.arange(n, device="cuda")).all())
    np.testing.assert_allclose((dist.double() ** 2).cpu().numpy(), want_d.cpu().numpy(), rtol=2e-5, atol=1e-5)
    same = (idx.long() == want_i).float().mean().item()
    assert same > 0.9999, same  # identities can only differ between exact ties
    i0, d0 = knn_device(Pd, k, q0=0, nq=10000)  # shard boundaries are multiples of chunksize, not of the 128-point tile
    i1, d1 = knn_device(Pd, k, q0=10000, nq=n - 10000)
    assert torch.equal(torch.cat([i0, i1]), idx) and torch.equal(torch.cat([d0, d1]), dist)


def test_knn_candidate_super_blocks_carry_the_lists():
    """Beyond 1536 candidate tiles (196 608 points) the candidates are streamed in L2-sized super-blocks, one launch each,
    and a query tile's candidate lists travel between the launches: 270 001 points, 1500 queries in the middle, against a
    float64 brute force."""
    import torch
    from infercnvpy_b200.pp._neighbors import knn_device

    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    n, d, k, q0, nq = 270001, 50, 15, 133337, 1500
    P = torch.randn((n, d), generator=g, device="cuda") + 2.5 * torch.randn((40, d), generator=g, device="cuda")[torch.randint(0, 40, (n,), generator=g, device="cuda")]
    idx, dist = knn_device(P, k, q0=q0, nq=nq)
    P64 = P.double()
    want_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    want_d = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    for a in range(0, nq, 250):
        D = torch.cdist(P64[q0 + a : q0 + a + 250], P64).pow(2)
        v, i = torch.sort(D, dim=1, stable=True)
        want_i[a : a + 250], want_d[a : a + 250] = i[:, :k], v[:, :k]
    assert bool((idx[:, 0].long() == torch.arange(q0, q0 + nq, device="cuda")).all())
    np.testing.assert_allclose((dist.double() ** 2).cpu().numpy(), want_d.cpu().numpy(), rtol=2e-5, atol=1e-5)
    assert (idx.long() == want_i).float().mean().item() > 0.9999
    # the neighbours come from every super-block, not only from the last launch
    assert int(idx.min()) < 1024 * 128 < int(idx.max())


def test_community_sweep_handles_hubs_exactly():
    """A star-like graph whose hubs have thousands of neighbours (the sweep used to truncate adjacency lists at
    96 edges): quality of the GPU partition vs the CPU Leiden restatement, and no overflow / truncation."""
    import torch

    from infercnvpy_b200.tl._leiden import _csr_to_device, leiden_device
    from oracle import leiden_oracle as lo

    rng = np.random.default_rng(11)
    n, n_blocks = 4000, 4
    block = rng.integers(0, n_blocks, size=n)
    rows, cols = [], []
    for b in range(n_blocks):  # one hub per block connected to every member, plus sparse random edges inside the block
        members = np.flatnonzero(block == b)
        hub = members[0]
        rows += [hub] * (len(members) - 1)
        cols += members[1:].tolist()
        m = len(members) * 3
        rows += rng.choice(members, m).tolist()
        cols += rng.choice(members, m).tolist()
    rows += rng.integers(0, n, 300).tolist()  # a few edges between blocks
    cols += rng.integers(0, n, 300).tolist()
    A = sp.csr_matrix((rng.uniform(0.2, 1.0, len(rows)), (rows, cols)), shape=(n, n))
    A = A.maximum(A.T)
    A.setdiag(0)
    A.eliminate_zeros()
    assert np.diff(A.indptr).max() > 900
    dev = torch.device("cuda", 0)
    lab = leiden_device(*_csr_to_device(A, dev)).cpu().numpy()
    q_gpu, q_cpu = lo.quality(A, lab), lo.quality(A, lo.leiden(A))
    print(f"\n[hubs] quality GPU {q_gpu:.4f} vs CPU Leiden {q_cpu:.4f}; clusters {len(np.unique(lab))}")
    assert q_gpu >= q_cpu - 0.02


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
    # quality and agreement against (a) the sequential CPU restatement of Leiden (oracle/leiden_oracle.py: the "CPU
    # Leiden-equivalent"; leidenalg itself is not installed) and (b) networkx's Louvain, on the same graph
    from oracle import leiden_oracle as lo

    A = adata.obsp["cnv_neighbors_connectivities"].tocsr()
    dev = torch.device("cuda", 0)
    indptr, indices, w = _csr_to_device(A, dev)
    q_ours = modularity_device(indptr, indices, w, torch.from_numpy(codes.astype(np.int64)).to(dev))
    assert abs(q_ours - lo.quality(A, codes.astype(np.int64))) < 1e-6  # the two quality evaluations agree
    cpu = lo.leiden(A, gamma=1.0, seed=0)
    q_cpu = lo.quality(A, cpu)
    G = nx.from_scipy_sparse_array(A)
    comms = nx.community.louvain_communities(G, weight="weight", resolution=1.0, seed=0)
    nxlab = np.zeros(A.shape[0], dtype=np.int64)
    for k, cset in enumerate(comms):
        nxlab[list(cset)] = k
    q_nx = lo.quality(A, nxlab)
    ari_cpu, ari_nx = adjusted_rand_score(cpu, codes), adjusted_rand_score(nxlab, codes)
    print(f"\n[leiden] quality: GPU {q_ours:.4f}, CPU Leiden {q_cpu:.4f}, networkx Louvain {q_nx:.4f}; "
          f"ARI vs CPU Leiden {ari_cpu:.3f}, vs networkx {ari_nx:.3f}; clusters {len(sizes)} / {cpu.max() + 1} / {len(comms)}")
    assert q_ours >= q_cpu - 0.01 and q_ours >= q_nx - 0.01, (q_ours, q_cpu, q_nx)
    assert ari_cpu > 0.6 and ari_nx > 0.5
    # Leiden guarantee: every cluster induces a connected subgraph
    from scipy.sparse.csgraph import connected_components

    for c in np.unique(codes):
        members = np.flatnonzero(codes == c)
        ncomp, _ = connected_components(A[members][:, members], directed=False)
        assert ncomp == 1, f"cluster {c} is split into {ncomp} components"
    # determinism: a second run gives the same labels
    # inplace=False: an annotated COPY comes back (scanpy's copy=True, tl/__init__.py:28) and the input is untouched
    before = adata.obs["cnv_leiden"].copy()
    copy = cnv.tl.leiden(adata, inplace=False, key_added="again")
    assert "again" not in adata.obs.columns and "again" not in adata.uns and adata.obs["cnv_leiden"].equals(before)
    assert list(copy.obs["again"]) == list(lab) and copy.uns["again"]["params"]["resolution"] == 1.0
    # cnv_score on the clusters: the altered clones score higher than the normal one
    cnv.tl.cnv_score(adata)
    score = adata.obs.groupby("clone", observed=True)["cnv_score"].mean()
    assert score["k0"] < score["k1"] and score["k0"] < score["k2"]
    res = cnv.tl.leiden(adata, inplace=False, resolution=0.5).obs["cnv_leiden"]
    assert len(res.cat.categories) <= len(lab.cat.categories)


def _knn_purity(emb, labels, k=10):
    from sklearn.neighbors import NearestNeighbors

    idx = NearestNeighbors(n_neighbors=k + 1).fit(emb).kneighbors(emb, return_distance=False)[:, 1:]
    return float((labels[idx] == labels[:, None]).mean())


def test_umap_and_tsne_embeddings(clones):
    """/root/reference/src/infercnvpy/tl/__init__.py:78-144 (wrappers around scanpy.tl.umap / scanpy.tl.tsne; the
    reference's test_workflow asserts nothing).  Ours: keys / shapes like the reference, the local structure of the PCA
    space survives (scikit-learn trustworthiness) and the planted clones stay apart in the plane."""
    from sklearn.manifold import trustworthiness

    adata, clone = clones
    if "cnv_neighbors" not in adata.uns:
        cnv.tl.pca(adata)
        cnv.pp.neighbors(adata)
    Y = np.asarray(adata.obsm["X_cnv_pca"])
    cnv.tl.umap(adata)
    U = adata.obsm["X_cnv_umap"]
    assert U.shape == (adata.n_obs, 2) and U.dtype == np.float32 and np.isfinite(U).all()
    tw_u, pur_u = trustworthiness(Y, U, n_neighbors=15), _knn_purity(U, clone)
    # deterministic schedule and hash-based negative samples, float atomics: repeat runs agree up to summation order
    U2 = cnv.tl.umap(adata, inplace=False)
    assert U2.shape == U.shape and _knn_purity(U2, clone) > 0.95
    V = cnv.tl.umap(adata, inplace=False, init_pos="random", min_dist=0.1, maxiter=100)
    assert np.isfinite(V).all() and _knn_purity(V, clone) > 0.9
    with pytest.raises(TypeError):
        cnv.tl.umap(adata, n_components=3)
    cnv.tl.tsne(adata)
    T = adata.obsm["X_cnv_tsne"]
    assert T.shape == (adata.n_obs, 2) and np.isfinite(T).all()
    tw_t, pur_t = trustworthiness(Y, T, n_neighbors=15), _knn_purity(T, clone)
    print(f"\n[embeddings] umap trustworthiness {tw_u:.3f}, clone purity {pur_u:.3f}; tsne {tw_t:.3f}, {pur_t:.3f}")
    assert tw_u > 0.80 and pur_u > 0.95
    assert tw_t > 0.90 and pur_t > 0.95
    with pytest.raises(ValueError, match="perplexity"):
        cnv.tl.tsne(adata, perplexity=1e9)


def test_tsne_affinities_match_numpy_restatement():
    """Conditional probabilities of the requested perplexity, symmetrised (scikit-learn's _joint_probabilities)."""
    import torch

    from infercnvpy_b200 import _lib

    rng = np.random.default_rng(0)
    X = rng.normal(size=(300, 7)).astype(np.float32)
    dev = torch.device("cuda", 0)
    Xd = torch.from_numpy(X).to(dev)
    P = torch.empty((300, 300), dtype=torch.float32, device=dev)
    lib = _lib.load()
    _lib.check(lib.icnv_tsne_affinities(_lib.ptr(Xd), 300, 7, 7, 20.0, _lib.ptr(P), _lib.stream_handle(dev)), "icnv_tsne_affinities")
    P = P.cpu().numpy().astype(np.float64)
    assert np.allclose(P, P.T) and abs(P.sum() - 1.0) < 1e-4 and (np.diag(P) == 0).all()
    # numpy restatement: per-row bisection on beta to entropy log(perplexity)
    D = ((X[:, None, :].astype(np.float64) - X[None, :, :]) ** 2).sum(-1)
    C = np.zeros_like(D)
    for i in range(300):
        lo, hi, beta = -np.inf, np.inf, 1.0
        d = np.delete(D[i], i)
        for _ in range(200):
            p = np.exp(-d * beta)
            s = p.sum()
            H = np.log(s) + beta * (d * p).sum() / s
            if abs(H - np.log(20.0)) < 1e-7:
                break
            if H > np.log(20.0):
                lo, beta = beta, beta * 2 if np.isinf(hi) else (beta + hi) / 2
            else:
                hi, beta = beta, beta / 2 if np.isinf(lo) else (beta + lo) / 2
        C[i, np.arange(300) != i] = p / s
    want = np.maximum((C + C.T) / 600.0, 1e-12)
    np.fill_diagonal(want, 0.0)
    np.testing.assert_allclose(P, want, rtol=2e-3, atol=1e-9)


def test_reference_workflow_sequence_on_183_cells():
    """/root/reference/tests/test_tools.py:206-218 (test_workflow) call for call, on a stand-in of the size of the
    reference's bundled oligodendroglioma data set (183 cells): infercnv -> pca -> neighbors -> tsne -> umap -> leiden ->
    cnv_score, then the group means behind pl.chromosome_heatmap_summary.  The reference asserts nothing there; here every
    key it would leave behind exists with the reference's shape / dtype and the planted clones are found."""
    var = cnv.datasets.synthetic_var(4000, seed=3)
    X, clone = cnv.datasets.synthetic_counts_with_cnv(183, var, seed=11)
    adata = cnv.AnnData(X, var=var)
    cnv.tl.infercnv(adata)
    cnv.tl.pca(adata)
    cnv.pp.neighbors(adata)
    cnv.tl.tsne(adata)
    cnv.tl.umap(adata)
    cnv.tl.leiden(adata)
    cnv.tl.cnv_score(adata)
    n = adata.n_obs
    assert sp.issparse(adata.obsm["X_cnv"]) and adata.obsm["X_cnv"].dtype == np.float64 and "chr_pos" in adata.uns["cnv"]
    assert adata.obsm["X_cnv_pca"].shape == (n, 50) and adata.obsm["X_cnv_pca"].dtype == np.float32
    assert adata.obsp["cnv_neighbors_connectivities"].shape == (n, n) and adata.obsp["cnv_neighbors_distances"].shape == (n, n)
    assert adata.uns["cnv_neighbors"]["params"]["n_neighbors"] == 15
    assert adata.obsm["X_cnv_tsne"].shape == (n, 2) and adata.obsm["X_cnv_umap"].shape == (n, 2)
    assert np.isfinite(adata.obsm["X_cnv_tsne"]).all() and np.isfinite(adata.obsm["X_cnv_umap"]).all()
    assert str(adata.obs["cnv_leiden"].dtype) == "category" and adata.obs["cnv_score"].dtype == np.float64
    codes = adata.obs["cnv_leiden"].cat.codes.values
    purity = sum(np.bincount(clone[codes == c]).max() for c in np.unique(codes)) / n
    assert purity > 0.9, purity
    groups, means = cnv.pl.group_means(adata, "cnv_leiden")
    assert means.shape == (len(groups), adata.obsm["X_cnv"].shape[1])
    g0 = groups[0]
    want = np.asarray(adata.obsm["X_cnv"][np.asarray(adata.obs["cnv_leiden"] == g0)].mean(axis=0)).ravel()
    np.testing.assert_allclose(means[0], want, rtol=1e-12, atol=1e-15)
