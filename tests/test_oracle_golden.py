"""Pin the CPU oracle: reference KATs + outputs of the real reference (CPU only)."""

import hashlib

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import infercnv_oracle as orc
from oracle import ref_loader
from tests.golden.cases import CASES, build_case


# ---- known-answer tests ported from the reference's own suite ---------------------------------
def _mock_matrix(container):
    # /root/reference/tests/conftest.py:42-58
    return container(np.array([[1, 1, 1, 2], [2, 1, 2, 2], [5, 5, 5, 5], [7, 5, 5, 7], [9, 9, 9, 9]]))


MOCK_CAT = np.array(["foo", "foo", "bar", "baz", "bar"])


@pytest.mark.parametrize("container", [np.array, sp.csr_matrix, sp.csc_matrix])
def test_reference_profile_key_and_cat(container):
    # /root/reference/tests/test_tools.py:11-22
    got = orc.reference_profile(_mock_matrix(container), MOCK_CAT, ["foo", "baz"])
    np.testing.assert_almost_equal(np.asarray(got), np.array([[1.5, 1, 1.5, 2], [7, 5, 5, 7]]))


@pytest.mark.parametrize("container", [np.array, sp.csr_matrix, sp.csc_matrix])
def test_reference_profile_all_cells(container):
    # /root/reference/tests/test_tools.py:25-28
    got = orc.reference_profile(_mock_matrix(container))
    np.testing.assert_almost_equal(np.asarray(got), np.array([[4.8, 4.2, 4.4, 5]]), decimal=5)


def test_reference_profile_explicit_and_errors():
    # /root/reference/tests/test_tools.py:31-39 and _infercnv.py:393-398
    X = _mock_matrix(np.array)
    got = orc.reference_profile(X, MOCK_CAT, "bar", np.array([1, 2, 3, 4]))
    np.testing.assert_equal(got[0], np.array([1, 2, 3, 4]))
    with pytest.raises(ValueError):
        orc.reference_profile(X, MOCK_CAT, "bar", np.array([1, 2, 3]))
    with pytest.raises(ValueError):
        orc.reference_profile(X, MOCK_CAT, ["foo", "nope"])


def test_running_mean_window_smaller_than_segment():
    # /root/reference/tests/test_tools.py:64-89
    x = np.array([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]])
    np.testing.assert_array_equal(orc.running_mean(x, 3, 1), np.array([[2, 3, 4], [7, 8, 9]]))


def test_running_mean_window_larger_than_segment():
    # /root/reference/tests/test_tools.py:92-117 (and n == width takes the same branch)
    x = np.array([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]])
    np.testing.assert_array_equal(orc.running_mean(x, 7, 1), np.array([[3], [8]]))
    np.testing.assert_array_equal(orc.running_mean(x, 5, 1), np.array([[3], [8]]))


def test_pyramid_weights():
    assert orc.pyramid_weights(5).tolist() == [1, 2, 3, 2, 1]
    assert orc.pyramid_weights(100).sum() == 2550
    assert orc.pyramid_weights(250).sum() == 15750


def test_natural_order():
    got = orc.natural_chromosome_order(["chr10", "chr2", "chrM", "chr1", "GL1", None, "chrX", "chr2"])
    assert got == ["chr1", "chr2", "chr10", "chrX"]


def test_chunk_known_answer(full_mock, x_res_actual):
    # /root/reference/tests/test_tools.py:143-169
    X, var = full_mock
    Xs = sp.csr_matrix(X)
    ref = orc.reference_profile(Xs)
    chr_pos, res = orc.infercnv_chunk(Xs, var["chromosome"].values, var["start"].values, ref, 1, 3, 1, 1)
    np.testing.assert_array_equal(res.toarray(), x_res_actual)
    assert chr_pos == {"chr1": 0, "chr2": 3}


def test_public_multi_chunk_known_answer(full_mock, x_res_actual):
    # /root/reference/tests/test_tools.py:172-191
    X, var = full_mock
    chr_pos, res = orc.infercnv(
        sp.csr_matrix(X),
        var["chromosome"].values,
        var["start"].values,
        chunksize=2,
        lfc_clip=1,
        window_size=3,
        step=1,
        dynamic_threshold=1,
        n_jobs=2,
    )
    np.testing.assert_array_equal(res.toarray(), x_res_actual)
    assert chr_pos == {"chr1": 0, "chr2": 3}


def test_cnv_score_known_answer():
    # /root/reference/tests/test_scores.py:18-21 with conftest.py:111-139
    X_cnv = np.array([[1, 1, 1, 2, 2, 1, 1, 1], [2, 2, 2, 1, 1, 2, 2, 2], [4, 4, 4, 2, 2, 3, 3, 3], [2, 2, 2, 4, 4, 4, 4, 4]]).T
    groups = np.array(list("AAAAABBB"))
    for container in (np.array, sp.csr_matrix, sp.csc_matrix):
        res = orc.cnv_score(container(X_cnv), groups)
        assert res["A"] == pytest.approx(2.25, abs=1e-3)
        assert res["B"] == pytest.approx(2.5, abs=1e-3)


# ---- per-gene layer (calculate_gene_values=True): the reference's own known answers ------------------
def test_gene_values_window_smaller_than_segment():
    # /root/reference/tests/test_tools.py:64-89
    x = np.array([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]])
    pos, vals = orc.gene_values_for_segment(orc.running_mean(x, 3, 1), 5, 3, 1)
    np.testing.assert_array_equal(pos, np.arange(5))
    np.testing.assert_array_equal(vals, np.array([[2.0, 2.5, 3.0, 3.5, 4.0], [7.0, 7.5, 8.0, 8.5, 9.0]]))


def test_gene_values_window_larger_than_segment():
    # /root/reference/tests/test_tools.py:92-117
    x = np.array([[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]])
    pos, vals = orc.gene_values_for_segment(orc.running_mean(x, 7, 1), 5, 7, 1)
    np.testing.assert_array_equal(pos, np.arange(5))
    np.testing.assert_array_equal(vals, np.array([[3.0] * 5, [8.0] * 5]))


def test_gene_averages_known_answer():
    # /root/reference/tests/test_tools.py:120-140: three windows of three genes over five genes
    smoothed = np.array([[2, 3, 4], [4, 4, 6], [6, 2, 1]], dtype=np.float64)
    pos, vals = orc.gene_values_for_segment(smoothed, 5, 3, 1)
    want = np.array([[2.0, 2.5, 3.0, 3.5, 4.0], [4.0, 4.0, 14 / 3, 5.0, 6.0], [6.0, 4.0, 3.0, 1.5, 1.0]])
    np.testing.assert_allclose(vals, want, rtol=1e-15)


def test_chunk_with_gene_values_known_answer(full_mock, x_res_actual, gene_res_actual):
    # /root/reference/tests/test_tools.py:143-155
    X, var = full_mock
    ref = orc.reference_profile(X)
    chr_pos, res, (cols, gene_res) = orc.infercnv_chunk(
        X, var["chromosome"].values, var["start"].values, ref, 1, 3, 1, 1, gene_values=True
    )
    np.testing.assert_array_equal(res.toarray(), x_res_actual)
    np.testing.assert_array_equal(cols, np.arange(10))
    np.testing.assert_allclose(gene_res, gene_res_actual, rtol=1e-8)
    assert {k: int(v) for k, v in chr_pos.items()} == {"chr1": 0, "chr2": 3}


def test_public_with_gene_values_known_answer(full_mock, x_res_actual):
    # /root/reference/tests/test_tools.py:172-191 (two chunks of two cells)
    X, var = full_mock
    chr_pos, res, per_gene = orc.infercnv(
        X, var["chromosome"].values, var["start"].values, chunksize=2, lfc_clip=1, window_size=3, step=1,
        dynamic_threshold=1, calculate_gene_values=True,
    )
    np.testing.assert_array_equal(per_gene[0], np.array([0.75, 0.0, 0.0, 0.0, -0.75, 0.0, 0.0, 0.0, 0.0, 0.75]))
    np.testing.assert_array_equal(per_gene[3], np.array([0, 0, 0, 0, 0, 0.921875, 0.703125, 0, 0, 0]))
    np.testing.assert_array_equal(res.toarray(), x_res_actual)


# ---- outputs of the real reference, generated by tests/golden/make_golden.py ----------------------
def _run_oracle(case):
    X, var, obs, kw = build_case(case)
    kw = dict(kw)
    key = kw.pop("reference_key", None)
    Xin = sp.csr_matrix(X) if case.get("container") == "csr" else X
    out = orc.infercnv(
        Xin,
        var["chromosome"].values,
        var["start"].values,
        obs_column=None if key is None else obs[key].values,
        calculate_gene_values=bool(case.get("gene_values")),
        **kw,
    )
    return (X, *out) if case.get("gene_values") else (X, *out, None)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case, golden_loader):
    gold = golden_loader(case["name"])
    X, chr_pos, res, per_gene = _run_oracle(case)
    assert hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() == str(gold["x_sha256"]), "RNG stream drifted"
    assert {k: int(v) for k, v in chr_pos.items()} == gold["chr_pos"]
    assert list(chr_pos.keys()) == list(gold["chr_pos"].keys())
    assert res.shape == gold["csr"].shape
    # same numpy calls in the same order -> bit-identical float64
    np.testing.assert_array_equal(res.indptr, gold["csr"].indptr)
    np.testing.assert_array_equal(res.indices, gold["csr"].indices)
    np.testing.assert_array_equal(res.data, gold["csr"].data)
    assert res.dtype == np.float64
    if case.get("gene_values"):
        want = gold["per_gene"]
        assert per_gene.shape == want.shape
        np.testing.assert_array_equal(np.isnan(per_gene), np.isnan(want))
        np.testing.assert_array_equal(np.nan_to_num(per_gene), np.nan_to_num(want))  # bit-identical float64


def test_oracle_cnv_score_golden(golden_loader):
    gold = golden_loader("small_default")
    got = orc.cnv_score(gold["csr"], gold["score_labels"])
    want = dict(zip(gold["score_keys"].tolist(), gold["score_vals"].tolist()))
    assert set(got) == set(want)
    for k in want:
        assert float(got[k]) == want[k]


def test_csr_and_dense_goldens_identical(golden_loader):
    # dense and CSR containers differ only through the reference profile: numpy's float32
    # column mean vs scipy.sparse's ``.mean(axis=0)`` round differently (~1e-7), everything
    # after it is the same arithmetic
    ga, gb = golden_loader("small_default"), golden_loader("small_csr")
    assert np.abs(ga["profile"] - gb["profile"]).max() < 5e-7
    a, b = ga["csr"].toarray(), gb["csr"].toarray()
    assert ((a == 0) != (b == 0)).sum() == 0
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-7)


# ---- live cross-check against the reference when it is mounted (build container only) ----------
@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_live_reference_random(seed):
    ref_inf, _ = ref_loader.load()
    rng = np.random.default_rng(seed)
    n, g = int(rng.integers(5, 40)), int(rng.integers(200, 900))
    from infercnvpy_b200.datasets import synthetic_counts, synthetic_var

    var = synthetic_var(g, seed=seed + 10, with_extras=bool(seed % 2))
    X = synthetic_counts(n, g, seed=seed)
    window = int(rng.integers(3, 60))
    step = int(rng.integers(1, 12))
    chunksize = int(rng.integers(3, 20))
    adata = ref_loader.MiniAnnData(X, var=var)
    chr_pos_r, res_r, _ = ref_inf.infercnv(
        adata, window_size=window, step=step, chunksize=chunksize, inplace=False, n_jobs=1
    )
    chr_pos_o, res_o = orc.infercnv(
        X, var["chromosome"].values, var["start"].values, window_size=window, step=step, chunksize=chunksize
    )
    assert {k: int(v) for k, v in chr_pos_r.items()} == {k: int(v) for k, v in chr_pos_o.items()}
    np.testing.assert_array_equal(sp.csr_matrix(res_r).toarray(), res_o.toarray())


# ---- ITH scores (tl/_scores.py:77-221) --------------------------------------------------------------
ITH_EXPR = np.array([[1, 1, 1, 1, 1, 1, 2, 3], [2, 2, 2, 2, 2, 2, 8, 0], [3, 3, 3, 3, 3, 10, 3, 7]]).T
ITH_CNV = np.array([[1, 1, 1, 2, 2, 1, 1, 1], [2, 2, 2, 1, 1, 2, 2, 2], [4, 4, 4, 2, 2, 3, 3, 3], [2, 2, 2, 4, 4, 4, 4, 4]]).T
ITH_GROUPS = list("AAAAABBB")


@pytest.mark.parametrize("container", [np.array, sp.csr_matrix, sp.csc_matrix])
def test_ith_scores_reference_known_answers(container):
    # /root/reference/tests/test_scores.py:6-15 on /root/reference/tests/conftest.py:111-139
    gex = orc.ith_score(container(ITH_EXPR), ITH_GROUPS)
    assert gex["A"] == 0 and gex["B"] == pytest.approx(1.2628, abs=0.001)
    cna = orc.ith_score(container(ITH_CNV), ITH_GROUPS)
    assert cna["A"] == pytest.approx(1.053, abs=0.001) and cna["B"] == 0


def test_ith_scores_match_reference_golden():
    from tests.golden.make_golden_ith import ith_case

    from tests.conftest import GOLDEN

    z = np.load(GOLDEN / "ith_scores.npz", allow_pickle=False)
    expr, x_cnv, labels = ith_case()
    gex, cna = orc.ith_score(expr, labels), orc.ith_score(x_cnv, labels)
    assert sorted(gex) == z["keys"].tolist() and "solo" not in gex
    for k, g, c in zip(z["keys"].tolist(), z["ithgex"].tolist(), z["ithcna"].tolist()):
        assert float(gex[k]) == g and float(cna[k]) == c  # same numpy calls: bit-identical
