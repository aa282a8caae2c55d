"""Parity of the CUDA path (through the C ABI / public API) with the oracle and the goldens.

Tolerances.  The kernels accumulate every window in float64 and round once to float32, so
against the reference's float64 result the float32 output is within 1 ulp(float32) ~ 6e-8
relative; ``north_star`` asks for 1e-5.  We assert 1e-6 relative on every entry both sides keep,
and with float64 output 1e-11.  Entries the noise filter keeps on one side and zeroes on the other
("flips") are only legal if the value sits within float32 rounding of the chunk threshold.
"""

from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

import infercnvpy_b200 as cnv
from oracle import infercnv_oracle as orc
from tests.golden.cases import CASES, build_case

pytestmark = pytest.mark.gpu

RTOL32 = 1e-6


def _torch():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _adata(X, var, obs=None):
    return cnv.AnnData(X, obs=obs, var=var)


def _compare_thresholded(got: np.ndarray, want: np.ndarray, chunk: int, rtol=RTOL32, what="", max_flips=None):
    """Both dense float64 [n, K].  Values both sides keep must agree to ``rtol``; an entry kept on one
    side only ("flip") is legal only if it sits on the chunk's noise threshold, i.e. its magnitude is
    within 1e-5 relative of the smallest magnitude the reference keeps in that chunk.  Returns #flips."""
    assert got.shape == want.shape
    both = (got != 0) & (want != 0)
    np.testing.assert_allclose(got[both], want[both], rtol=rtol, atol=0, err_msg=what)
    n_flips = 0
    for r0 in range(0, got.shape[0], chunk):
        g, w = got[r0 : r0 + chunk], want[r0 : r0 + chunk]
        flips = (g != 0) != (w != 0)
        if not flips.any():
            continue
        n_flips += int(flips.sum())
        kept_min = np.abs(w[w != 0]).min() if (w != 0).any() else 0.0
        val = np.where(g[flips] != 0, np.abs(g[flips]), np.abs(w[flips]))
        assert np.all(val <= kept_min * (1 + 1e-5)), f"{what}: flip away from the threshold in chunk at row {r0}"
    budget = max(1, int(1e-6 * got.size)) if max_flips is None else max_flips
    assert n_flips <= budget, f"{what}: {n_flips} flips (budget {budget})"
    return n_flips


# ------------------------------------------------------------------------------------------------
GPU_CASES = [c for c in CASES if c.get("gpu", True)]


@pytest.mark.parametrize("case", GPU_CASES, ids=[c["name"] for c in GPU_CASES])
def test_public_api_matches_reference_golden(case, golden_loader):
    """cnv.tl.infercnv with the reference's own profile == the real reference's output."""
    gold = golden_loader(case["name"])
    X, var, obs, kw = build_case(case)
    kw = {k: v for k, v in kw.items() if k not in ("reference_key", "reference_cat", "reference")}
    Xin = sp.csr_matrix(X) if case.get("container") == "csr" else X
    adata = _adata(Xin, var, obs)
    chr_pos, res, per_gene = cnv.tl.infercnv(adata, reference=gold["profile"], inplace=False, **kw)
    assert per_gene is None
    assert {k: int(v) for k, v in chr_pos.items()} == gold["chr_pos"]
    assert list(chr_pos) == list(gold["chr_pos"])
    assert sp.issparse(res) and res.dtype == np.float64 and res.shape == gold["csr"].shape
    _compare_thresholded(res.toarray(), gold["csr"].toarray(), kw.get("chunksize", 5000), what=case["name"])


GV_CASES = [c for c in CASES if c.get("gene_values")]


def _compare_gene_layer(got, want, chunk, what):
    assert got.shape == want.shape and got.dtype == np.float64
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg=what)  # same genes covered (:146)
    g, w = np.nan_to_num(got), np.nan_to_num(want)
    both = (g != 0) & (w != 0)
    # float64 end to end: window sums differ from np.convolve only in summation order
    np.testing.assert_allclose(g[both], w[both], rtol=1e-9, atol=1e-13, err_msg=what)
    return _compare_thresholded(g, w, chunk, rtol=1e-6, what=what)


@pytest.mark.parametrize("case", GV_CASES, ids=[c["name"] for c in GV_CASES])
def test_gene_values_match_reference_golden(case, golden_loader):
    """calculate_gene_values=True (_infercnv.py:141-151, :247-291, :443-444, :452-453) == the real reference's layer."""
    gold = golden_loader(case["name"])
    X, var, obs, kw = build_case(case)
    kw = {k: v for k, v in kw.items() if k not in ("reference_key", "reference_cat", "reference")}
    adata = _adata(X, var, obs)
    chr_pos, res, per_gene = cnv.tl.infercnv(
        adata, reference=gold["profile"], inplace=False, calculate_gene_values=True, **kw
    )
    assert {k: int(v) for k, v in chr_pos.items()} == gold["chr_pos"]
    chunk = kw.get("chunksize", 5000)
    _compare_thresholded(res.toarray(), gold["csr"].toarray(), chunk, what=case["name"])
    _compare_gene_layer(per_gene, gold["per_gene"], chunk, case["name"])
    # inplace key (:157-158)
    cnv.tl.infercnv(adata, reference=gold["profile"], calculate_gene_values=True, key_added="k2", **kw)
    np.testing.assert_array_equal(np.nan_to_num(adata.layers["gene_values_k2"]), np.nan_to_num(per_gene))


def test_gene_values_reference_known_answers(full_mock, x_res_actual, gene_res_actual):
    """/root/reference/tests/test_tools.py:143-191 with the per-gene rows (conftest.py:78-95)."""
    X, var = full_mock
    adata = _adata(sp.csr_matrix(X), var)
    chr_pos, res, per_gene = cnv.tl.infercnv(
        adata, chunksize=2, lfc_clip=1, window_size=3, step=1, dynamic_threshold=1, inplace=False,
        calculate_gene_values=True,
    )
    np.testing.assert_allclose(res.toarray(), x_res_actual, rtol=1e-6, atol=0)
    # test_tools.py:185-191 pins rows 0 and 3 of the two-chunk run
    np.testing.assert_allclose(per_gene[0], [0.75, 0.0, 0.0, 0.0, -0.75, 0.0, 0.0, 0.0, 0.0, 0.75], rtol=1e-12)
    np.testing.assert_allclose(per_gene[3], [0, 0, 0, 0, 0, 0.921875, 0.703125, 0, 0, 0], rtol=1e-12)
    # one chunk of four cells == conftest's gene_res_actual (test_tools.py:143-169)
    _, _, per_gene1 = cnv.tl.infercnv(
        adata, chunksize=5000, lfc_clip=1, window_size=3, step=1, dynamic_threshold=1, inplace=False,
        calculate_gene_values=True,
    )
    np.testing.assert_allclose(per_gene1, gene_res_actual, rtol=1e-8)


@pytest.mark.parametrize("g,window,step", [(30000, 100, 10), (20000, 250, 10), (3000, 20, 3), (6000, 260, 1)])
def test_gene_values_match_oracle_other_shapes(g, window, step):
    """30000 genes: the per-gene means no longer fit in shared memory (L2 scratch path); window 260 / step 1: every
    gene sits in up to 260 windows, i.e. numpy's pairwise recursion beyond one block of 128."""
    n = 6
    var = cnv.datasets.synthetic_var(g, seed=5, with_extras=True)
    X = cnv.datasets.synthetic_counts(n, g, seed=g + window + step)
    ref = X.mean(axis=0, dtype=np.float64).astype(np.float32)
    adata = _adata(X, var)
    chr_pos, res, per_gene = cnv.tl.infercnv(
        adata, reference=ref, window_size=window, step=step, chunksize=4, inplace=False, calculate_gene_values=True
    )
    _, want_res, want_gene = orc.infercnv(
        X, var["chromosome"].values, var["start"].values, reference=ref, window_size=window, step=step, chunksize=4,
        calculate_gene_values=True,
    )
    _compare_thresholded(res.toarray(), want_res.toarray(), 4, what="windows")
    _compare_gene_layer(per_gene, want_gene, 4, f"g={g} window={window} step={step}")


def test_gene_values_properties_at_scale():
    """Bench-shaped gene axis, a few thousand cells: size-independent properties of the layer."""
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    G, N, chunk = 20000, 3000, 1000
    var = cnv.datasets.synthetic_var(G, seed=0)
    Xd = cnv.datasets.device_counts(N, G, dev, seed=77)
    layout = build_layout(var, 100, 10)
    with DevicePlan(layout, dev) as plan:
        sums, counts = plan.colsum(Xd)
        plan.set_reference(plan.mean_from_sums(sums, counts))
        tmp = plan.smooth(Xd, 3.0)
        out, stats = plan.center(tmp)
        thr, _, _ = plan.threshold(out, stats, chunk, 1.5)
        raw = plan.gene_values(tmp, chunk, None)
        filt = plan.gene_values(tmp, chunk, thr)
        n_cov = plan.n_covered
        # (1) the same columns are NaN in every row, and exactly n_covered are not
        nan = torch.isnan(raw)
        assert torch.equal(nan, nan[0:1].expand_as(nan)) and int((~nan[0]).sum()) == n_cov
        assert torch.equal(torch.isnan(filt), nan)
        # (2) every row of the unfiltered layer has median 0 over its covered genes (np.median, even n: middle pair)
        vals = raw[:, ~nan[0]]
        srt = vals.sort(dim=1).values
        mid = 0.5 * (srt[:, (n_cov - 1) // 2] + srt[:, n_cov // 2])
        assert float(mid.abs().max()) < 1e-15
        # (3) the filter zeroes exactly the entries below the row's chunk threshold and keeps the others unchanged
        t_row = thr.repeat_interleave(chunk)[:N, None]
        keep = vals.abs() >= t_row
        assert torch.equal(filt[:, ~nan[0]], torch.where(keep, vals, torch.zeros_like(vals)))
        # (4) a gene covered by exactly one window equals that window's pre-median value up to the two medians:
        #     value - window is constant along the row for all such genes (first `step` genes of every chromosome)
        pre = plan.center(tmp, out_dtype=torch.float64)[0]
        first_gene = torch.from_numpy(layout.gene_idx[layout.seg_off[:-1]].astype(np.int64)).to(dev)
        first_col = torch.from_numpy(np.asarray(layout.out_off[:-1], dtype=np.int64)).to(dev)
        d = raw[:, first_gene] - pre[:, first_col]
        assert float((d - d[:, :1]).abs().max()) < 1e-14


@pytest.mark.parametrize("name", ["small_default", "small_cats", "small_onecat", "small_csr", "g20k_win100"])
def test_data_derived_reference_profile(name, golden_loader):
    """Reference profile computed on the device (all cells / per category, dense and CSR)."""
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout
    from infercnvpy_b200.tl._infercnv import _reference_categories, _rows_to_device

    case = next(c for c in CASES if c["name"] == name)
    gold = golden_loader(name)
    X, var, obs, kw = build_case(case)
    Xin = sp.csr_matrix(X) if case.get("container") == "csr" else X
    adata = _adata(Xin, var, obs)
    layout = build_layout(var, kw.get("window_size", 100), kw.get("step", 10))
    dev = torch.device("cuda", 0)
    with DevicePlan(layout, dev) as plan:
        row_cat, n_cat = None, 1
        if "reference_key" in kw:
            rc, cats = _reference_categories(adata, kw["reference_key"], kw["reference_cat"])
            row_cat, n_cat = torch.from_numpy(rc).to(dev), len(cats)
        sums, counts = plan.colsum(_rows_to_device(Xin, 0, X.shape[0], dev), row_cat, n_cat)
        ref = plan.mean_from_sums(sums, counts).cpu().numpy()
    # numpy's float32 running mean vs our float64-accumulated mean: a few float32 ulps
    np.testing.assert_allclose(ref, gold["profile"], rtol=2e-6, atol=1e-7)

    # and the whole public path with the data-derived profile: north_star's 1e-5 relative on every entry both sides keep,
    # flips only ON the chunk threshold and at most a handful (measured on the CPU with the same float64-accumulated
    # profile: max rel 5e-7, 0 flips on these cases)
    chr_pos, res, _ = cnv.tl.infercnv(adata, inplace=False, **dict(kw))
    got, want = res.toarray(), gold["csr"].toarray()
    both = (got != 0) & (want != 0)
    rel = np.abs(got[both] - want[both]) / np.abs(want[both])
    n_flips = _compare_thresholded(got, want, kw.get("chunksize", 5000), rtol=1e-5, what=name, max_flips=2)
    print(f"\n[{name}] data-derived reference: max rel err {rel.max():.3e}, flips {n_flips} of {got.size}")


@pytest.mark.parametrize("window", [100, 250])
def test_bench_chunk_shape_default_reference(window):
    """The configuration bench.py times: one full chunk of 5000 cells x 20 000 genes, reference = mean of all cells
    (_infercnv.py:380-385), chunksize 5000, window 100 and 250, against the oracle (numpy float32 running mean).
    north_star tolerance: 1e-5 relative; flips only on the threshold, explicit budget."""
    G, N = 20000, 5000
    var = cnv.datasets.synthetic_var(G, seed=0)
    X = cnv.datasets.synthetic_counts(N, G, seed=4242)
    adata = _adata(X, var)
    chr_pos, res, _ = cnv.tl.infercnv(adata, window_size=window, chunksize=5000, inplace=False)
    chr_pos_o, want = orc.infercnv(X, var["chromosome"].values, var["start"].values, window_size=window, chunksize=5000)[:2]
    assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in chr_pos_o.items()}
    got, want = res.toarray(), want.toarray()
    both = (got != 0) & (want != 0)
    rel = np.abs(got[both] - want[both]) / np.abs(want[both])
    # CPU measurement with a float64-accumulated profile: max rel 2.0e-6, 0 (w=100) / 2 (w=250) flips of 9e6 / 7.3e6
    n_flips = _compare_thresholded(got, want, 5000, rtol=1e-5, what=f"bench chunk w={window}", max_flips=16)
    print(f"\n[bench chunk w={window}] max rel err {rel.max():.3e}, flips {n_flips} of {got.size}")
    # CSR input of the same matrix: bit-identical result (densify-on-load / sparse-aware smoothing, :115-116,:423)
    adata_s = _adata(sp.csr_matrix(X), var)
    _, res_s, _ = cnv.tl.infercnv(adata_s, window_size=window, chunksize=5000, inplace=False)
    d = (res_s - res)
    assert d.nnz == 0 or np.abs(d.data).max() <= 1e-5 * np.abs(res.data).max()
    print(f"[bench chunk w={window}] CSR vs dense input: {0 if d.nnz == 0 else np.abs(d.data).max():.3e} max abs diff")


@pytest.mark.parametrize("n_cat", [1, 3])
def test_csr_column_sums_are_deterministic(n_cat):
    """icnv_colsum_csr_f32 (icnv_sparse.cu): no atomics — two runs give the same bits, and the sums equal the dense
    kernel's to fp64 rounding; with categories (per-CTA accumulators in the global workspace) too."""
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    G, N = 20000, 7001
    var = cnv.datasets.synthetic_var(G, seed=0)
    Xd = cnv.datasets.device_counts(N, G, dev, seed=99)
    csr = Xd.to_sparse_csr()
    triple = (csr.crow_indices().to(torch.int64), csr.col_indices().to(torch.int32), csr.values())
    row_cat = None
    if n_cat > 1:
        row_cat = torch.from_numpy(np.random.default_rng(3).integers(-1, n_cat, size=N).astype(np.int32)).to(dev)
    with DevicePlan(build_layout(var, 100, 10), dev) as plan:
        s1, c1 = plan.colsum(triple, row_cat, n_cat)
        s2, c2 = plan.colsum(triple, row_cat, n_cat)
        sd, cd = plan.colsum(Xd, row_cat, n_cat)
    assert torch.equal(s1, s2) and torch.equal(c1, c2)
    assert torch.equal(c1, cd)
    np.testing.assert_allclose(s1.cpu().numpy(), sd.cpu().numpy(), rtol=1e-13, atol=1e-12)
    Xh = Xd.cpu().numpy().astype(np.float64)
    for k in range(n_cat):
        rows = np.ones(N, bool) if row_cat is None else (row_cat.cpu().numpy() == k)
        np.testing.assert_allclose(s1[k].cpu().numpy(), Xh[rows].sum(axis=0), rtol=1e-12, atol=1e-12)


def test_float64_output_matches_oracle_to_1e11():
    """C-ABI level: pre-threshold matrix with float64 output against the oracle (no noise filter)."""
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    for g, n_cells, window, step in [(20000, 64, 100, 10), (20000, 32, 250, 10), (6000, 40, 50, 10), (3000, 24, 20, 3)]:
        var = cnv.datasets.synthetic_var(g, seed=3, with_extras=True)
        X = cnv.datasets.synthetic_counts(n_cells, g, seed=g + window)
        ref = X.mean(axis=0, dtype=np.float64).astype(np.float32)[None, :]
        layout = build_layout(var, window, step)
        with DevicePlan(layout, dev) as plan:
            plan.set_reference(torch.from_numpy(ref).to(dev))
            tmp = plan.smooth(torch.from_numpy(X).to(dev), 3.0)
            out, stats = plan.center(tmp, out_dtype=torch.float64)
            got = out.cpu().numpy()
            stats = stats.cpu().numpy()
            tier = plan.tier
        chr_pos, want = orc.infercnv(
            X, var["chromosome"].values, var["start"].values, reference=ref, window_size=window, step=step,
            dynamic_threshold=None,
        )
        want = want.toarray()
        assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in layout.chr_pos.items()}
        scale = np.abs(want).max()
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13 * scale, err_msg=f"tier {tier} window {window}")
        np.testing.assert_allclose(stats[:, 0], want.sum(axis=1), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(stats[:, 1], (want**2).sum(axis=1), rtol=1e-10)


@pytest.mark.parametrize(
    "g,window,step,gene_values",
    [
        (58000, 100, 10, False),  # gene axis longer than shared memory (the tutorial dataset's var): staged in parts
        (12000, 100, 1, False),   # K > 28 tiles of tasks: CTA-per-row median
        (58000, 250, 1, True),    # both, K*8 beyond shared memory, plus the per-gene layer on the same paths
        (9000, 37, 7, True),      # step does not divide the window, odd window
    ],
)
def test_general_fallback_kernels_match_oracle(g, window, step, gene_values):
    """Direct-form smoothing in parts + wide-row centring (csrc/icnv_direct.cu) against the oracle."""
    n = 5
    var = cnv.datasets.synthetic_var(g, seed=11, with_extras=True)
    X = cnv.datasets.synthetic_counts(n, g, seed=g + window)
    ref = X.mean(axis=0, dtype=np.float64).astype(np.float32)
    adata = _adata(X, var)
    chr_pos, res, per_gene = cnv.tl.infercnv(
        adata, reference=ref, window_size=window, step=step, chunksize=3, inplace=False,
        calculate_gene_values=gene_values,
    )
    out = orc.infercnv(
        X, var["chromosome"].values, var["start"].values, reference=ref, window_size=window, step=step, chunksize=3,
        calculate_gene_values=gene_values,
    )
    assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in out[0].items()}
    _compare_thresholded(res.toarray(), out[1].toarray(), 3, what=f"g={g} window={window} step={step}")
    if gene_values:
        _compare_gene_layer(per_gene, out[2], 3, f"gene layer g={g} window={window} step={step}")
    # float64 output of the same path, no filter: 1e-11
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    with DevicePlan(build_layout(var, window, step), dev) as plan:
        plan.set_reference(torch.from_numpy(ref[None, :]).to(dev))
        assert plan.tier in (1, 2)
        got = plan.center(plan.smooth(torch.from_numpy(X).to(dev), 3.0), out_dtype=torch.float64)[0].cpu().numpy()
    want = orc.infercnv(
        X, var["chromosome"].values, var["start"].values, reference=ref, window_size=window, step=step,
        dynamic_threshold=None,
    )[1].toarray()
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13 * np.abs(want).max())


def test_reference_known_answers(full_mock, x_res_actual):
    """/root/reference/tests/test_tools.py:143-191 through the public API (integer matrix, CSR, 2 chunks)."""
    X, var = full_mock
    adata = _adata(sp.csr_matrix(X), var)
    chr_pos, res, _ = cnv.tl.infercnv(
        adata, chunksize=2, lfc_clip=1, window_size=3, step=1, dynamic_threshold=1, inplace=False
    )
    np.testing.assert_allclose(res.toarray(), x_res_actual, rtol=1e-6, atol=0)
    assert chr_pos == {"chr1": 0, "chr2": 3}
    # inplace keys (_infercnv.py:153-155)
    cnv.tl.infercnv(adata, chunksize=2, lfc_clip=1, window_size=3, step=1, dynamic_threshold=1)
    assert "X_cnv" in adata.obsm and adata.uns["cnv"]["chr_pos"] == {"chr1": 0, "chr2": 3}


def test_errors_match_reference():
    var = cnv.datasets.synthetic_var(300, seed=1)
    X = cnv.datasets.synthetic_counts(8, 300, seed=1)
    with pytest.raises(ValueError, match="Genomic positions not found"):
        cnv.tl.infercnv(_adata(X, var.drop(columns=["start"])))
    dup = var.copy()
    dup.index = ["a"] * 300
    with pytest.raises(ValueError, match="unique"):
        cnv.tl.infercnv(_adata(X, dup))
    with pytest.raises(ValueError, match="Reference must match"):
        cnv.tl.infercnv(_adata(X, var), reference=np.ones(299))
    obs = pd.DataFrame({"ct": ["a"] * 8}, index=[str(i) for i in range(8)])
    with pytest.raises(ValueError, match="reference categories were not found"):
        cnv.tl.infercnv(_adata(X, var, obs), reference_key="ct", reference_cat=["a", "zzz"])


def test_layer_equals_x():
    # /root/reference/tests/test_tools.py:221-239
    var = cnv.datasets.synthetic_var(2400, seed=0)
    X = cnv.datasets.synthetic_counts(64, 2400, seed=5)
    a = cnv.AnnData(np.zeros_like(X), var=var, layers={"LogNormalize": X})
    b = cnv.AnnData(X, var=var)
    cnv.tl.infercnv(a, layer="LogNormalize")
    cnv.tl.infercnv(b)
    assert (a.obsm["X_cnv"] != b.obsm["X_cnv"]).nnz == 0


@pytest.mark.parametrize("container", ["dense", "csr"])
def test_multi_block_equals_single_block(container, monkeypatch):
    """A matrix that does not fit the device budget is processed in row blocks (multiples of chunksize, each block
    uploaded for the reference pass and again for the smoothing): same result as the resident single-block path."""
    var = cnv.datasets.synthetic_var(2400, seed=0, with_extras=True)
    X = cnv.datasets.synthetic_counts(700, 2400, seed=21)
    Xin = sp.csr_matrix(X) if container == "csr" else X
    ref = X.mean(axis=0, dtype=np.float64).astype(np.float32)
    kw = dict(chunksize=100, inplace=False, calculate_gene_values=True, reference=ref)
    _, res1, gene1 = cnv.tl.infercnv(_adata(Xin, var), **kw)
    small = str(300 * (4 * 2400 + 14 * res1.shape[1] + 64 + 8 * 2400 + 10 * res1.shape[1]))  # three blocks of <= 300 rows
    monkeypatch.setenv("ICNV_BLOCK_BYTES", small)
    _, res2, gene2 = cnv.tl.infercnv(_adata(Xin, var), **kw)
    assert (res1 != res2).nnz == 0 and res1.shape == res2.shape  # same profile -> bit-identical
    np.testing.assert_array_equal(np.nan_to_num(gene1, nan=-7.0), np.nan_to_num(gene2, nan=-7.0))
    # data-derived profiles (all cells / per category) summed over several blocks: the fp64 sums are added in a
    # different order, so the float32 profile may move by an ulp in a few columns
    obs = pd.DataFrame({"ct": np.array(["a", "b", "c"])[np.arange(700) % 3]}, index=[str(i) for i in range(700)])
    for extra in (dict(), dict(reference_key="ct", reference_cat=["a", "c"])):
        kw = dict(chunksize=100, inplace=False, **extra)
        monkeypatch.setenv("ICNV_BLOCK_BYTES", small)
        _, res3, _ = cnv.tl.infercnv(_adata(Xin, var, obs), **kw)
        monkeypatch.delenv("ICNV_BLOCK_BYTES")
        _, res4, _ = cnv.tl.infercnv(_adata(Xin, var, obs), **kw)
        a3, a4 = res3.toarray(), res4.toarray()
        both = (a3 != 0) & (a4 != 0)
        np.testing.assert_allclose(a3[both], a4[both], rtol=1e-5, atol=1e-7)
        assert ((a3 != 0) != (a4 != 0)).sum() <= 3


def test_csr_with_duplicates_and_unsorted_columns_equals_dense():
    # scipy's toarray() (_infercnv.py:423) sums duplicate entries and does not care about column order
    import scipy.sparse as sp

    var = cnv.datasets.synthetic_var(2400, seed=0)
    X = cnv.datasets.synthetic_counts(70, 2400, seed=9)
    csr = sp.csr_matrix(X)
    # split every stored value into two halves stored at the same (row, col), then shuffle the columns of each row
    rng = np.random.default_rng(0)
    indptr, indices, data = [0], [], []
    for r in range(X.shape[0]):
        cols = csr.indices[csr.indptr[r] : csr.indptr[r + 1]]
        vals = csr.data[csr.indptr[r] : csr.indptr[r + 1]]
        c2 = np.concatenate([cols, cols])
        v2 = np.concatenate([vals * np.float32(0.5), vals * np.float32(0.5)])
        perm = rng.permutation(len(c2))
        indices.append(c2[perm])
        data.append(v2[perm])
        indptr.append(indptr[-1] + len(c2))
    messy = sp.csr_matrix((np.concatenate(data), np.concatenate(indices), np.array(indptr)), shape=X.shape)
    assert not messy.has_canonical_format
    np.testing.assert_array_equal(messy.toarray(), X)
    a = cnv.AnnData(messy, var=var)
    b = cnv.AnnData(csr, var=var)
    cnv.tl.infercnv(a)
    cnv.tl.infercnv(b)
    assert (a.obsm["X_cnv"] != b.obsm["X_cnv"]).nnz == 0


# ---- size-independent properties at (close to) bench shape ----------------------------------------
def test_properties_at_scale():
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    G, N, chunk = 20000, 12000, 5000
    var = cnv.datasets.synthetic_var(G, seed=0)
    Xd = cnv.datasets.device_counts(N, G, dev, seed=1234)
    layout = build_layout(var, 100, 10)
    with DevicePlan(layout, dev) as plan:
        assert plan.K == 1792 and plan.tier == 0
        sums, counts = plan.colsum(Xd)
        ref = plan.mean_from_sums(sums, counts)
        plan.set_reference(ref)
        tmp = plan.smooth(Xd, 3.0)
        pre, stats = plan.center(tmp)
        out = pre.clone()
        thr, row_abs, row_nnz = plan.threshold(out, stats, chunk, 1.5)
        # (1) every row of the pre-threshold matrix has median 0 (even K: mean of the middle pair)
        med = pre.double().median(dim=1).values  # lower median
        srt = pre.double().sort(dim=1).values
        mid = 0.5 * (srt[:, plan.K // 2 - 1] + srt[:, plan.K // 2])
        assert float(mid.abs().max()) < 1e-7
        # (2) per-chunk threshold equals dyn * population std of the chunk
        for c in range(3):
            blk = pre[c * chunk : (c + 1) * chunk].double()
            want = 1.5 * blk.std(unbiased=False).item()
            assert abs(thr[c].item() - want) / want < 1e-6
            kept = out[c * chunk : (c + 1) * chunk]
            assert float(kept[kept != 0].abs().min()) >= thr[c].item() * (1 - 1e-6)
        # (3) statistics rows
        np.testing.assert_allclose(row_abs.cpu().numpy(), out.double().abs().sum(dim=1).cpu().numpy(), rtol=1e-12)
        assert torch.equal(row_nnz.long(), (out != 0).sum(dim=1))
        # (4) row-shard invariance: two shards cut at a chunk boundary give the same matrix
        o1 = plan.center(plan.smooth(Xd[:5000], 3.0))[0]
        o2 = plan.center(plan.smooth(Xd[5000:], 3.0))[0]
        assert torch.equal(torch.cat([o1, o2]), pre)
        # odd row counts (a CTA iteration may stage rows in pairs: the last pair is then half empty)
        for n_odd in (1, 297, 2001):
            assert torch.equal(plan.center(plan.smooth(Xd[7:7 + n_odd], 3.0))[0], pre[7:7 + n_odd])
        # (5) CSR input (deltas of the stored entries added to the smoothed constant row in 48-bit fixed point,
        #     icnv_sparse_delta.cu) == dense input up to the fixed-point rounding (2^-49 per entry, ~1e-15 on a window
        #     mean): after the single rounding to fp32 a few entries in a million may differ, by one ulp
        sub = Xd[:3000]
        csr = sub.to_sparse_csr()
        triple = (csr.crow_indices().to(torch.int64), csr.col_indices().to(torch.int32), csr.values())
        o3 = plan.center(plan.smooth(triple, 3.0))[0]
        diff = o3 != pre[:3000]
        assert int(diff.sum()) <= max(3, int(2e-5 * o3.numel())), f"{int(diff.sum())} entries differ between CSR and dense input"
        print(f"\n[scale] CSR vs dense: {int(diff.sum())} of {o3.numel()} fp32 entries differ (one ulp)")
        assert float((o3 - pre[:3000]).abs().max()) <= 1.2e-7 * float(pre[:3000].abs().max())
        # CSR smoothing is run-to-run bit-reproducible
        assert torch.equal(plan.smooth(triple, 3.0), plan.smooth(triple, 3.0))
        # (6) CSR conversion round trip
        indptr, indices, data = plan.to_csr(out, row_nnz)
        back = torch.sparse_csr_tensor(indptr, indices.long(), data, size=out.shape).to_dense()
        assert torch.equal(back, out)
        # (6b) the fused path (count + scan + filtering compaction on the UNFILTERED block) gives the same CSR, bit for bit
        pre_copy = pre.clone()
        thr2, ra2, nz2, (ip2, ix2, dv2) = plan.filter_to_csr(pre, stats, chunk, 1.5)
        assert torch.equal(pre, pre_copy), "filter_to_csr must not rewrite the dense block"
        assert torch.equal(thr2, thr) and torch.equal(nz2, row_nnz) and torch.equal(ip2, indptr)
        assert torch.equal(ix2, indices) and torch.equal(dv2, data)
        np.testing.assert_allclose(ra2.cpu().numpy(), row_abs.cpu().numpy(), rtol=1e-14)
        _, _, _, (_, ix3, dv3) = plan.filter_to_csr(pre, stats, chunk, 1.5, data_dtype=torch.float64)
        assert dv3.dtype == torch.float64 and torch.equal(dv3, data.double()) and torch.equal(ix3, indices)
        # no filter: every non-zero survives
        _, _, nz4, (ip4, ix4, dv4) = plan.filter_to_csr(pre, stats, chunk, None)
        assert torch.equal(torch.sparse_csr_tensor(ip4, ix4.long(), dv4, size=pre.shape).to_dense(), pre)
    # (7) permuting the gene columns together with var leaves the result unchanged
    perm = np.random.default_rng(0).permutation(G)
    var_p = var.iloc[perm]
    layout_p = build_layout(var_p, 100, 10)
    with DevicePlan(layout_p, dev) as plan_p:
        plan_p.set_reference(ref[:, torch.from_numpy(perm).to(dev)].contiguous())
        o4 = plan_p.center(plan_p.smooth(Xd[:2000][:, torch.from_numpy(perm).to(dev)].contiguous(), 3.0))[0]
    assert torch.equal(o4, pre[:2000])


@pytest.mark.parametrize("n", [0, 1, 1023, 8192, 8193, 70001])
def test_indptr_scan_matches_cumsum(n):
    torch = _torch()
    from infercnvpy_b200 import _lib

    dev = torch.device("cuda", 0)
    lib = _lib.load()
    g = torch.Generator(device=dev)
    g.manual_seed(n)
    nnz = torch.randint(0, 1793, (n,), generator=g, device=dev, dtype=torch.int32)
    indptr = torch.full((n + 1,), -7, dtype=torch.int64, device=dev)
    _lib.check(lib.icnv_nnz_to_indptr(_lib.ptr(nnz), n, _lib.ptr(indptr), _lib.stream_handle(dev)), "icnv_nnz_to_indptr")
    want = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    want[1:] = torch.cumsum(nnz.long(), 0)
    assert torch.equal(indptr, want)


def test_properties_at_full_bench_size():
    """BASELINE configs[1] at full size (100 000 cells x 20 000 genes fp32, window 100, chunks of 5000): properties that
    need no oracle — bit-reproducibility, shard invariance, exact medians, thresholds, statistics."""
    torch = _torch()
    from infercnvpy_b200._engine import DevicePlan
    from infercnvpy_b200._layout import build_layout

    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 24 << 30:
        pytest.skip("needs 24 GiB of free device memory")
    G, N, chunk = 20000, 100_000, 5000
    var = cnv.datasets.synthetic_var(G, seed=0)
    Xd = cnv.datasets.device_counts(N, G, dev, seed=4321)
    with DevicePlan(build_layout(var, 100, 10), dev) as plan:
        sums, counts = plan.colsum(Xd)
        # the reference profile is the column mean (fp64 sums): check a strided sample of columns against torch
        cols = torch.arange(0, G, 97, device=dev)
        want_mean = Xd[:, cols].double().mean(dim=0)
        ref = plan.mean_from_sums(sums, counts)
        assert int(counts[0]) == N
        torch.testing.assert_close(ref[0, cols].double(), want_mean, rtol=1e-6, atol=1e-9)
        plan.set_reference(ref)
        tmp = plan.smooth(Xd, 3.0)
        pre, stats = plan.center(tmp)
        # (1) bit-reproducible from launch to launch (dynamic work hand-out must not change a single bit)
        tmp2 = plan.smooth(Xd, 3.0)
        assert torch.equal(tmp2.view(torch.int64), tmp.view(torch.int64))
        pre2, stats2 = plan.center(tmp2)
        assert torch.equal(pre2, pre) and torch.equal(stats2, stats)
        del tmp2, pre2, stats2
        # (2) row shards cut at chunk boundaries give the same matrix (what the multi-GPU path relies on)
        cut = 35_000
        top = plan.center(plan.smooth(Xd[:cut], 3.0))[0]
        assert torch.equal(top, pre[:cut])
        bottom = plan.center(plan.smooth(Xd[cut:], 3.0))[0]
        assert torch.equal(bottom, pre[cut:])
        del top, bottom
        # (3) every row is centred on its exact median (even K: mean of the middle pair), checked on 4000 sampled rows
        rows = torch.randperm(N, device=dev, generator=torch.Generator(device=dev).manual_seed(1))[:4000]
        srt = pre[rows].double().sort(dim=1).values
        mid = 0.5 * (srt[:, plan.K // 2 - 1] + srt[:, plan.K // 2])
        assert float(mid.abs().max()) < 1e-7
        # (4) per-chunk threshold = 1.5 * population std of the chunk; kept entries are never below it
        out = pre.clone()
        thr, row_abs, row_nnz = plan.threshold(out, stats, chunk, 1.5)
        assert thr.numel() == N // chunk
        for c in (0, 7, 19):
            blk = pre[c * chunk : (c + 1) * chunk].double()
            want = 1.5 * blk.std(unbiased=False).item()
            assert abs(thr[c].item() - want) / want < 1e-6
            kept = out[c * chunk : (c + 1) * chunk]
            assert float(kept[kept != 0].abs().min()) >= thr[c].item() * (1 - 1e-6)
            assert torch.equal(kept != 0, blk.abs() >= thr[c])  # strict `<` of _infercnv.py:451, compared in float64
        # (5) statistics and CSR round trip over the whole matrix
        np.testing.assert_allclose(row_abs.cpu().numpy(), out.double().abs().sum(dim=1).cpu().numpy(), rtol=1e-12)
        assert torch.equal(row_nnz.long(), (out != 0).sum(dim=1))
        indptr, indices, data = plan.to_csr(out, row_nnz)
        assert int(indptr[-1]) == int(row_nnz.sum()) and 0.05 < int(indptr[-1]) / out.numel() < 0.3
        r = 54_321
        dense_row = torch.zeros(plan.K, device=dev)
        dense_row[indices[indptr[r] : indptr[r + 1]].long()] = data[indptr[r] : indptr[r + 1]]
        assert torch.equal(dense_row, out[r])


def test_cnv_score_known_answer_and_golden(golden_loader):
    # /root/reference/tests/test_scores.py:18-21
    X_cnv = np.array([[1, 1, 1, 2, 2, 1, 1, 1], [2, 2, 2, 1, 1, 2, 2, 2], [4, 4, 4, 2, 2, 3, 3, 3], [2, 2, 2, 4, 4, 4, 4, 4]]).T
    obs = pd.DataFrame({"group": list("AAAAABBB")}, index=[f"c{i}" for i in range(8)])
    for container in (np.array, sp.csr_matrix, sp.csc_matrix):
        a = cnv.AnnData(np.zeros((8, 3)), obs=obs.copy(), obsm={"X_cnv": container(X_cnv)})
        res = cnv.tl.cnv_score(a, "group", inplace=False)
        assert res["A"] == pytest.approx(2.25, abs=1e-3) and res["B"] == pytest.approx(2.5, abs=1e-3)
        cnv.tl.cnv_score(a, "group")
        np.testing.assert_allclose(a.obs["cnv_score"].values, [2.25] * 5 + [2.5] * 3)
    with pytest.raises(ValueError, match="cnv_leiden"):
        cnv.tl.cnv_score(cnv.AnnData(np.zeros((8, 3)), obs=obs.copy(), obsm={"X_cnv": X_cnv}))
    gold = golden_loader("small_default")
    obs = pd.DataFrame({"grp": gold["score_labels"]})
    a = cnv.AnnData(np.zeros((gold["csr"].shape[0], 1)), obs=obs, obsm={"X_cnv": gold["csr"]})
    res = cnv.tl.cnv_score(a, "grp", inplace=False)
    for k, v in zip(gold["score_keys"].tolist(), gold["score_vals"].tolist()):
        assert float(res[k]) == pytest.approx(v, rel=1e-12)


# ---- ITH scores (tl/_scores.py:77-221) --------------------------------------------------------------
def test_ith_scores_known_answers_and_golden():
    from tests.golden.make_golden_ith import ith_case
    from tests.test_oracle_golden import ITH_CNV, ITH_EXPR, ITH_GROUPS

    # /root/reference/tests/test_scores.py:6-15
    obs = pd.DataFrame({"group": ITH_GROUPS}, index=[f"c{i}" for i in range(8)])
    for container in (np.array, sp.csr_matrix, sp.csc_matrix):
        a = cnv.AnnData(container(ITH_EXPR), obs=obs.copy(), obsm={"X_cnv": container(ITH_CNV)})
        gex = cnv.tl.ithgex(a, "group", inplace=False)
        assert gex["A"] == 0 and gex["B"] == pytest.approx(1.2628, abs=0.001)
        cna = cnv.tl.ithcna(a, "group", inplace=False)
        assert cna["A"] == pytest.approx(1.053, abs=0.001) and cna["B"] == 0
        cnv.tl.ithgex(a, "group")
        cnv.tl.ithcna(a, "group")
        np.testing.assert_allclose(a.obs["ithgex"].values, [gex["A"]] * 5 + [gex["B"]] * 3)
        np.testing.assert_allclose(a.obs["ithcna"].values, [cna["A"]] * 5 + [cna["B"]] * 3)
    with pytest.raises(ValueError, match="both layer and raw"):
        cnv.tl.ithgex(a, "group", use_raw=True, layer="x")
    # outputs of the unmodified reference on a seeded 260-cell case (tests/golden/make_golden_ith.py)
    z = np.load(Path(__file__).resolve().parent / "golden" / "ith_scores.npz", allow_pickle=False)
    expr, x_cnv, labels = ith_case()
    obs = pd.DataFrame({"patient": labels}, index=[f"c{i}" for i in range(len(labels))])
    a = cnv.AnnData(expr, obs=obs, obsm={"X_cnv": x_cnv})
    gex = cnv.tl.ithgex(a, "patient", inplace=False)
    cna = cnv.tl.ithcna(a, "patient", inplace=False)
    assert sorted(gex) == z["keys"].tolist() == sorted(cna)
    for k, g, c in zip(z["keys"].tolist(), z["ithgex"].tolist(), z["ithcna"].tolist()):
        assert float(gex[k]) == pytest.approx(g, rel=1e-9) and float(cna[k]) == pytest.approx(c, rel=1e-9)
    with pytest.raises(KeyError):  # the one-cell group has no score to broadcast (_scores.py:144-146)
        cnv.tl.ithgex(a, "patient")
    # a cell without variance poisons its group with NaN, like np.corrcoef / np.percentile
    flat = expr.copy()
    flat[3] = 0.25
    b = cnv.AnnData(flat, obs=obs.copy())
    res = cnv.tl.ithgex(b, "patient", inplace=False)
    want = orc.ith_score(flat, labels)
    for k in want:
        assert np.isnan(res[k]) == np.isnan(want[k])
        if not np.isnan(want[k]):
            assert float(res[k]) == pytest.approx(float(want[k]), rel=1e-9)
    # a group above 2048 cells takes the selection path (no sort of the n_g^2 correlations): same numbers as numpy
    rng = np.random.default_rng(11)
    big = (rng.normal(size=(2600, 90)) + rng.normal(size=(1, 90)) * 0.7).astype(np.float64)
    lab = np.where(np.arange(2600) < 2500, "big", "small")
    c = cnv.AnnData(big, obs=pd.DataFrame({"g": lab}, index=[f"c{i}" for i in range(2600)]))
    got = cnv.tl.ithgex(c, "g", inplace=False)
    want = orc.ith_score(big, lab)
    for k in want:
        assert float(got[k]) == pytest.approx(float(want[k]), rel=1e-9), k


def test_group_means_match_numpy(golden_loader):
    """pl.chromosome_heatmap_summary's per-group column means (/root/reference/src/infercnvpy/pl/_chromosome_heatmap.py:
    149-158) against the reference expression evaluated with numpy/scipy on the same CSR, sparse and dense containers."""
    case = next(c for c in CASES if c["name"] == "small_default")
    X, var, obs, kw = build_case(case)
    adata = _adata(X, var, obs)
    cnv.tl.infercnv(adata, **kw)
    rng = np.random.default_rng(4)
    adata.obs = pd.DataFrame({"grp": rng.choice(["b", "a", "z", "m"], size=X.shape[0])}, index=adata.obs.index)
    res = cnv.pl.chromosome_heatmap_summary(adata, groupby="grp")
    assert res["groups"] == list(adata.obs["grp"].unique())
    Xc = adata.obsm["X_cnv"]
    for i, g in enumerate(res["groups"]):
        want = np.asarray(np.mean(Xc[adata.obs["grp"].values == g, :], axis=0)).ravel()
        np.testing.assert_allclose(res["mean"][i], want, rtol=1e-12, atol=1e-15)
    assert res["var_group_positions"][0][0] == 0 and res["var_group_positions"][-1][1] == Xc.shape[1]
    assert res["var_group_labels"] == list(adata.uns["cnv"]["chr_pos"].keys())
    adata.obsm["X_dense"] = Xc.toarray()
    g2, m2 = cnv.pl.group_means(adata, "grp", use_rep="dense")
    np.testing.assert_allclose(m2, res["mean"], rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError, match="cnv_leiden"):
        cnv.pl.chromosome_heatmap_summary(adata)


def test_disk_to_hbm_loader_shards_and_feeds_infercnv(tmp_path):
    """infercnvpy_b200.io.read_matrix: every container lands in HBM bit-identical to the file, row
    shards are cut at multiples of chunksize, and a DeviceCSR / device tensor feeds tl.infercnv like the host matrix."""
    torch = _torch()
    from infercnvpy_b200 import io as cio

    var = cnv.datasets.synthetic_var(2400, seed=0, with_extras=True)
    X = cnv.datasets.synthetic_counts(730, 2400, seed=5)
    A = sp.csr_matrix(X)
    np.save(tmp_path / "x.npy", X)
    sp.save_npz(tmp_path / "a.npz", A, compressed=False)
    d = tmp_path / "csrdir"
    d.mkdir()
    for name, arr in (("indptr", A.indptr), ("indices", A.indices), ("data", A.data), ("shape", np.array(A.shape))):
        np.save(d / f"{name}.npy", arr)
    # dense, small slabs so several staging rounds happen
    Xd, (r0, r1), n = cio.read_matrix(tmp_path / "x.npy", slab_bytes=1 << 20)
    assert (r0, r1, n) == (0, 730, 730) and Xd.is_cuda and np.array_equal(Xd.cpu().numpy(), X)
    parts = [cio.read_matrix(tmp_path / "x.npy", rank=r, world=3, chunksize=100, slab_bytes=1 << 19) for r in range(3)]
    assert [p[1] for p in parts] == [cnv.shard_rows(730, 100, r, 3) for r in range(3)]
    assert np.array_equal(np.concatenate([p[0].cpu().numpy() for p in parts]), X)
    for path in (tmp_path / "a.npz", d):
        S, (r0, r1), n = cio.read_matrix(path, slab_bytes=1 << 18)
        assert isinstance(S, cio.DeviceCSR) and S.shape == A.shape and n == 730
        assert np.array_equal(S.indptr.cpu().numpy(), A.indptr) and np.array_equal(S.indices.cpu().numpy(), A.indices)
        assert np.array_equal(S.data.cpu().numpy(), A.data)
        S1, (a, b), _ = cio.read_matrix(path, rank=1, world=2, chunksize=100)
        sub = A[a:b]
        assert np.array_equal(S1.indptr.cpu().numpy(), sub.indptr) and np.array_equal(S1.data.cpu().numpy(), sub.data)
    # the loaded containers are drop-in matrices for tl.infercnv
    kw = dict(chunksize=100, inplace=False)
    _, want, _ = cnv.tl.infercnv(_adata(A, var), **kw)
    _, got_csr, _ = cnv.tl.infercnv(_adata(S, var), **kw)
    assert (got_csr != want).nnz == 0
    _, want_d, _ = cnv.tl.infercnv(_adata(X, var), **kw)
    _, got_d, _ = cnv.tl.infercnv(_adata(Xd, var), **kw)
    assert (got_d != want_d).nnz == 0


def test_edge_cases_empty_ragged_and_degenerate_rows():
    """Empty and ragged inputs through the public API against the oracle: no cells, one cell, a chunk size larger than
    the matrix, CSR rows without any stored entry, a fully dense CSR row, explicitly stored zeros."""
    var = cnv.datasets.synthetic_var(2400, seed=0, with_extras=True)
    chrom, start = var["chromosome"].values, var["start"].values
    X = cnv.datasets.synthetic_counts(23, 2400, seed=77)
    ref = X.mean(axis=0, dtype=np.float64).astype(np.float32)
    # no cells: the reference's process_map gets no chunk and vstack fails; here an empty CSR of the right width
    for empty in (X[:0], sp.csr_matrix(X[:0])):
        chr_pos, res, _ = cnv.tl.infercnv(_adata(empty, var), reference=ref, inplace=False)
        assert res.shape[0] == 0 and res.shape[1] > 0 and res.nnz == 0 and list(chr_pos)[0] == "chr1"
    # one cell / chunksize larger than the matrix / ragged last chunk
    for n, chunk in ((1, 5000), (23, 5000), (23, 7)):
        chr_pos_o, want = orc.infercnv(X[:n], chrom, start, reference=ref, chunksize=chunk)[:2]
        for Xin in (X[:n], sp.csr_matrix(X[:n])):
            chr_pos, res, _ = cnv.tl.infercnv(_adata(Xin, var), reference=ref, chunksize=chunk, inplace=False)
            assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in chr_pos_o.items()}
            _compare_thresholded(res.toarray(), want.toarray(), chunk, what=f"n={n} chunk={chunk}", max_flips=2)
    # CSR with empty rows, one fully dense row and explicitly stored zeros
    Y = X.copy()
    Y[3] = 0.0
    Y[11] = 0.0
    Y[5] = np.float32(0.5) + Y[5]
    S = sp.csr_matrix(Y)
    S.data[::17] = 0.0  # stored zeros (the dense twin gets the same zeros)
    Yd = S.toarray()
    assert S.nnz > (Yd != 0).sum()
    _, want = orc.infercnv(Yd, chrom, start, reference=ref, chunksize=10)[:2]
    _, got_s, _ = cnv.tl.infercnv(_adata(S, var), reference=ref, chunksize=10, inplace=False)
    _, got_d, _ = cnv.tl.infercnv(_adata(Yd, var), reference=ref, chunksize=10, inplace=False)
    _compare_thresholded(got_s.toarray(), want.toarray(), 10, what="ragged csr", max_flips=2)
    _compare_thresholded(got_d.toarray(), want.toarray(), 10, what="ragged dense", max_flips=2)
