"""Host-side logic and the C-ABI surface, no GPU needed."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import infercnvpy_b200 as cnv
from infercnvpy_b200 import _lib
from infercnvpy_b200._layout import build_layout
from oracle import infercnv_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "icnv.h").read_text()
    declared = set(re.findall(r"\b(icnv_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), "ctypes table and include/icnv.h disagree"
    lib_path = _lib.lib_path()
    assert lib_path.exists(), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(str(lib_path))
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported"
    handle.icnv_version.restype = ctypes.c_int
    assert handle.icnv_version() >= 100


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    var = cnv.datasets.synthetic_var(300, seed=1)
    X = cnv.datasets.synthetic_counts(4, 300, seed=1)
    with pytest.raises(_lib.IcnvError, match="CUDA"):
        cnv.tl.infercnv(cnv.AnnData(X, var=var))


@pytest.mark.parametrize("window,step,extras", [(100, 10, True), (250, 10, False), (11, 1, True), (20, 3, False), (5000, 10, True)])
def test_layout_matches_oracle_indices(window, step, extras):
    """chr_pos / gene permutation are integers and must match the oracle exactly."""
    var = cnv.datasets.synthetic_var(2400, seed=4, with_extras=extras)
    lay = build_layout(var, window, step)
    X = np.zeros((1, 2400), dtype=np.float32)
    chr_pos, res = orc.infercnv(X, var["chromosome"].values, var["start"].values, window_size=window, step=step, dynamic_threshold=None)
    assert list(chr_pos) == lay.chromosomes
    assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in lay.chr_pos.items()}
    assert res.shape[1] == lay.n_out
    keep = ~lay.var_mask
    chrom_k, start_k = var["chromosome"].values[keep], var["start"].values[keep]
    kept_cols = np.flatnonzero(keep)
    for i, c in enumerate(lay.chromosomes):
        want = kept_cols[orc.gene_order(chrom_k, start_k, c)]
        np.testing.assert_array_equal(lay.gene_idx[lay.seg_off[i] : lay.seg_off[i + 1]], want)


def test_layout_ties_follow_pandas():
    var = pd.DataFrame({"chromosome": ["chr1"] * 40, "start": [5] * 20 + [1] * 20, "end": 0}, index=[f"g{i}" for i in range(40)])
    lay = build_layout(var, 3, 1)
    want = var.loc[var["chromosome"] == "chr1"].sort_values("start").index.map(lambda s: int(s[1:])).to_numpy()
    np.testing.assert_array_equal(lay.gene_idx, want)


def test_layout_errors():
    var = cnv.datasets.synthetic_var(50, seed=0)
    with pytest.raises(ValueError, match="Genomic positions not found"):
        build_layout(var.drop(columns=["end"]), 10, 1)
    with pytest.raises(ValueError):
        build_layout(var, 0, 1)


def test_duck_anndata_slicing():
    var = cnv.datasets.synthetic_var(30, seed=0)
    a = cnv.AnnData(np.arange(60, dtype=np.float32).reshape(2, 30), var=var)
    b = a[:, np.arange(30) % 2 == 0]
    assert b.shape == (2, 15) and list(b.var_names) == list(var.index[::2])
    assert a.copy().X is not a.X


def _group_table(var, window, step):
    """Position-ordered groups of `step` genes per chromosome, like icnv_plan_create builds them."""
    lay = build_layout(var, window, step)
    rows = []
    for c in range(len(lay.chromosomes)):
        s0, s1 = int(lay.seg_off[c]), int(lay.seg_off[c + 1])
        g_c = s1 - s0
        flat = not (window < g_c)
        n_out = 1 if flat else (g_c - window) // step + 1
        n_grp = -(-g_c // step) if flat else ((n_out - 1) * step + window) // step
        for g in range(n_grp):
            rows.append([lay.gene_idx[s0 + g * step + j] if g * step + j < g_c else -1 for j in range(step)])
    return np.asarray(rows, dtype=np.int32), var.shape[0]


@pytest.mark.parametrize("g,sort_in_memory", [(20000, False), (20000, True), (2400, False)])
def test_gather_schedule_invariants(g, sort_in_memory):
    """csrc/icnv_schedule.cu on the host: every group gets exactly one lane slot (lane % 8 == group % 8), every lane's
    walk is a permutation of its group, and the permuted walk costs far fewer shared-memory wavefronts than the
    natural one (the kernel itself only decodes what this produces; GPU parity tests cover the arithmetic)."""
    lib = _lib.load()
    var = cnv.datasets.synthetic_var(g, seed=0)
    if sort_in_memory:
        var = var.sort_values(["chromosome", "start"])
    gcol, n_genes = _group_table(var, 100, 10)
    n_groups, gs = gcol.shape
    nsets = -(-(-(-n_groups // 4)) // 32) * 4
    cost = {}
    for permute in (0, 1):
        slot = np.empty(nsets * 32, dtype=np.int32)
        order = np.empty(nsets * 32 * gs, dtype=np.uint8)
        cost[permute] = lib.icnv_host_schedule_gathers(
            gcol.ctypes.data_as(_lib.c_i32p), n_groups, gs, n_genes, nsets, permute,
            slot.ctypes.data_as(_lib.c_i32p), order.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
        )
        used = slot[slot >= 0]
        assert sorted(used.tolist()) == list(range(n_groups))
        lanes = np.flatnonzero(slot >= 0) % 32
        assert np.all(lanes % 8 == used % 8)
        walks = order.reshape(nsets * 32, gs)[slot >= 0]
        assert np.all(np.sort(walks, axis=1) == np.arange(gs))
        if not permute:
            assert np.all(walks == np.arange(gs))
        # recompute the cost independently: distinct words per bank, the pad word counted once in bank n_genes % 32
        total = 0
        for s in range(nsets):
            grp = slot[s * 32 : (s + 1) * 32]
            for t in range(gs):
                cols = np.array([gcol[grp[l], order[(s * 32 + l) * gs + t]] if grp[l] >= 0 else -1 for l in range(32)])
                cnt = np.bincount(cols[cols >= 0] & 31, minlength=32)
                if (cols < 0).any():
                    cnt[n_genes & 31] += 1
                total += max(1, int(cnt.max()))
        assert cost[permute] == pytest.approx(total / (nsets * gs), rel=1e-12)
    assert cost[1] < 1.5 and cost[1] < 0.75 * cost[0]
    # too few slots -> refused
    assert lib.icnv_host_schedule_gathers(gcol.ctypes.data_as(_lib.c_i32p), n_groups, gs, n_genes, 4, 1, None, None) < 0


def test_quantile_rule_matches_numpy_percentile():
    """tl/_ith.py reads the quartiles of the sorted correlation entries with numpy's default ('linear') rule
    (_scores.py:141,214 call np.percentile(pcorr, [75, 25]))."""
    import torch

    from infercnvpy_b200.tl._ith import _np_linear_quantile

    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 5, 9, 16, 25, 1000, 4097):
        a = np.sort(rng.normal(size=n))
        t = torch.from_numpy(a)
        for q in (0.25, 0.75):
            assert _np_linear_quantile(t, n, q) == float(np.percentile(a, 100 * q)), (n, q)
    # ties and a constant vector
    a = np.sort(np.repeat(rng.normal(size=7), 5))
    for q in (0.25, 0.75):
        assert _np_linear_quantile(torch.from_numpy(a), a.size, q) == float(np.percentile(a, 100 * q))
    ones = torch.ones(36, dtype=torch.float64)
    assert _np_linear_quantile(ones, 36, 0.75) - _np_linear_quantile(ones, 36, 0.25) == 0.0


def test_block_rows_are_multiples_of_chunksize(monkeypatch):
    """Row blocks never cut a chunk (its std must see all of its rows, _infercnv.py:123,450)."""
    from infercnvpy_b200.tl._infercnv import _block_rows

    per_row = 4 * 20000 + 14 * 1792 + 64
    monkeypatch.setenv("ICNV_BLOCK_BYTES", str(12_345 * per_row))
    assert _block_rows(100_000, 20000, 1792, 5000) == 10_000
    assert _block_rows(7_000, 20000, 1792, 5000) == 7_000          # everything fits: one block
    assert _block_rows(100_000, 20000, 1792, 20_000) == 20_000     # budget below one chunk: still a whole chunk
    assert _block_rows(0, 20000, 1792, 5000) == 0
    # the per-gene layer needs 8*G more bytes per row
    assert _block_rows(100_000, 20000, 1792, 1000, gene_values=True) < _block_rows(100_000, 20000, 1792, 1000)


def test_layout_matches_oracle_on_random_var_tables():
    """Integer parity of the gene axis on adversarial var tables: ties in `start`, chromosomes that are excluded, null,
    not `chr*`, `chrM`, names that only differ in case or need the natural sort (chr2 < chr10 < chr10_alt)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    names = ["chr1", "chr2", "chr10", "chr10_alt", "chrX", "chrY", "chrM", "chrUn_x", "scaffold7", "Chr3", "chr3", None]

    @settings(max_examples=60, deadline=None)
    @given(
        st.integers(min_value=1, max_value=400),
        st.integers(min_value=0, max_value=2**31 - 1),
        st.sampled_from([(5, 1), (7, 3), (10, 10), (4, 9), (100, 10)]),
        st.sampled_from([("chrX", "chrY"), None, ("chr1",)]),
    )
    def check(g, seed, ws, exclude):
        rng = np.random.default_rng(seed)
        window, step = ws
        chrom = rng.choice(np.array(names, dtype=object), size=g)
        start = rng.integers(0, max(2, g // 3), size=g)  # plenty of ties
        var = pd.DataFrame({"chromosome": chrom, "start": start, "end": start + 10}, index=[f"g{i}" for i in range(g)])
        lay = build_layout(var, window, step, exclude)
        keep = ~lay.var_mask
        chrom_s = pd.Series(chrom, dtype=object)
        drop = chrom_s.isnull()
        if exclude is not None:
            drop = drop | chrom_s.isin(exclude)
        np.testing.assert_array_equal(lay.var_mask, drop.to_numpy())
        order = orc.natural_chromosome_order(chrom[keep])
        assert lay.chromosomes == order
        kept_cols = np.flatnonzero(keep)
        widths = []
        for i, c in enumerate(order):
            want = kept_cols[orc.gene_order(chrom[keep], start[keep], c)]
            np.testing.assert_array_equal(lay.gene_idx[lay.seg_off[i] : lay.seg_off[i + 1]], want)
            g_c = want.size
            widths.append((g_c - window) // step + 1 if window < g_c else 1)
        np.testing.assert_array_equal(lay.out_off, np.cumsum([0] + widths))
        if order:
            X = rng.normal(size=(2, g)).astype(np.float32)
            chr_pos, res = orc.infercnv(X, chrom, start, window_size=window, step=step, dynamic_threshold=None,
                                        exclude_chromosomes=exclude, reference=np.zeros(g, np.float32))
            assert {k: int(v) for k, v in chr_pos.items()} == {k: int(v) for k, v in lay.chr_pos.items()}
            assert res.shape[1] == lay.n_out

    check()


@pytest.mark.parametrize("g,window,step", [(20000, 100, 10), (9000, 100, 10)])
def test_banded_layout_reproduces_the_oracle_smoothing(g, window, step):
    """EXPERIMENTAL banded kernel (csrc/icnv_smooth_banded.cu, off by default): its table layout is host-only code
    (csrc/icnv_banded_host.cu).  Here the kernel's data flow is emulated in numpy FROM THOSE TABLES — gathers through
    (address, element position) entries into per-group partial sums A = sum x, B = sum j*x at the physical group
    indices, then the window formula per task — and compared with the oracle's smoothing.  Also checks that the two
    bands touch disjoint ranges of the partial-sum array (what lets their phases overlap)."""
    lib = _lib.load()
    var = cnv.datasets.synthetic_var(g, seed=0)
    lay = build_layout(var, window, step)
    gcol, n_genes = _group_table(var, window, step)
    n_groups, gs = gcol.shape
    nq = window // step
    # task list like icnv_plan_create: up to 9 consecutive outputs per task, one task per flat chromosome
    tasks, flat_genes, gbase = [], [], 0
    for ci in range(len(lay.chromosomes)):
        g_c = int(lay.seg_off[ci + 1] - lay.seg_off[ci])
        out0 = int(lay.out_off[ci])
        if window < g_c:
            n_out = (g_c - window) // step + 1
            for k in range(0, n_out, 9):
                tasks.append((gbase + k, out0 + k, min(9, n_out - k), 0))
            gbase += n_out - 1 + nq
        else:
            n_grp = -(-g_c // step)
            tasks.append((gbase, out0, n_grp, 1 | (len(flat_genes) << 8)))
            flat_genes.append(g_c)
            gbase += n_grp
    assert gbase == n_groups
    tasks = np.asarray(tasks, dtype=np.int32)
    raw_base = 0x420
    meta = np.zeros(9, dtype=np.int32)
    cap_e, cap_g = (n_groups // 128 + 4) * gs * 128, (n_groups // 128 + 4) * 128
    off = np.zeros(cap_e, dtype=np.uint32)
    cols = np.zeros(cap_e, dtype=np.int32)
    grp = np.zeros(cap_g, dtype=np.int32)
    tasks_b = np.zeros_like(tasks)
    rc = lib.icnv_host_banded_layout(
        gcol.ctypes.data_as(_lib.c_i32p), n_groups, gs, nq, tasks.ctypes.data_as(_lib.c_i32p), len(tasks), n_genes, raw_base,
        meta.ctypes.data_as(_lib.c_i32p), off.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cols.ctypes.data_as(_lib.c_i32p),
        cap_e, grp.ctypes.data_as(_lib.c_i32p), cap_g, tasks_b.ctypes.data_as(_lib.c_i32p),
    )
    assert rc == 0
    on, NG, NGpad, units_a, units_b, tile0_b, tiles_a, tiles_b, n_entries = (int(v) for v in meta)
    n_tiles = -(-len(tasks) // 32)
    assert on == 1 and tiles_a + tiles_b == n_tiles and tile0_b == tiles_a and max(tiles_a, tiles_b) <= 4
    n_units = units_a + units_b
    assert n_entries == n_units * gs * 128
    off, cols = off[:n_entries].reshape(n_units, gs, 32, 4), cols[:n_entries].reshape(n_units, gs, 32, 4)
    grp = grp[: n_units * 128].reshape(n_units, 32, 4)
    # entries: address of the gene (or of the zero pad slot) + element position in bits 24..27
    addr, pos = off & 0xFFFFFF, off >> 24
    np.testing.assert_array_equal(addr, raw_base + 4 * np.where(cols < 0, n_genes, cols))
    assert pos.max() < gs
    # the bands own disjoint ranges of the partial-sum array, separated by >= 12 untouched groups
    a_hi = grp[:units_a].max()
    b_lo = grp[units_a:].min()
    real_a = sorted(set(grp[:units_a].ravel().tolist()))
    assert b_lo - (real_a[-2] if len(real_a) > 1 else a_hi) > 12 and b_lo % 8 == 0 and grp.max() <= NGpad
    real_slot = (cols >= 0).any(axis=1)  # [unit, lane, u]: the slot holds a group with at least one real gene
    lanes = np.broadcast_to(np.arange(32)[None, :, None], grp.shape)
    assert np.all(grp[real_slot] % 8 == lanes[real_slot] % 8)  # a quarter-warp's partial-sum stores hit 8 bank groups
    tA = 32 * tiles_a
    np.testing.assert_array_equal(tasks_b[:tA], tasks[:tA])
    np.testing.assert_array_equal(tasks_b[tA:, 1:], tasks[tA:, 1:])
    assert tasks_b[tA, 0] == b_lo
    # ---- emulate the kernel on two cell rows
    rng = np.random.default_rng(5)
    X = cnv.datasets.synthetic_counts(2, g, seed=17)
    ref = rng.uniform(0.0, 0.5, size=g).astype(np.float32)
    d = np.clip(X - ref, -3, 3).astype(np.float32)
    dpad = np.concatenate([d, np.zeros((2, 1), np.float32)], axis=1).astype(np.float64)  # slot n_genes = zero pad
    AB = np.zeros((2, NGpad + 12, 2))
    src = np.where(cols < 0, n_genes, cols)
    for u in range(n_units):
        vals = dpad[:, src[u]]                               # [2, gs, 32, 4]
        a = vals.sum(axis=1)
        b = (vals * pos[u][None].astype(np.float64)).sum(axis=1)
        AB[:, grp[u].ravel(), 0] = a.reshape(2, -1)
        AB[:, grp[u].ravel(), 1] = b.reshape(2, -1)
    w0 = np.array([min(10 * q + 1, window - 10 * q) for q in range(nq)], dtype=np.float64)            # weight of j = 0
    w1 = np.array([min(10 * q + 2, window - 10 * q - 1) for q in range(nq)], dtype=np.float64) - w0  # slope inside the group
    out = np.full((2, lay.n_out), np.nan)
    sumw = float(sum(min(j + 1, window - j) for j in range(window)))
    for x, y, z, w in tasks_b.tolist():
        if (w & 0xFF) == 0:
            for i in range(z):
                seg = AB[:, x + i : x + i + nq]
                out[:, y + i] = (seg[:, :, 0] * w0 + seg[:, :, 1] * w1).sum(axis=1) / sumw
        else:
            out[:, y] = AB[:, x : x + z, 0].sum(axis=1) / flat_genes[w >> 8]
    chr_pos, want = orc.smooth_by_chromosome(d, var["chromosome"].values, var["start"].values, window, step)
    np.testing.assert_allclose(out, want, rtol=1e-12, atol=1e-15)
