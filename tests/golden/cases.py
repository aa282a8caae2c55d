"""Seeded parity cases shared by the golden generator and the tests.

Every case is small enough for the CPU oracle to finish in seconds and is
chosen to hit one branch of ``tl/_infercnv.py`` (reference file:line in the
comments).  ``build_case`` regenerates the inputs from the seeds; the outputs
of the real reference live next to this file as ``<name>.npz``.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from infercnvpy_b200.datasets import synthetic_counts, synthetic_var

CASES = [
    # default parameters; 2400 genes over 22 chromosomes + chrX/chrY/chrM/contig/NaN genes
    # -> masks of :104-108 and :327, many chromosomes with G_c <= window (:227-236),
    # three chunks with a ragged last one (:123), all-cells float32 mean (:380-385)
    dict(name="small_default", n=300, g=2400, extras=True, seed=1013, kw=dict(chunksize=128), score_labels=5),
    # same input handed over as CSR: the reference densifies (:115-116,:423)
    dict(name="small_csr", n=300, g=2400, extras=True, seed=1013, container="csr", kw=dict(chunksize=128)),
    # two reference categories -> bounded centring (:424-432)
    dict(
        name="small_cats",
        n=300,
        g=2400,
        extras=True,
        obs_cats=3,
        kw=dict(chunksize=128, reference_key="cell_type", reference_cat=["c0", "c2"]),
    ),
    # one reference category -> plain centring with a category mean (:388-400)
    dict(
        name="small_onecat",
        n=200,
        g=2400,
        extras=False,
        obs_cats=3,
        kw=dict(chunksize=5000, reference_key="cell_type", reference_cat="c1"),
    ),
    # window 250: pyramid peak falls inside a step-aligned group of genes
    dict(name="win250", n=120, g=6000, extras=False, kw=dict(window_size=250, chunksize=50)),
    # odd window, step 1 (no decimation)
    dict(name="win11_step1", n=40, g=1500, extras=True, kw=dict(window_size=11, step=1, chunksize=16)),
    # step does not divide the window
    dict(name="win20_step3", n=40, g=1500, extras=False, kw=dict(window_size=20, step=3, chunksize=5000)),
    # step 20 > default, window 100, lfc_clip small enough to bite (:436)
    dict(name="win100_step20_clip", n=64, g=6000, extras=False, kw=dict(step=20, lfc_clip=0.5, chunksize=40)),
    # no noise filter (:449 dynamic_threshold=None)
    dict(name="no_threshold", n=48, g=2400, extras=True, kw=dict(dynamic_threshold=None, chunksize=20)),
    # explicit float64 reference -> the subtraction happens in float64 (:423 dtype promotion)
    dict(name="explicit_ref_f64", n=48, g=2400, extras=True, explicit_ref=11, kw=dict(chunksize=5000)),
    # keep the genosomes (:107 exclude_chromosomes=None)
    dict(name="keep_xy", n=48, g=2400, extras=True, kw=dict(exclude_chromosomes=None, chunksize=5000)),
    # bench-shaped gene axis: G = 20000 with the SURVEY var spec -> K = 1792 / 1463
    dict(name="g20k_win100", n=48, g=20000, extras=False, kw=dict(chunksize=32)),
    dict(name="g20k_win250", n=24, g=20000, extras=False, kw=dict(window_size=250, chunksize=5000)),
    # ---- calculate_gene_values=True (:141-151, :220-223, :247-291, :443-444, :452-453): per-gene layer ----
    # masked genes + flat chromosomes + three chunks
    dict(name="gv_small", n=24, g=2400, extras=True, seed=2001, gene_values=True, kw=dict(chunksize=10)),
    # step does not divide the window; uncovered tail genes -> NaN
    dict(name="gv_win11_step3", n=16, g=1500, extras=False, seed=2002, gene_values=True,
         kw=dict(window_size=11, step=3, chunksize=5000)),
    # step > window: genes between kept windows are not covered at all
    dict(name="gv_win10_step20", n=12, g=1500, extras=True, seed=2003, gene_values=True,
         kw=dict(window_size=10, step=20, chunksize=5)),
    # bounded centring, no noise filter
    dict(name="gv_cats_nothr", n=20, g=2400, extras=True, seed=2004, obs_cats=3, gene_values=True,
         kw=dict(chunksize=8, dynamic_threshold=None, reference_key="cell_type", reference_cat=["c0", "c1"])),
    # bench-shaped gene axis
    dict(name="gv_g20k", n=8, g=20000, extras=False, seed=2005, gene_values=True, kw=dict(chunksize=5)),
    # ---- shapes of the reference's two tutorial notebooks, synthetic stand-ins ----
    # (bounded centring on the direct kernel staged in parts / wide-row median)
    # tutorial_3k: ~58k genes, window 250, step 10, several reference categories -> ~5300 columns
    dict(name="tutorial_3k_like", n=16, g=58000, extras=True, seed=3001, obs_cats=4,
         kw=dict(window_size=250, chunksize=5000, reference_key="cell_type", reference_cat=["c0", "c1", "c3"])),
    # reproduce_infercnv: ~11k genes, window 100, step 1, two reference categories -> ~9000 columns
    dict(name="reproduce_like", n=12, g=11000, extras=False, seed=3002, obs_cats=3,
         kw=dict(window_size=100, step=1, chunksize=5000, reference_key="cell_type", reference_cat=["c0", "c2"])),
]


def build_case(case):
    """-> ``(X float32 dense, var, obs, kwargs for infercnv)``."""
    var = synthetic_var(case["g"], seed=0, with_extras=case.get("extras", False))
    X = synthetic_counts(case["n"], case["g"], seed=case.get("seed", 1000 + len(case["name"])))
    n = case["n"]
    obs = pd.DataFrame(index=pd.Index([f"cell{i}" for i in range(n)]))
    kw = dict(case.get("kw", {}))
    if case.get("obs_cats"):
        rng = np.random.default_rng(99)
        lab = rng.integers(0, case["obs_cats"], size=n)
        obs = obs.assign(cell_type=[f"c{i}" for i in lab])
        # give the categories different depth so the reference rows really differ
        scale = 1.0 + 0.35 * lab[:, None]
        X = (X * scale).astype(np.float32)
    if case.get("explicit_ref") is not None:
        rng = np.random.default_rng(case["explicit_ref"])
        kw["reference"] = rng.uniform(0.0, 0.6, size=case["g"])  # float64 on purpose
    return X, var, obs, kw
