"""CPU oracle for the ``tl.infercnv`` / ``tl.cnv_score`` hot path.

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the reference
algorithm (icbi-lab/infercnvpy @ 89aac1e).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it;
the product package ``infercnvpy_b200`` never does and fails loudly if its CUDA
library is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against (a) the known-answer vectors the reference's own test-suite holds
for this path (``tests/conftest.py:61-108``, ``tests/test_tools.py:11-191``,
``tests/test_scores.py:18-21``) and (b) outputs of the unmodified reference
module run in the build container on seeded inputs
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

All ``file:line`` citations are into ``/root/reference/src/infercnvpy/``.
The oracle works on plain arrays (matrix, chromosome labels, start positions)
rather than on AnnData, which is not installed in this image.
"""

from __future__ import annotations

import re
from concurrent.futures import ProcessPoolExecutor
from typing import Sequence

import numpy as np
import pandas as pd
import scipy.sparse as sp

__all__ = [
    "natural_chromosome_order",
    "gene_order",
    "reference_profile",
    "pyramid_weights",
    "running_mean",
    "smooth_by_chromosome",
    "infercnv_chunk",
    "infercnv",
    "cnv_score",
    "ith_score",
]


# --------------------------------------------------------------------------- #
# chromosome bookkeeping
# --------------------------------------------------------------------------- #
def _natural_key(name: str):
    """tl/_infercnv.py:164-176 — digits compare as ints, the rest lower-cased."""
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split("([0-9]+)", name)]


def natural_chromosome_order(chromosomes: Sequence) -> list[str]:
    """Chromosomes that take part in the smoothing, in output order.

    tl/_infercnv.py:327 — keep the unique labels that start with ``"chr"`` and
    are not ``"chrM"``; natural-sort them (chr1, chr2, ..., chr10, ...).
    Labels that survive the exclusion mask but do not start with ``chr`` are
    silently never used.
    """
    uniq = pd.unique(pd.Series(chromosomes, dtype=object))
    keep = [c for c in uniq if isinstance(c, str) and c.startswith("chr") and c != "chrM"]
    return sorted(keep, key=_natural_key)


def gene_order(chromosome: np.ndarray, start: np.ndarray, which: str) -> np.ndarray:
    """Column indices of chromosome ``which`` ordered by ``start``.

    tl/_infercnv.py:350-351 — the reference does
    ``var.loc[var.chromosome == chr].sort_values("start").index`` followed by
    ``var.index.get_indexer``.  ``sort_values`` defaults to numpy's
    (non-stable) quicksort, so ties in ``start`` come out in whatever order
    that algorithm leaves them; we make the identical pandas call so the
    integer permutation is the same on the same machine.
    """
    frame = pd.DataFrame({"start": np.asarray(start)})
    member = np.asarray(pd.Series(chromosome, dtype=object) == which)
    return frame.loc[member].sort_values("start").index.to_numpy()


# --------------------------------------------------------------------------- #
# reference profile
# --------------------------------------------------------------------------- #
def _as_ndarray(a):
    """_util.py:4-9 — ``np.matrix`` (from sparse arithmetic) -> ndarray."""
    return a.A if isinstance(a, np.matrix) else a


def reference_profile(X, obs_column=None, reference_cat=None, reference=None) -> np.ndarray:
    """``[n_cat, G]`` reference expression.  tl/_infercnv.py:359-408.

    * explicit ``reference`` wins (``:379``);
    * without a key/category: column mean over all cells (``:380-385``) — for a
      float32 matrix this is numpy's float32 mean;
    * otherwise one mean per category over the cells whose ``obs_column``
      equals it (``:388-400``); unknown categories raise ``ValueError``
      (``:393-398``).
    A 1-D result is promoted to ``[1, G]`` (``:402-403``) and the width is
    checked against the matrix (``:405-406``).
    """
    if reference is None:
        if obs_column is None or reference_cat is None:
            reference = np.mean(X, axis=0)
        else:
            col = np.asarray(obs_column)
            cats = np.array([reference_cat] if isinstance(reference_cat, str) else list(reference_cat))
            present = np.isin(cats, col)
            if not present.all():
                raise ValueError(
                    f"The following reference categories were not found in adata.obs[reference_key]: {cats[~present]}"
                )
            reference = np.vstack([np.mean(X[col == c, :], axis=0) for c in cats])
    if reference.ndim == 1:
        reference = reference[np.newaxis, :]
    if reference.shape[1] != X.shape[1]:
        raise ValueError("Reference must match the number of genes in AnnData. ")
    return reference


# --------------------------------------------------------------------------- #
# smoothing
# --------------------------------------------------------------------------- #
def pyramid_weights(n: int) -> np.ndarray:
    """tl/_infercnv.py:206-207 — ``min(r, reversed r)`` for r = 1..n (int64).

    n=100 -> 1..50,50..1 (sum 2550); n=5 -> 1,2,3,2,1.
    """
    ramp = np.arange(1, n + 1)
    return np.minimum(ramp, ramp[::-1])


def running_mean(x: np.ndarray, n: int, step: int) -> np.ndarray:
    """Pyramid-weighted running mean along axis 1.  tl/_infercnv.py:179-244.

    * ``n < width``: ``np.convolve(row, pyramid, "valid") / pyramid.sum()`` per
      row (``:205-212``) — the int64 kernel promotes float32 rows to float64 —
      then keep every ``step``-th window (``:215-218``);
    * ``n >= width`` (``:227-236``): ONE column, the flat (unweighted) mean of
      all genes of the segment.
    """
    x = np.asarray(x)
    width = x.shape[1]
    if n < width:
        w = pyramid_weights(n)
        full = np.empty((x.shape[0], width - n + 1), dtype=np.result_type(x.dtype, w.dtype, np.float64))
        for i in range(x.shape[0]):
            full[i] = np.convolve(x[i], w, mode="valid")
        full /= np.sum(w)
        return full[:, np.arange(0, full.shape[1], step)]
    flat = np.ones(width, dtype=np.int64)
    out = np.empty((x.shape[0], 1), dtype=np.result_type(x.dtype, flat.dtype, np.float64))
    for i in range(x.shape[0]):
        out[i] = np.convolve(x[i], flat, mode="valid")
    return out / np.sum(flat)


def gene_values_for_segment(smoothed: np.ndarray, width: int, n: int, step: int):
    """Per-gene values of one chromosome.  tl/_infercnv.py:214-223, 238-242, 247-291.

    Regular branch: the value of the gene at sorted position ``p`` is ``np.mean`` of the decimated
    windows ``k`` that contain it (``k*step <= p < k*step + n``), in window order
    (``_calculate_gene_averages`` appends them left to right and takes ``np.mean`` of the list,
    ``:278-287``); genes no kept window covers do not appear at all.  Flat branch
    (``width <= n``): every gene gets the single flat mean (``:240``).
    Returns ``(positions, values [rows, len(positions)])``.
    """
    rows = smoothed.shape[0]
    if n < width:
        n_out = smoothed.shape[1]
        pos, cols = [], []
        for p in range(width):
            k_lo = max(0, -((n - 1 - p) // step))  # ceil((p - n + 1) / step)
            k_hi = min(n_out - 1, p // step)
            if k_lo > k_hi:
                continue
            pos.append(p)
            # np.mean of a 1-D float64 array: add.reduce (pairwise) then divide by the count
            block = np.ascontiguousarray(smoothed[:, k_lo : k_hi + 1], dtype=np.float64)
            cols.append(np.array([np.mean(block[r]) for r in range(rows)]))
        vals = np.stack(cols, axis=1) if cols else np.empty((rows, 0))
        return np.asarray(pos, dtype=np.int64), vals
    return np.arange(width, dtype=np.int64), np.repeat(smoothed, width, axis=1)


def smooth_by_chromosome(x: np.ndarray, chromosome, start, window: int, step: int, gene_values: bool = False):
    """tl/_infercnv.py:301-343 — smooth every chromosome on its own and stack.

    Returns ``(chr_pos, smoothed)`` where ``chr_pos[chr]`` is the first output
    column of that chromosome (``:335-337``, numpy ints from ``np.cumsum``).
    """
    order = natural_chromosome_order(chromosome)
    genes = [gene_order(chromosome, start, c) for c in order]
    pieces = [running_mean(x[:, g], window, step) for g in genes]
    offsets = np.cumsum([0] + [p.shape[1] for p in pieces])
    chr_pos = {c: off for c, off in zip(order, offsets)}
    if not gene_values:
        return chr_pos, np.hstack(pieces)
    # per-gene values, concatenated over chromosomes (:339-341); ``gene_cols`` are the matrix columns they belong to
    gene_cols, gene_vals = [], []
    for g, piece in zip(genes, pieces):
        pos, vals = gene_values_for_segment(piece, len(g), window, step)
        gene_cols.append(np.asarray(g)[pos])
        gene_vals.append(vals)
    return chr_pos, np.hstack(pieces), np.concatenate(gene_cols), np.hstack(gene_vals)


def infercnv_chunk(x, chromosome, start, reference, lfc_clip, window, step, dynamic_threshold, gene_values=False):
    """One row-chunk of the method.  tl/_infercnv.py:411-457.

    1. centre: one reference row -> ``x - ref`` (``:422-423``); several ->
       "bounded" difference: 0 inside ``[min, max]`` of the reference rows,
       distance to the nearer bound outside (``:424-432``; the result buffer
       takes ``x.dtype``, ``:428``);
    2. clip to ``+-lfc_clip`` (``:436``);
    3. per-chromosome pyramid smoothing (``:438``);
    4. subtract the per-row median (``:442``);
    5. zero entries with ``|v| < dynamic_threshold * np.std(chunk)``
       (population std over *every element of the chunk*, ``:449-451``);
    6. to CSR (``:455``).
    """
    reference = np.asarray(reference) if not isinstance(reference, np.matrix) else reference
    if reference.shape[0] == 1:
        centred = x - reference[0, :]
    else:
        lo = np.min(reference, axis=0)
        hi = np.max(reference, axis=0)
        centred = np.zeros(x.shape, dtype=x.dtype)
        above = _as_ndarray(x > hi)
        below = _as_ndarray(x < lo)
        centred[above] = _as_ndarray(x - hi)[above]
        centred[below] = _as_ndarray(x - lo)[below]
    centred = np.asarray(_as_ndarray(centred))
    clipped = np.clip(centred, -lfc_clip, lfc_clip)
    if gene_values:
        chr_pos, smoothed, gene_cols, conv = smooth_by_chromosome(clipped, chromosome, start, window, step, True)
    else:
        chr_pos, smoothed = smooth_by_chromosome(clipped, chromosome, start, window, step)
    res = smoothed - np.median(smoothed, axis=1)[:, np.newaxis]
    gene_res = None
    if gene_values:
        # the per-gene layer is centred on ITS OWN row median (:444) but filtered with the window threshold (:453)
        gene_res = conv - np.median(conv, axis=1)[:, np.newaxis]
    if dynamic_threshold is not None:
        thr = dynamic_threshold * np.std(res)
        res[np.abs(res) < thr] = 0
        if gene_values:
            gene_res[np.abs(gene_res) < thr] = 0
    if gene_values:
        return chr_pos, sp.csr_matrix(res), (gene_cols, gene_res)
    return chr_pos, sp.csr_matrix(res)


def _chunk_job(args):
    return infercnv_chunk(*args)


def infercnv(
    X,
    chromosome,
    start,
    *,
    obs_column=None,
    reference_cat=None,
    reference=None,
    lfc_clip=3,
    window_size=100,
    step=10,
    dynamic_threshold=1.5,
    exclude_chromosomes=("chrX", "chrY"),
    chunksize=5000,
    n_jobs=1,
    calculate_gene_values=False,
):
    """Whole-matrix driver.  tl/_infercnv.py:97-161.

    Gene mask = null chromosome or chromosome in ``exclude_chromosomes``
    (``:104-108``); the reference profile is computed on ALL genes and then
    masked (``:111``); rows are cut into ``chunksize`` blocks that are processed
    independently — each has its own std (``:120-135``) — and stacked
    (``:137``); ``chr_pos`` is taken from the first chunk (``:139``).
    ``n_jobs`` > 1 fans the chunks out to a process pool like the reference's
    ``process_map`` (``:121,132``).
    Returns ``(chr_pos, csr float64 [N, K])``; with ``calculate_gene_values`` also the per-gene matrix
    ``[N, n_var]`` float64, NaN for genes no kept window covers and for masked genes (``:141-148``).
    """
    chrom_s = pd.Series(np.asarray(chromosome, dtype=object))
    drop = chrom_s.isnull()
    if exclude_chromosomes is not None:
        drop = drop | chrom_s.isin(exclude_chromosomes)
    keep = ~drop.to_numpy()

    ref = reference_profile(X, obs_column, reference_cat, reference)[:, keep]
    expr = X[:, keep]
    if sp.issparse(expr):
        expr = expr.tocsr()
    chrom_k = chrom_s.to_numpy()[keep]
    start_k = np.asarray(start)[keep]

    jobs = [
        (expr[i : i + chunksize, :], chrom_k, start_k, ref, lfc_clip, window_size, step, dynamic_threshold,
         calculate_gene_values)
        for i in range(0, X.shape[0], chunksize)
    ]
    if n_jobs is None or n_jobs > 1:
        with ProcessPoolExecutor(max_workers=n_jobs) as pool:
            results = list(pool.map(_chunk_job, jobs))
    else:
        results = [_chunk_job(j) for j in jobs]
    chr_pos = results[0][0]
    res = sp.vstack([r[1] for r in results])
    if not calculate_gene_values:
        return chr_pos, res
    kept_to_var = np.flatnonzero(keep)
    per_gene = np.full((X.shape[0], X.shape[1]), np.nan)
    r0 = 0
    for _, blk, (cols, vals) in results:
        per_gene[r0 : r0 + blk.shape[0], kept_to_var[cols]] = vals
        r0 += blk.shape[0]
    return chr_pos, res, per_gene


# --------------------------------------------------------------------------- #
# cnv_score
# --------------------------------------------------------------------------- #
def cnv_score(X_cnv, labels) -> dict:
    """tl/_scores.py:65-68 — per label ``mean(abs(X_cnv[rows of label, :]))``.

    The mean runs over ALL entries of the row block (zeros included), for
    dense or sparse ``X_cnv``.
    """
    labels = np.asarray(labels)
    return {lab: np.mean(np.abs(X_cnv[labels == lab, :])) for lab in pd.unique(pd.Series(labels))}


# --------------------------------------------------------------------------- #
# ITH scores
# --------------------------------------------------------------------------- #
def ith_score(X, labels) -> dict:
    """tl/_scores.py:128-141 (ithgex) / :201-214 (ithcna) — per label the inter-quartile range of the cell-cell Pearson
    correlation matrix of its rows (``np.corrcoef(X, rowvar=True)``, all entries, ``np.percentile(.., [75, 25])``).

    ``X`` is the matrix the score is taken of (expression for ithgex, ``obsm["X_cnv"]`` for ithcna), dense or sparse;
    groups with a single cell are skipped (``:135,208``).
    """
    labels = np.asarray(labels)
    out = {}
    for lab in pd.unique(pd.Series(labels)):
        block = X[labels == lab, :]
        if sp.issparse(block):
            block = block.todense()
        if block.shape[0] <= 1:
            continue
        pcorr = np.corrcoef(block, rowvar=True)
        q75, q25 = np.percentile(pcorr, [75, 25])
        out[lab] = q75 - q25
    return out
