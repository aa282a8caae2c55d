"""Host-side gene-axis preparation (pure pandas/numpy, integers only).

Everything here is metadata work on ``adata.var`` (G <= ~60k rows) that the
reference redoes inside every chunk
(``/root/reference/src/infercnvpy/tl/_infercnv.py:104-108``, ``:327``,
``:350-351``).  We do it once per call, with the reference's own pandas calls
where the result could otherwise differ (tie order of ``sort_values``), and
hand the integer permutation to ``icnv_plan_create``.
"""

from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np
import pandas as pd


def _natural_key(name: str):
    # tl/_infercnv.py:164-176
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split("([0-9]+)", name)]


@dataclass
class GeneLayout:
    n_genes: int                      # columns of the expression matrix
    var_mask: np.ndarray              # bool [G]; True = gene dropped (null / excluded chromosome)
    chromosomes: list[str]            # output order
    gene_idx: np.ndarray              # int32, original column of the p-th position-sorted gene
    seg_off: np.ndarray               # int32 [n_seg + 1]
    window: int
    step: int
    out_off: np.ndarray = field(default=None)   # int64 [n_seg + 1] first output column per chromosome
    n_null: int = 0

    @property
    def n_out(self) -> int:
        return int(self.out_off[-1])

    @property
    def chr_pos(self) -> dict:
        # tl/_infercnv.py:335-337 — values are numpy ints from np.cumsum
        return {c: self.out_off[i] for i, c in enumerate(self.chromosomes)}


def build_layout(var: pd.DataFrame, window_size: int, step: int, exclude_chromosomes=("chrX", "chrY")) -> GeneLayout:
    if {"chromosome", "start", "end"} - set(var.columns) != set():
        # tl/_infercnv.py:99-102
        raise ValueError(
            "Genomic positions not found. There need to be `chromosome`, `start`, and `end` columns in `adata.var`. "
        )
    window_size = int(window_size)
    step = int(step)
    if window_size < 1 or step < 1:
        raise ValueError("window_size and step must be positive integers")
    chrom = var["chromosome"]
    var_mask = chrom.isnull()  # :104
    n_null = int(np.sum(var_mask))
    if exclude_chromosomes is not None:
        var_mask = var_mask | chrom.isin(exclude_chromosomes)  # :107-108
    var_mask = np.asarray(var_mask, dtype=bool)
    kept_cols = np.flatnonzero(~var_mask)

    # the frame the reference works on after `adata[:, ~var_mask]` (:110,:118); positional index
    sub = pd.DataFrame(
        {"chromosome": np.asarray(chrom, dtype=object)[kept_cols], "start": np.asarray(var["start"])[kept_cols]}
    )
    uniq = [c for c in sub["chromosome"].unique() if isinstance(c, str) and c.startswith("chr") and c != "chrM"]  # :327
    chromosomes = sorted(uniq, key=_natural_key)

    pieces, seg_off = [], [0]
    for c in chromosomes:
        # same call as :350 so ties in `start` fall the same way on the same machine
        order = sub.loc[sub["chromosome"] == c].sort_values("start").index.to_numpy()
        pieces.append(kept_cols[order])
        seg_off.append(seg_off[-1] + order.size)
    gene_idx = np.concatenate(pieces).astype(np.int32) if pieces else np.zeros(0, np.int32)
    seg_off = np.asarray(seg_off, dtype=np.int32)

    # window grid: :205 (n < G_c regular, else one flat column), :215-218 decimation
    widths = []
    for i in range(len(chromosomes)):
        g_c = int(seg_off[i + 1] - seg_off[i])
        widths.append((g_c - window_size) // step + 1 if window_size < g_c else 1)
    out_off = np.cumsum([0] + widths).astype(np.int64)
    return GeneLayout(
        n_genes=int(var.shape[0]),
        var_mask=var_mask,
        chromosomes=chromosomes,
        gene_idx=gene_idx,
        seg_off=seg_off,
        window=window_size,
        step=step,
        out_off=out_off,
        n_null=n_null,
    )


def shard_rows(n_rows: int, chunksize: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range of ``rank``; boundaries are multiples of ``chunksize``.

    Each per-chunk standard deviation (tl/_infercnv.py:123,450) then lives on
    exactly one GPU and the partition equals the reference's.
    """
    n_chunks = (n_rows + chunksize - 1) // chunksize
    base, extra = divmod(n_chunks, world)
    c0 = rank * base + min(rank, extra)
    c1 = c0 + base + (1 if rank < extra else 0)
    return min(n_rows, c0 * chunksize), min(n_rows, c1 * chunksize)
