"""``cnv.tl.infercnv`` — same signature, keys and errors as the reference
(``/root/reference/src/infercnvpy/tl/_infercnv.py:18-161``); the arithmetic runs in
``libicnv.so`` on the current CUDA device.  No CPU fallback.
"""

from __future__ import annotations

import logging
import os
from collections.abc import Sequence

import numpy as np
import scipy.sparse

from .. import _lib
from .._engine import DevicePlan, allreduce_host_counts, allreduce_sums
from .._layout import build_layout

log = logging.getLogger("infercnvpy_b200")


def _block_rows(n_rows: int, n_genes: int, n_out: int, chunksize: int, gene_values: bool = False) -> int:
    """Rows per device block: a multiple of ``chunksize`` (so every per-chunk std sees a whole
    chunk, _infercnv.py:123,450) that keeps input + intermediates + output under ICNV_BLOCK_BYTES (default: 60 % of
    the free device memory, at least 4 GiB — a matrix that fits stays resident and crosses PCIe once)."""
    if "ICNV_BLOCK_BYTES" in os.environ:
        budget = int(os.environ["ICNV_BLOCK_BYTES"])
    else:
        import torch

        budget = max(4 << 30, int(0.6 * torch.cuda.mem_get_info()[0]))
    per_row = 4 * n_genes + 14 * n_out + 64 + (8 * n_genes + 10 * n_out if gene_values else 0)
    chunks = max(1, (budget // per_row) // chunksize)
    return min(n_rows, chunks * chunksize) if n_rows else 0


def _rows_to_device(expr, r0, r1, device):
    """Rows [r0, r1) of the host matrix as device float32 (dense tensor or CSR triple)."""
    import torch

    from ..io import DeviceCSR

    if isinstance(expr, DeviceCSR):  # already in HBM (infercnvpy_b200.io.read_matrix): canonical CSR, views only
        return expr.rows(r0, r1)
    if scipy.sparse.issparse(expr):
        whole = r0 == 0 and r1 == expr.shape[0] and expr.format == "csr"
        blk = expr if whole else expr[r0:r1].tocsr()
        triple = (
            _upload_1d(blk.indptr.astype(np.int64, copy=False), device),
            _upload_1d(blk.indices.astype(np.int32, copy=False), device),
            _upload_1d(blk.data.astype(np.float32, copy=False), device),
        )
        if not _csr_is_canonical(*triple):
            # duplicates / unsorted columns: scipy's toarray() sums duplicates (_infercnv.py:423); canonicalise on
            # the host (rare) and upload again
            blk = blk.copy()
            blk.sum_duplicates()
            triple = (
                _upload_1d(blk.indptr.astype(np.int64, copy=False), device),
                _upload_1d(blk.indices.astype(np.int32, copy=False), device),
                _upload_1d(blk.data.astype(np.float32, copy=False), device),
            )
        return triple
    if isinstance(expr, torch.Tensor):
        return expr[r0:r1].to(device=device, dtype=torch.float32).contiguous()
    blk = np.ascontiguousarray(np.asarray(expr[r0:r1]), dtype=np.float32)
    LAST_TRANSFER["h2d_bytes"] += blk.nbytes
    return torch.from_numpy(blk).to(device)


def _upload_1d(arr: np.ndarray, device):
    """Host 1-D array -> device tensor.  Pinned sources go as one async DMA; large pageable ones are staged through
    two pinned buffers so the host memcpy of slab i+1 overlaps the DMA of slab i."""
    import torch

    arr = np.ascontiguousarray(arr)
    src = torch.from_numpy(arr)
    n = src.numel()
    LAST_TRANSFER["h2d_bytes"] += arr.nbytes
    if n * src.element_size() < (64 << 20):
        return src.to(device)
    dst = torch.empty((n,), dtype=src.dtype, device=device)
    if src.is_pinned():
        dst.copy_(src, non_blocking=True)
        return dst
    slab = (128 << 20) // src.element_size()
    staging = [torch.empty((slab,), dtype=src.dtype, pin_memory=True) for _ in range(2)]
    done = [None, None]
    for i, a in enumerate(range(0, n, slab)):
        b = min(n, a + slab)
        buf = staging[i % 2]
        if done[i % 2] is not None:
            done[i % 2].synchronize()
        buf[: b - a].copy_(src[a:b])
        dst[a:b].copy_(buf[: b - a], non_blocking=True)
        done[i % 2] = torch.cuda.Event()
        done[i % 2].record()
    for ev in done:
        if ev is not None:
            ev.synchronize()
    return dst


def _csr_is_canonical(indptr, indices, data) -> bool:
    """Column indices strictly increasing inside every row (checked on the device: one pass over ``indices``)."""
    import torch

    nnz = indices.numel()
    if nnz < 2:
        return True
    bad = indices[1:] <= indices[:-1]
    starts = indptr[1:-1]
    starts = starts[(starts > 0) & (starts < nnz)]
    bad[starts - 1] = False  # a new row may start with any column
    return not bool(bad.any().item())


def _upload_pipelined(expr: np.ndarray, device, on_slab):
    """Host float32 C-contiguous matrix -> resident device tensor, copied in ~256 MB row slabs on a side stream.
    ``on_slab(Xd[r0:r1], r0, r1)`` is enqueued on the compute stream as soon as a slab has landed, so the
    column-sum pass overlaps the PCIe transfer.  Pageable input goes through two pinned staging buffers."""
    import torch

    n, G = expr.shape
    LAST_TRANSFER["h2d_bytes"] += expr.nbytes
    Xd = torch.empty((n, G), dtype=torch.float32, device=device)
    src = torch.from_numpy(expr)
    pinned = src.is_pinned()
    slab = max(1, (256 << 20) // (4 * G))
    main = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device)
    side.wait_stream(main)
    staging, free_ev = None, None
    if not pinned:
        staging = [torch.empty((min(slab, n), G), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        free_ev = [None, None]
    landed = []
    for i, r0 in enumerate(range(0, n, slab)):
        r1 = min(n, r0 + slab)
        with torch.cuda.stream(side):
            if pinned:
                Xd[r0:r1].copy_(src[r0:r1], non_blocking=True)
            else:
                buf = staging[i % 2]
                if free_ev[i % 2] is not None:
                    free_ev[i % 2].synchronize()
                buf[: r1 - r0].copy_(src[r0:r1])
                Xd[r0:r1].copy_(buf[: r1 - r0], non_blocking=True)
                free_ev[i % 2] = torch.cuda.Event()
                free_ev[i % 2].record(side)
            ev = torch.cuda.Event()
            ev.record(side)
        landed.append((ev, r0, r1))
        if not pinned:  # staging keeps the host busy anyway: consume as we go
            main.wait_event(ev)
            on_slab(Xd[r0:r1], r0, r1)
    if pinned:  # every DMA is already queued; now hang the per-slab work behind the events
        for ev, r0, r1 in landed:
            main.wait_event(ev)
            on_slab(Xd[r0:r1], r0, r1)
    return Xd


_PINNED: dict = {}
_POOL: dict = {}
LAST_TRANSFER = {"h2d_bytes": 0, "d2h_bytes": 0}  # bytes the last infercnv() call moved over PCIe (bench.py reports them)


def _pool():
    """Host copy threads of this process (created once)."""
    from concurrent.futures import ThreadPoolExecutor

    n = _copy_threads()
    if _POOL.get("n") != n:
        _POOL["n"] = n
        _POOL["pool"] = ThreadPoolExecutor(max_workers=n)
    return _POOL["pool"], n


def _to_host(t, out_dtype=None, slab_bytes: int = 32 << 20):
    """Device tensor -> fresh numpy array (optionally widened to ``out_dtype`` on the host, so float32 values cross PCIe
    as 4 bytes and land as the reference's float64).  The tensor is copied in slabs through two cached pinned buffers:
    the DMA of slab i+1 overlaps the host threads that move slab i into the (first-touched) result array."""
    import torch

    n = t.numel()
    src_np_dtype = np.dtype(str(t.dtype).replace("torch.", ""))
    out_dtype = np.dtype(out_dtype or src_np_dtype)
    if n < (1 << 16):
        return t.cpu().numpy().astype(out_dtype, copy=False)
    LAST_TRANSFER["d2h_bytes"] += n * t.element_size()
    flat = t.reshape(-1)
    per = max(1, slab_bytes // t.element_size())
    key = (t.dtype, per)
    bufs = _PINNED.get(key)
    if bufs is None:
        bufs = [torch.empty((per,), dtype=t.dtype, pin_memory=True) for _ in range(2)]
        _PINNED[key] = bufs
    out = np.empty(n, dtype=out_dtype)
    pool, n_thr = _pool()
    spans = [(a, min(n, a + per)) for a in range(0, n, per)]
    evs = [None, None]
    main = torch.cuda.current_stream(t.device)
    skey = ("stream", str(t.device))
    if skey not in _PINNED:  # one copy stream per device, created once
        _PINNED[skey] = torch.cuda.Stream(t.device)
    side = _PINNED[skey]
    side.wait_stream(main)

    def issue(i):
        a, b = spans[i]
        with torch.cuda.stream(side):
            bufs[i % 2][: b - a].copy_(flat[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        evs[i % 2] = ev

    def move(i):
        a, b = spans[i]
        src = bufs[i % 2][: b - a].numpy()
        step = max(1 << 18, -(-(b - a) // n_thr))
        parts = [(x, min(b - a, x + step)) for x in range(0, b - a, step)]
        if len(parts) > 1:
            list(pool.map(lambda xy: np.copyto(out[a + xy[0] : a + xy[1]], src[xy[0] : xy[1]], casting="same_kind"), parts))
        else:
            np.copyto(out[a:b], src, casting="same_kind")

    issue(0)
    for i in range(len(spans)):
        evs[i % 2].synchronize()
        if i + 1 < len(spans):
            issue(i + 1)  # the other buffer: its previous content was moved out in the last iteration
        move(i)
    main.wait_stream(side)  # the source tensor may be freed / reused by the caller from here on
    return out.reshape(t.shape)


def _copy_threads() -> int:
    """Host copy threads of this rank: at most 16, and no more than its share of the cores when several ranks (LOCAL_WORLD_SIZE
    / WORLD_SIZE of torchrun) run on the box (first-touching the result arrays is page-fault bound, ~2-4 GB/s per thread)."""
    if os.environ.get("ICNV_COPY_THREADS"):  # developer override
        return max(1, int(os.environ["ICNV_COPY_THREADS"]))
    ranks = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, min(16, (os.cpu_count() or 8) // max(1, ranks)))


def _to_host_into(t, dst: np.ndarray, slab_bytes: int = 256 << 20):
    """Device matrix -> rows of a preallocated C-contiguous numpy array, in row slabs through two pinned buffers
    (the DMA of slab i+1 overlaps the host copy of slab i)."""
    import torch

    n, w = t.shape
    assert dst.shape == (n, w) and dst.flags.c_contiguous and dst.dtype == np.dtype(str(t.dtype).replace("torch.", ""))
    if n == 0:
        return
    rows = max(1, slab_bytes // max(1, w * t.element_size()))
    bufs = [torch.empty((min(rows, n), w), dtype=t.dtype, pin_memory=True) for _ in range(2 if n > rows else 1)]
    evs = [None] * len(bufs)
    spans = [(a, min(n, a + rows)) for a in range(0, n, rows)]

    def issue(i):
        a, b = spans[i]
        bufs[i % len(bufs)][: b - a].copy_(t[a:b], non_blocking=True)
        evs[i % len(bufs)] = torch.cuda.Event()
        evs[i % len(bufs)].record()

    issue(0)
    for i, (a, b) in enumerate(spans):
        evs[i % len(bufs)].synchronize()
        if i + 1 < len(spans) and len(bufs) > 1:
            issue(i + 1)
        np.copyto(dst[a:b], bufs[i % len(bufs)][: b - a].numpy())
        if i + 1 < len(spans) and len(bufs) == 1:
            issue(i + 1)


def _host_csr(indptr, indices, data, shape):
    """Device CSR pieces -> scipy CSR float64 (the reference's container, _infercnv.py:455)."""
    import torch

    small = indices.numel() < 2**31 - 1
    ip = _to_host(indptr.to(torch.int32) if small else indptr)
    ix = _to_host(indices if small else indices.to(torch.int64))
    dv = _to_host(data, np.float64)  # float32 on the wire, widened by the host copy threads
    res = scipy.sparse.csr_matrix((dv, ix, ip), shape=shape, copy=False)
    res.has_sorted_indices = True  # compacted in column order
    return res


_PLAN_CACHE: dict = {}


def _cached_plan(var, window_size, step, exclude_chromosomes, device):
    """(GeneLayout, DevicePlan) for this gene axis; both only depend on var/window/step, so repeated calls on
    the same AnnData (or on row shards of it) reuse the tables already in HBM."""
    import hashlib

    h = hashlib.sha1()
    import pandas as pd

    h.update(pd.util.hash_pandas_object(var["chromosome"], index=False).to_numpy().tobytes())  # vectorised, 64 bits per gene
    h.update(np.ascontiguousarray(np.asarray(var["start"], dtype=np.float64)).tobytes())
    key = (h.hexdigest(), int(window_size), int(step), None if exclude_chromosomes is None else tuple(exclude_chromosomes), str(device))
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        layout = build_layout(var, window_size, step, exclude_chromosomes)
        hit = (layout, DevicePlan(layout, device))
        if len(_PLAN_CACHE) >= 4:
            old = _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))  # oldest entry first
            old[1].close()
        _PLAN_CACHE[key] = hit
    return hit


def _reference_categories(adata, reference_key, reference_cat):
    """-> (row_cat int32 [n] with -1 for non-reference cells, categories).  _infercnv.py:388-398.

    A category listed twice gives two identical reference rows in the reference, i.e. the same bounds as listing it
    once: duplicates are dropped.  Under an initialised ``torch.distributed`` group the "not found" check runs on the
    cell counts summed over all ranks (a shard without any cell of a rare category is fine) and raises on every rank."""
    obs_col = adata.obs[reference_key]
    if isinstance(reference_cat, str):
        reference_cat = [reference_cat]
    reference_cat = np.array(list(dict.fromkeys(reference_cat)))
    values = np.asarray(obs_col.values)
    row_cat = np.full(values.shape[0], -1, dtype=np.int32)
    local = np.zeros(len(reference_cat), dtype=np.int64)
    for i, cat in enumerate(reference_cat):
        m = values == cat
        row_cat[m] = i
        local[i] = int(np.count_nonzero(m))
    present = allreduce_host_counts(local) > 0
    if not np.all(present):
        raise ValueError(
            "The following reference categories were not found in "
            "adata.obs[reference_key]: "
            f"{reference_cat[~present]}"
        )
    return row_cat, reference_cat


def infercnv(
    adata,
    *,
    reference_key: str | None = None,
    reference_cat: None | str | Sequence[str] = None,
    reference: np.ndarray | None = None,
    lfc_clip: float = 3,
    window_size: int = 100,
    step: int = 10,
    dynamic_threshold: float | None = 1.5,
    exclude_chromosomes: Sequence[str] | None = ("chrX", "chrY"),
    chunksize: int = 5000,
    n_jobs: int | None = None,
    inplace: bool = True,
    layer: str | None = None,
    key_added: str = "cnv",
    calculate_gene_values: bool = False,
):
    """Infer copy number variation by averaging expression over genomic windows (GPU).

    Parameters, return value, AnnData keys and raised errors follow the reference
    (``_infercnv.py:18-161``).  Differences, all documented in DESIGN.md:

    * ``n_jobs`` is accepted and ignored (the reference forks CPU workers, ``:132``);
      ``chunksize`` keeps its meaning for the noise filter — every block of ``chunksize`` cells has
      its own standard deviation (``:123,450``);
    * the matrix is processed as float32 (other dtypes are cast); values of ``X_cnv`` are computed
      with float64 accumulation and stored as float32-rounded float64 CSR;
    * when the reference profile is derived from the data it is a float64-accumulated mean (numpy's
      float32 running sum drifts by ~1e-5 relative at 1e5 cells);
    * under an initialised ``torch.distributed`` group every rank passes its own row shard (cut at
      multiples of ``chunksize``, ``infercnvpy_b200.shard_rows``) and the reference profile is the
      mean over all ranks (one all-reduce).
    """
    import torch

    if not adata.var_names.is_unique:
        raise ValueError("Ensure your var_names are unique!")
    if {"chromosome", "start", "end"} - set(adata.var.columns) != set():
        raise ValueError(
            "Genomic positions not found. There need to be `chromosome`, `start`, and `end` columns in `adata.var`. "
        )
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if device is None:
        raise _lib.IcnvError("infercnvpy_b200.tl.infercnv needs a CUDA device; there is no CPU fallback")
    layout, plan = _cached_plan(adata.var, window_size, step, exclude_chromosomes, device)
    LAST_TRANSFER["h2d_bytes"] = LAST_TRANSFER["d2h_bytes"] = 0
    if layout.n_null:
        log.warning(f"Skipped {layout.n_null} genes because they don't have a genomic position annotated. ")
    chunksize = int(chunksize)
    if chunksize < 1:
        raise ValueError("chunksize must be positive")

    expr = adata.X if layer is None else adata.layers[layer]
    n_rows, n_genes = adata.shape
    src_dtype = np.dtype(str(expr.dtype).replace("torch.", "")) if not hasattr(expr.dtype, "kind") else expr.dtype

    # ---- reference profile: explicit one is validated before any GPU work (_infercnv.py:402-406)
    ref_host = None
    row_cat_host, n_cat = None, 1
    if reference is not None:
        ref_host = np.asarray(reference)
        if ref_host.ndim == 1:
            ref_host = ref_host[np.newaxis, :]
        if ref_host.shape[1] != n_genes:
            raise ValueError("Reference must match the number of genes in AnnData. ")
    elif reference_key is None or reference_cat is None:
        log.warning(
            "Using mean of all cells as reference. For better results, "
            "provide either `reference`, or both `reference_key` and `reference_cat`. "
        )
    else:
        row_cat_host, cats = _reference_categories(adata, reference_key, reference_cat)
        n_cat = len(cats)

    K = plan.K
    block = _block_rows(n_rows, n_genes, K, chunksize, calculate_gene_values)
    blocks = [(r0, min(n_rows, r0 + block)) for r0 in range(0, n_rows, max(block, 1))]
    resident = None
    need_sums = ref_host is None
    sums = counts = None
    fast_host = (
        len(blocks) == 1
        and isinstance(expr, np.ndarray)
        and expr.dtype == np.float32
        and expr.flags.c_contiguous
        and n_rows > 0
    )
    if fast_host:
        # one resident copy; the reference-profile pass rides on the transfer
        acc = {"s": None, "c": None}

        def on_slab(Xs, r0, r1):
            if not need_sums:
                return
            rc = torch.from_numpy(row_cat_host[r0:r1]).to(device) if row_cat_host is not None else None
            s_, c_ = plan.colsum(Xs, rc, n_cat)
            acc["s"] = s_ if acc["s"] is None else acc["s"].add_(s_)
            acc["c"] = c_ if acc["c"] is None else acc["c"].add_(c_)

        resident = _upload_pipelined(expr, device, on_slab)
        sums, counts = acc["s"], acc["c"]
    elif len(blocks) == 1:
        resident = _rows_to_device(expr, 0, n_rows, device)

    # ---- reference profile on the device
    if ref_host is not None:
        c64 = np.result_type(src_dtype, ref_host.dtype) == np.float64
        ref_dev = torch.from_numpy(np.ascontiguousarray(ref_host, dtype=np.float64 if c64 else np.float32)).to(device)
    else:
        row_cat_dev = None
        if sums is None:
            for r0, r1 in blocks:
                Xb = resident if resident is not None else _rows_to_device(expr, r0, r1, device)
                if row_cat_host is not None:
                    row_cat_dev = torch.from_numpy(row_cat_host[r0:r1]).to(device)
                s, c = plan.colsum(Xb, row_cat_dev, n_cat)
                sums = s if sums is None else sums.add_(s)
                counts = c if counts is None else counts.add_(c)
        if sums is None:  # no rows on this rank
            sums = torch.zeros((n_cat, n_genes), dtype=torch.float64, device=device)
            counts = torch.zeros((n_cat,), dtype=torch.int64, device=device)
        sums, counts = allreduce_sums(sums, counts)
        # numpy: float32 matrix -> float32 mean, anything else -> float64 (_infercnv.py:385,400)
        ref_dev = plan.mean_from_sums(sums, counts, f64=(src_dtype != np.float32))
    plan.set_reference(ref_dev)

    # ---- smoothing, noise filter, CSR (blocks are multiples of chunksize)
    parts = []
    # per-gene layer (_infercnv.py:141-148): dense float64 [n_obs, n_vars], NaN for genes without a value
    per_gene = np.empty((n_rows, n_genes), dtype=np.float64) if calculate_gene_values else None
    for r0, r1 in blocks:
        Xb = resident if resident is not None else _rows_to_device(expr, r0, r1, device)
        if isinstance(Xb, tuple) and plan.tier == 2:
            Xb = _densify(Xb, n_genes, device)
        tmp = plan.smooth(Xb, lfc_clip)
        out, stats = plan.center(tmp)
        # noise filter + CSR in two read-only passes over `out` (count, compact); the dense block is never rewritten
        thr, _, _, (indptr, indices, data) = plan.filter_to_csr(out, stats, chunksize, dynamic_threshold)
        if calculate_gene_values:
            # own row median (:444), the window matrix's chunk thresholds (:453); copied out in row slabs
            gv = plan.gene_values(tmp, chunksize, thr)
            _to_host_into(gv, per_gene[r0:r1])
            del gv
        del tmp
        parts.append(_host_csr(indptr, indices, data, (r1 - r0, K)))
        del out, stats, Xb
    if parts:
        res = scipy.sparse.vstack(parts, format="csr") if len(parts) > 1 else parts[0]
    else:
        res = scipy.sparse.csr_matrix((0, K), dtype=np.float64)

    chr_pos = layout.chr_pos
    if inplace:
        adata.obsm[f"X_{key_added}"] = res
        adata.uns[key_added] = {"chr_pos": chr_pos}
        if calculate_gene_values:
            adata.layers[f"gene_values_{key_added}"] = per_gene
    else:
        return chr_pos, res, per_gene


def _densify(csr_triple, n_genes, device):
    """CSR block -> dense float32 block for the direct-form kernel (rare (window, step) pairs)."""
    import torch

    indptr, indices, data = csr_triple
    n = indptr.numel() - 1
    dense = torch.zeros((n, n_genes), dtype=torch.float32, device=device)
    rows = torch.repeat_interleave(torch.arange(n, device=device), (indptr[1:] - indptr[:-1]))
    dense[rows, indices.long()] = data
    return dense
