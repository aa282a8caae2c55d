"""Device pipeline: thin Python over the C ABI.  torch is only the allocator /
stream provider / collective plumbing here; every arithmetic step is a kernel
of ``libicnv.so``.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._layout import GeneLayout


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.IcnvError("infercnvpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


class DevicePlan:
    """Owns an ``icnv_plan*`` (gene-axis tables in HBM) for one layout on one device."""

    def __init__(self, layout: GeneLayout, device=None):
        torch = _torch()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.layout = layout
        handle = C.c_void_p()
        gi = np.ascontiguousarray(layout.gene_idx, dtype=np.int32)
        so = np.ascontiguousarray(layout.seg_off, dtype=np.int32)
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.icnv_plan_create(
                    self.device.index,
                    layout.n_genes,
                    len(layout.chromosomes),
                    gi.ctypes.data_as(_lib.c_i32p),
                    so.ctypes.data_as(_lib.c_i32p),
                    layout.window,
                    layout.step,
                    C.byref(handle),
                ),
                "icnv_plan_create",
            )
        self.handle = handle
        k = C.c_int64()
        _lib.check(self.lib.icnv_plan_out_width(self.handle, C.byref(k)))
        self.K = int(k.value)
        off = np.zeros(len(layout.chromosomes) + 1, dtype=np.int64)
        _lib.check(self.lib.icnv_plan_out_offsets(self.handle, off.ctypes.data_as(_lib.c_i64p)))
        self.out_off = off
        if not np.array_equal(off, layout.out_off):  # native window grid vs the host restatement
            raise _lib.IcnvError("internal: native and host output offsets disagree")
        self._ref_keepalive = None
        self.launches = 0  # kernels launched through this plan (counted per C-ABI call; bench.py's gpu_launches)

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            self.lib.icnv_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- info ---------------------------------------------------------------------------------
    @property
    def tier(self) -> int:
        return int(self.lib.icnv_plan_kernel_tier(self.handle))

    def launch_info(self) -> dict:
        v = [C.c_int32() for _ in range(4)]
        _lib.check(self.lib.icnv_plan_launch_info(self.handle, *[C.byref(x) for x in v]), "icnv_plan_launch_info")
        w = C.c_double()
        _lib.check(self.lib.icnv_plan_gather_cost(self.handle, C.byref(w)), "icnv_plan_gather_cost")
        return dict(
            ctas_per_sm=v[0].value, threads=v[1].value, smem_bytes=v[2].value, n_sm=v[3].value, tier=self.tier,
            rows=int(self.lib.icnv_plan_rows_per_iteration(self.handle)),
            wavefronts_per_gather=round(w.value, 4),
        )

    def _stream(self):
        return _lib.stream_handle(self.device)

    @property
    def table_sets(self) -> int:
        """Gather-table sets the plan keeps (2 when dense input runs the row-pair kernel and CSR the single-row one):
        ``icnv_plan_set_reference`` launches one bounds kernel per set."""
        info = self.launch_info()
        return 2 if (info["tier"] == 0 and info["rows"] == 2) else 1

    # -- reference profile ----------------------------------------------------------------------
    def colsum(self, X, row_cat=None, n_cat: int = 1):
        """Per-category column sums of this shard.  ``X``: dense float32 tensor or ``(indptr, indices, data)``."""
        torch = _torch()
        G = self.layout.n_genes
        sums = torch.empty((n_cat, G), dtype=torch.float64, device=self.device)
        counts = torch.empty((n_cat,), dtype=torch.int64, device=self.device)
        if isinstance(X, tuple):
            indptr, indices, data = X
            n = indptr.numel() - 1
            rc = self.lib.icnv_colsum_csr_f32(
                _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), n, G, _lib.ptr(row_cat), n_cat,
                _lib.ptr(sums), _lib.ptr(counts), self._stream(),
            )
        else:
            assert X.dtype == torch.float32 and X.stride(1) == 1
            rc = self.lib.icnv_colsum_dense_f32(
                _lib.ptr(X), X.shape[0], X.stride(0), G, _lib.ptr(row_cat), n_cat, _lib.ptr(sums), _lib.ptr(counts),
                self._stream(),
            )
        _lib.check(rc, "icnv_colsum")
        self.launches += 2 if isinstance(X, tuple) else 3
        return sums, counts

    def mean_from_sums(self, sums, counts, f64: bool = False):
        torch = _torch()
        ref = torch.empty(sums.shape, dtype=torch.float64 if f64 else torch.float32, device=self.device)
        _lib.check(
            self.lib.icnv_mean_from_sums(
                _lib.ptr(sums), _lib.ptr(counts), sums.shape[0], sums.shape[1], _lib.ptr(ref), int(f64), self._stream()
            )
        )
        self.launches += 1
        return ref

    def set_reference(self, ref):
        """``ref``: device tensor [n_cat, G], float32 or float64 (float64 -> float64 centring)."""
        torch = _torch()
        assert ref.dim() == 2 and ref.shape[1] == self.layout.n_genes and ref.is_contiguous()
        assert ref.dtype in (torch.float32, torch.float64)
        _lib.check(
            self.lib.icnv_plan_set_reference(self.handle, _lib.ptr(ref), ref.shape[0], int(ref.dtype == torch.float64), self._stream()),
            "icnv_plan_set_reference",
        )
        self.launches += self.table_sets
        self._ref_keepalive = ref

    # -- smoothing --------------------------------------------------------------------------------
    def tmp_width(self) -> int:
        w = C.c_int64()
        _lib.check(self.lib.icnv_plan_tmp_width(self.handle, C.byref(w)), "icnv_plan_tmp_width")
        return int(w.value)

    def smooth(self, X, lfc_clip: float, tmp=None):
        """Steps 1-3 for the rows of ``X`` -> ``tmp [n, tmp_width]`` float64 (smoothed rows, warp-tile order)."""
        torch = _torch()
        n = X[0].numel() - 1 if isinstance(X, tuple) else X.shape[0]
        if tmp is None:
            tmp = torch.empty((n, self.tmp_width()), dtype=torch.float64, device=self.device)
        assert tmp.dtype == torch.float64
        ld = tmp.stride(0) if n else self.tmp_width()
        if isinstance(X, tuple):
            indptr, indices, data = X
            rc = self.lib.icnv_smooth_csr_f32(
                self.handle, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), n, float(lfc_clip), _lib.ptr(tmp), ld, self._stream()
            )
        else:
            assert X.dtype == torch.float32 and X.stride(1) == 1
            rc = self.lib.icnv_smooth_dense_f32(
                self.handle, _lib.ptr(X), n, X.stride(0), float(lfc_clip), _lib.ptr(tmp), ld, self._stream()
            )
        _lib.check(rc, "icnv_smooth")
        self.launches += 1
        return tmp

    def center(self, tmp, out=None, row_stats=None, out_dtype=None):
        """Step 4: exact row median + centring: ``tmp`` -> ``(out [n, K] natural order, row_stats [n, 2])``."""
        torch = _torch()
        n = tmp.shape[0]
        if out is None:
            out = torch.empty((n, self.K), dtype=out_dtype or torch.float32, device=self.device)
        if row_stats is None:
            row_stats = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        if n:
            _lib.check(
                self.lib.icnv_center_rows(
                    self.handle, _lib.ptr(tmp), n, tmp.stride(0), _lib.ptr(out), int(out.dtype == torch.float64), out.stride(0),
                    _lib.ptr(row_stats), self._stream(),
                ),
                "icnv_center_rows",
            )
            self.launches += 1
        return out, row_stats

    def threshold(self, out, row_stats, chunk_rows: int, dynamic_threshold):
        """Step 5 in place on ``out`` (``dynamic_threshold=None``: statistics only).
        Returns ``(thr or None, row_abs_sum, row_nnz)``."""
        torch = _torch()
        n, K = out.shape
        row_abs = torch.empty((n,), dtype=torch.float64, device=self.device)
        row_nnz = torch.empty((n,), dtype=torch.int32, device=self.device)
        thr = None
        is64 = int(out.dtype == torch.float64)
        if dynamic_threshold is not None and n > 0:
            n_chunks = math.ceil(n / chunk_rows)
            thr = torch.empty((n_chunks,), dtype=torch.float64, device=self.device)
            _lib.check(
                self.lib.icnv_chunk_threshold(_lib.ptr(row_stats), n, K, chunk_rows, float(dynamic_threshold), _lib.ptr(thr), self._stream()),
                "icnv_chunk_threshold",
            )
            self.launches += 1
        if n > 0:
            self.launches += 1
            _lib.check(
                self.lib.icnv_apply_threshold(
                    _lib.ptr(out), is64, n, K, out.stride(0), chunk_rows, _lib.ptr(thr), _lib.ptr(row_abs), _lib.ptr(row_nnz), self._stream()
                ),
                "icnv_apply_threshold",
            )
        return thr, row_abs, row_nnz

    def chunk_thresholds(self, row_stats, n: int, K: int, chunk_rows: int, dynamic_threshold):
        """Per-chunk noise thresholds ``dynamic_threshold * std(chunk)`` (``:450``) from the row moments; ``None`` when the
        filter is off."""
        torch = _torch()
        if dynamic_threshold is None or n == 0:
            return None
        thr = torch.empty((math.ceil(n / chunk_rows),), dtype=torch.float64, device=self.device)
        _lib.check(
            self.lib.icnv_chunk_threshold(_lib.ptr(row_stats), n, K, chunk_rows, float(dynamic_threshold), _lib.ptr(thr), self._stream()),
            "icnv_chunk_threshold",
        )
        self.launches += 1
        return thr

    def filter_to_csr(self, out, row_stats, chunk_rows: int, dynamic_threshold, indptr=None, indices=None, data=None, data_dtype=None):
        """Steps 5 + CSR (``:449-455``) without rewriting ``out``: thresholds, a counting pass, the indptr scan and one
        compaction pass that applies the filter on the fly.  Returns ``(thr, row_abs_sum, row_nnz, (indptr, indices, data))``.
        With caller-provided ``indices`` / ``data`` nothing is read back (asynchronous); ``data_dtype`` float64 widens on
        the device (the reference's container dtype)."""
        torch = _torch()
        n, K = out.shape
        is64 = int(out.dtype == torch.float64)
        thr = self.chunk_thresholds(row_stats, n, K, chunk_rows, dynamic_threshold)
        row_abs = torch.empty((n,), dtype=torch.float64, device=self.device)
        row_nnz = torch.empty((n,), dtype=torch.int32, device=self.device)
        if indptr is None:
            indptr = torch.empty((n + 1,), dtype=torch.int64, device=self.device)
        if n:
            _lib.check(
                self.lib.icnv_filter_count(_lib.ptr(out), is64, n, K, out.stride(0), chunk_rows, _lib.ptr(thr), _lib.ptr(row_abs),
                                           _lib.ptr(row_nnz), self._stream()),
                "icnv_filter_count",
            )
            self.launches += 1
        _lib.check(self.lib.icnv_nnz_to_indptr(_lib.ptr(row_nnz), n, _lib.ptr(indptr), self._stream()), "icnv_nnz_to_indptr")
        self.launches += 1
        if indices is None:
            nnz = int(indptr[-1].item())
            indices = torch.empty((nnz,), dtype=torch.int32, device=self.device)
            data = torch.empty((nnz,), dtype=data_dtype or out.dtype, device=self.device)
        else:
            assert data is not None and indices.dtype == torch.int32
            nnz = indices.numel()
        if n and nnz:
            _lib.check(
                self.lib.icnv_filter_to_csr(_lib.ptr(out), is64, n, K, out.stride(0), chunk_rows, _lib.ptr(thr), _lib.ptr(indptr),
                                            _lib.ptr(indices), _lib.ptr(data), int(data.dtype == torch.float64), self._stream()),
                "icnv_filter_to_csr",
            )
            self.launches += 1
        return thr, row_abs, row_nnz, (indptr, indices, data)

    def gene_values(self, tmp, chunk_rows: int, thr=None, out=None):
        """Per-gene layer of ``calculate_gene_values=True`` for the rows of ``tmp`` (output of :meth:`smooth`):
        ``[n, n_genes]`` float64 in the matrix's column order, NaN where no kept window covers the gene.
        ``thr``: the per-chunk thresholds :meth:`threshold` returned (``None`` = no noise filter)."""
        torch = _torch()
        n = tmp.shape[0]
        G = self.layout.n_genes
        if out is None:
            out = torch.empty((n, G), dtype=torch.float64, device=self.device)
        assert out.dtype == torch.float64 and out.stride(1) == 1
        if self.K == 0:  # no chromosome takes part: nothing is covered
            out.fill_(float("nan"))
        elif n:
            _lib.check(
                self.lib.icnv_gene_values(
                    self.handle, _lib.ptr(tmp), n, tmp.stride(0), int(chunk_rows), _lib.ptr(thr), _lib.ptr(out), out.stride(0),
                    self._stream(),
                ),
                "icnv_gene_values",
            )
        return out

    @property
    def n_covered(self) -> int:
        v = C.c_int32()
        _lib.check(self.lib.icnv_plan_gene_coverage(self.handle, C.byref(v)), "icnv_plan_gene_coverage")
        return int(v.value)

    def to_csr(self, out, row_nnz, indptr=None, indices=None, data=None):
        """Dense thresholded block -> device CSR ``(indptr int64, indices int32, data)``.  With caller-provided buffers
        (``indices`` / ``data`` at least nnz long) nothing is read back, so the call stays asynchronous."""
        torch = _torch()
        n, K = out.shape
        if indptr is None:
            indptr = torch.empty((n + 1,), dtype=torch.int64, device=self.device)
        _lib.check(self.lib.icnv_nnz_to_indptr(_lib.ptr(row_nnz), n, _lib.ptr(indptr), self._stream()), "icnv_nnz_to_indptr")
        self.launches += 1
        if indices is None:
            nnz = int(indptr[-1].item())
            indices = torch.empty((nnz,), dtype=torch.int32, device=self.device)
            data = torch.empty((nnz,), dtype=out.dtype, device=self.device)
        else:
            assert data is not None and data.dtype == out.dtype and indices.dtype == torch.int32
            nnz = indices.numel()
        if n and nnz:
            _lib.check(
                self.lib.icnv_dense_to_csr(
                    _lib.ptr(out), int(out.dtype == torch.float64), n, K, out.stride(0), _lib.ptr(indptr), _lib.ptr(indices),
                    _lib.ptr(data), self._stream(),
                ),
                "icnv_dense_to_csr",
            )
            self.launches += 1
        return indptr, indices, data


# ---------------------------------------------------------------------------------------------
def allreduce_sums(sums, counts):
    """The single exchange of the path: sum the per-category column sums and
    row counts over all ranks when ``torch.distributed`` is initialised.  NCCL reduces the device
    tensors in place over NVLink; a gloo group (CPU tests) goes through host copies."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sums, counts
    if dist.get_backend() == "nccl":
        dist.all_reduce(sums)
        dist.all_reduce(counts)
        return sums, counts
    s, c = sums.cpu(), counts.cpu()
    dist.all_reduce(s)
    dist.all_reduce(c)
    return s.to(sums.device), c.to(counts.device)


def _dist_world():
    """(initialised, world size) of the default torch.distributed group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return True, dist.get_world_size()
    return False, 1


def allreduce_host_counts(vec: np.ndarray) -> np.ndarray:
    """Sum a small host int64 vector over all ranks (host-side validation that every rank has to agree on:
    a rank raising alone would leave the others hanging in the next collective).  Identity without a group."""
    import torch
    import torch.distributed as dist

    on, world = _dist_world()
    v = np.ascontiguousarray(vec, dtype=np.int64)
    if not on or world == 1:
        return v
    t = torch.from_numpy(v.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()
    dist.all_reduce(t)
    return t.numpy()


def global_label_order(local_labels) -> list:
    """Deterministic label list shared by every rank: labels in order of first appearance over the ranks' shards taken
    in rank order, i.e. ``pd.unique`` of the concatenated column when shards are contiguous row blocks — the key order
    of the reference's result dict (tl/_scores.py:65-68).  Without a group: the local order."""
    import torch.distributed as dist

    local = list(local_labels)
    on, world = _dist_world()
    if not on or world == 1:
        return local
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    seen, order = set(), []
    for part in gathered:
        for lab in part:
            if lab not in seen:
                seen.add(lab)
                order.append(lab)
    return order


def label_scores(lib, row_abs, labels_dev, n_labels: int, K: int, device):
    """Per-label ``sum|x| / (rows * K)`` with the cross-rank reduction (tl/_scores.py:65-68)."""
    import torch

    label_sum = torch.empty((n_labels,), dtype=torch.float64, device=device)
    label_rows = torch.empty((n_labels,), dtype=torch.int64, device=device)
    _lib.check(
        lib.icnv_label_sums(
            _lib.ptr(row_abs), _lib.ptr(labels_dev), row_abs.numel(), n_labels, _lib.ptr(label_sum), _lib.ptr(label_rows),
            _lib.stream_handle(device),
        ),
        "icnv_label_sums",
    )
    label_sum, label_rows = allreduce_sums(label_sum, label_rows)
    return label_sum / (label_rows.to(torch.float64) * float(K))
