"""On-disk matrix -> this rank's rows in HBM, and ``X_cnv`` back to disk.

The reference reaches the hot path through an in-memory AnnData whose matrix it first copies twice
(``adata[:, ~var_mask]`` and ``.tocsr()``, ``/root/reference/src/infercnvpy/tl/_infercnv.py:110-116``).  Here the
matrix goes from the file straight into device buffers, row shard by row shard (``shard_rows``: cuts at multiples of
``chunksize``), through two pinned staging buffers: a reader thread fills one from the file (page cache / disk) while
the DMA engine drains the other.  The result is what ``cnv.tl.infercnv`` takes as ``adata.X`` without another copy:
a device ``torch.Tensor`` (dense) or a :class:`DeviceCSR`.

Formats (the image has no ``h5py``; ``.h5ad`` is read through it when it is importable):

* ``.npy``                     dense ``[n_obs, n_vars]``, memory-mapped;
* directory with ``indptr.npy`` / ``indices.npy`` / ``data.npy`` (+ optional ``shape.npy``)   CSR, memory-mapped;
* ``.npz`` written by ``scipy.sparse.save_npz(..., compressed=False)``   CSR, members read in place by offset;
* ``.h5ad``                    ``X`` dense dataset or CSR group (needs ``h5py``).
"""

from __future__ import annotations

import threading
import zipfile
from pathlib import Path

import numpy as np

from ._layout import shard_rows


class DeviceCSR:
    """CSR matrix whose three arrays live in HBM (``indptr`` int64 ``[n + 1]``, ``indices`` int32, ``data`` float32)."""

    format = "csr"

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.shape = (int(shape[0]), int(shape[1]))
        self.dtype = np.dtype("float32")

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    def rows(self, r0: int, r1: int):
        """Device CSR triple of rows ``[r0, r1)`` (indptr rebased to 0; views, no copy of the entries)."""
        if r0 == 0 and r1 == self.shape[0]:
            return self.indptr, self.indices, self.data
        e0, e1 = int(self.indptr[r0].item()), int(self.indptr[r1].item())
        return self.indptr[r0 : r1 + 1] - e0, self.indices[e0:e1], self.data[e0:e1]


def _device(device):
    import torch

    if not torch.cuda.is_available():
        from . import _lib

        raise _lib.IcnvError("infercnvpy_b200.io needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _dist_rank_world(rank, world):
    if rank is not None and world is not None:
        return int(rank), int(world)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _stream_to_device(read_into, n_items: int, item_shape: tuple, np_dtype, torch_dtype, device, slab_items: int):
    """``read_into(host_view, a, b)`` fills ``host_view`` with items ``[a, b)`` of the source.  Items are staged through two
    pinned buffers: the reader thread works on slab i + 1 while the copy engine moves slab i."""
    import torch

    out = torch.empty((n_items,) + tuple(item_shape), dtype=torch_dtype, device=device)
    if n_items == 0:
        return out
    slab_items = max(1, min(slab_items, n_items))
    bufs = [torch.empty((slab_items,) + tuple(item_shape), dtype=torch_dtype, pin_memory=True) for _ in range(2)]
    views = [b.numpy() for b in bufs]
    spans = [(a, min(n_items, a + slab_items)) for a in range(0, n_items, slab_items)]
    stream = torch.cuda.Stream(device)
    done = [None, None]  # DMA-complete event of the slab last sent from each buffer
    filled = [threading.Event(), threading.Event()]
    issued = [threading.Event(), threading.Event()]
    err = []

    def reader():
        try:
            for i, (a, b) in enumerate(spans):
                k = i % 2
                if i >= 2:  # the buffer is free again once the DMA of slab i - 2 has finished
                    issued[k].wait()
                    issued[k].clear()
                    done[k].synchronize()
                read_into(views[k][: b - a], a, b)
                filled[k].set()
        except Exception as e:  # pragma: no cover
            err.append(e)
            for ev in filled:
                ev.set()

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    for i, (a, b) in enumerate(spans):
        k = i % 2
        filled[k].wait()
        filled[k].clear()
        if err:
            raise err[0]
        with torch.cuda.stream(stream):
            out[a:b].copy_(bufs[k][: b - a], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        done[k] = ev
        issued[k].set()
    th.join()
    stream.synchronize()
    torch.cuda.current_stream(device).wait_stream(stream)
    return out


def _npz_member(path: Path, name: str):
    """(file offset of the array data, shape, dtype) of an UNCOMPRESSED .npy member of a zip archive."""
    with zipfile.ZipFile(path) as zf:
        info = zf.getinfo(name + ".npy")
        if info.compress_type != zipfile.ZIP_STORED:
            raise ValueError(f"{path}: member {name} is compressed; write it with scipy.sparse.save_npz(..., compressed=False)")
        with zf.open(info) as f:
            version = np.lib.format.read_magic(f)
            if version == (1, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
            else:
                shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
            if fortran and len(shape) > 1:
                raise ValueError(f"{path}: member {name} is Fortran-ordered")
            header_len = f.tell()
        with open(path, "rb") as raw:
            raw.seek(info.header_offset)
            local = raw.read(30)
            n_name, n_extra = int.from_bytes(local[26:28], "little"), int.from_bytes(local[28:30], "little")
        return info.header_offset + 30 + n_name + n_extra + header_len, shape, dtype


def _open_arrays(path: Path):
    """-> ("dense", memmap) or ("csr", indptr, indices_src, data_src, shape) with ``*_src`` sliceable array-likes."""
    if path.is_dir():
        ip = np.load(path / "indptr.npy", mmap_mode="r")
        ix = np.load(path / "indices.npy", mmap_mode="r")
        dv = np.load(path / "data.npy", mmap_mode="r")
        if (path / "shape.npy").exists():
            shape = tuple(int(v) for v in np.load(path / "shape.npy"))
        else:
            shape = (len(ip) - 1, int(ix.max()) + 1 if len(ix) else 0)
        return ("csr", ip, ix, dv, shape)
    suffix = path.suffix.lower()
    if suffix == ".npy":
        return ("dense", np.load(path, mmap_mode="r"))
    if suffix == ".npz":
        with np.load(path, allow_pickle=False) as z:
            fmt = z["format"].item() if "format" in z.files else b"csr"
            fmt = fmt.decode() if isinstance(fmt, bytes) else str(fmt)
            if fmt != "csr":
                raise ValueError(f"{path}: sparse format {fmt!r}; only CSR rows can be sharded")
            shape = tuple(int(v) for v in z["shape"])
        arrs = {}
        for name in ("indptr", "indices", "data"):
            off, shp, dt = _npz_member(path, name)
            arrs[name] = np.memmap(path, mode="r", dtype=dt, shape=shp, offset=off)
        return ("csr", arrs["indptr"], arrs["indices"], arrs["data"], shape)
    if suffix == ".h5ad":
        try:
            import h5py
        except ImportError as e:
            raise ImportError("reading .h5ad needs h5py, which is not installed in this environment") from e
        f = h5py.File(path, "r")
        X = f["X"]
        if isinstance(X, h5py.Dataset):
            return ("dense", X)
        enc = X.attrs.get("encoding-type", X.attrs.get("h5sparse_format", "csr_matrix"))
        enc = enc.decode() if isinstance(enc, bytes) else str(enc)
        if not enc.startswith("csr"):
            raise ValueError(f"{path}: X is stored as {enc}; only CSR rows can be sharded")
        shape = tuple(int(v) for v in X.attrs.get("shape", X.attrs.get("h5sparse_shape")))
        return ("csr", X["indptr"], X["indices"], X["data"], shape)
    raise ValueError(f"{path}: unknown matrix container (expected .npy, .npz, .h5ad or a CSR directory)")


def read_matrix(path, *, rank: int | None = None, world: int | None = None, chunksize: int = 5000, device=None,
                slab_bytes: int = 128 << 20):
    """Load this rank's row shard of the matrix at ``path`` into HBM.

    Returns ``(X, (row0, row1), n_obs_total)`` with ``X`` a device float32 tensor ``[row1 - row0, n_vars]`` or a
    :class:`DeviceCSR`.  ``rank`` / ``world`` default to the initialised ``torch.distributed`` group (else one shard).
    Shards are cut at multiples of ``chunksize`` like everywhere else on the path, so every rank's noise-filter chunks
    are the reference's (``_infercnv.py:123``)."""
    import torch

    path = Path(path)
    device = _device(device)
    rank, world = _dist_rank_world(rank, world)
    opened = _open_arrays(path)
    if opened[0] == "dense":
        src = opened[1]
        n, g = int(src.shape[0]), int(src.shape[1])
        r0, r1 = shard_rows(n, chunksize, rank, world)

        def read_rows(view, a, b):
            np.copyto(view, src[r0 + a : r0 + b], casting="unsafe")

        X = _stream_to_device(read_rows, r1 - r0, (g,), np.float32, torch.float32, device, max(1, slab_bytes // max(1, 4 * g)))
        return X, (r0, r1), n
    _, ip_src, ix_src, dv_src, shape = opened
    n, g = shape
    r0, r1 = shard_rows(n, chunksize, rank, world)
    ip = np.asarray(ip_src[r0 : r1 + 1], dtype=np.int64)
    e0, e1 = (int(ip[0]), int(ip[-1])) if len(ip) else (0, 0)
    indptr = torch.from_numpy(ip - e0).to(device)

    def read_ix(view, a, b):
        np.copyto(view, ix_src[e0 + a : e0 + b], casting="unsafe")

    def read_dv(view, a, b):
        np.copyto(view, dv_src[e0 + a : e0 + b], casting="unsafe")

    per = max(1, slab_bytes // 4)
    indices = _stream_to_device(read_ix, e1 - e0, (), np.int32, torch.int32, device, per)
    data = _stream_to_device(read_dv, e1 - e0, (), np.float32, torch.float32, device, per)
    return DeviceCSR(indptr, indices, data, (r1 - r0, g)), (r0, r1), n


def write_cnv(path, adata, key: str = "cnv") -> None:
    """``obsm["X_<key>"]`` (CSR float64, ``_infercnv.py:153-156``) and ``uns[key]["chr_pos"]`` into one uncompressed
    ``.npz`` (scipy's ``save_npz`` layout plus two ``chr_pos_*`` members), readable by :func:`read_cnv` and by
    ``scipy.sparse.load_npz``."""
    import scipy.sparse as sp

    X = sp.csr_matrix(adata.obsm[f"X_{key}"])
    chr_pos = adata.uns[key]["chr_pos"]
    np.savez(
        Path(path), format=np.array("csr".encode("ascii")), shape=np.array(X.shape), data=X.data, indices=X.indices, indptr=X.indptr,
        chr_pos_names=np.array(list(chr_pos.keys())), chr_pos_starts=np.array([int(v) for v in chr_pos.values()], dtype=np.int64),
    )


def read_cnv(path, adata=None, key: str = "cnv"):
    """Inverse of :func:`write_cnv`: returns ``(chr_pos, X_cnv)`` and, when ``adata`` is given, stores them under the
    reference's keys."""
    import scipy.sparse as sp

    with np.load(Path(path), allow_pickle=False) as z:
        X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(int(v) for v in z["shape"]))
        chr_pos = {str(k): int(v) for k, v in zip(z["chr_pos_names"], z["chr_pos_starts"])}
    if adata is not None:
        adata.obsm[f"X_{key}"] = X
        adata.uns[key] = {"chr_pos": chr_pos}
    return chr_pos, X
