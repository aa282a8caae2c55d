"""ctypes binding of ``libicnv.so`` (C ABI declared in ``include/icnv.h``).

There is no CPU fallback: importing the handle without the shared library, or
calling into it without a CUDA device, raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libicnv.so"
_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/icnv.h declares
SIGNATURES = {
    "icnv_last_error": (C.c_char_p, []),
    "icnv_version": (C.c_int, []),
    "icnv_plan_create": (C.c_int, [C.c_int, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32, C.c_int32, C.POINTER(c_vp)]),
    "icnv_plan_destroy": (None, [c_vp]),
    "icnv_plan_out_width": (C.c_int, [c_vp, c_i64p]),
    "icnv_plan_out_offsets": (C.c_int, [c_vp, c_i64p]),
    "icnv_plan_kernel_tier": (C.c_int, [c_vp]),
    "icnv_plan_rows_per_iteration": (C.c_int, [c_vp]),
    "icnv_plan_launch_info": (C.c_int, [c_vp, c_i32p, c_i32p, c_i32p, c_i32p]),
    "icnv_colsum_dense_f32": (C.c_int, [c_vp, C.c_int64, C.c_int64, C.c_int32, c_vp, C.c_int32, c_vp, c_vp, c_vp]),
    "icnv_colsum_csr_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_int64, C.c_int32, c_vp, C.c_int32, c_vp, c_vp, c_vp]),
    "icnv_mean_from_sums": (C.c_int, [c_vp, c_vp, C.c_int32, C.c_int32, c_vp, C.c_int32, c_vp]),
    "icnv_nnz_to_indptr": (C.c_int, [c_vp, C.c_int64, c_vp, c_vp]),
    "icnv_plan_set_reference": (C.c_int, [c_vp, c_vp, C.c_int32, C.c_int32, c_vp]),
    "icnv_smooth_dense_f32": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_double, c_vp, C.c_int64, c_vp]),
    "icnv_smooth_csr_f32": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int64, C.c_double, c_vp, C.c_int64, c_vp]),
    "icnv_center_rows": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, c_vp, C.c_int32, C.c_int64, c_vp, c_vp]),
    "icnv_chunk_threshold": (C.c_int, [c_vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_vp, c_vp]),
    "icnv_apply_threshold": (C.c_int, [c_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp, c_vp, c_vp]),
    "icnv_filter_count": (C.c_int, [c_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp, c_vp, c_vp]),
    "icnv_filter_to_csr": (C.c_int, [c_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp, c_vp, c_vp, C.c_int32, c_vp]),
    "icnv_gene_values": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp, C.c_int64, c_vp]),
    "icnv_plan_gene_coverage": (C.c_int, [c_vp, c_i32p]),
    "icnv_plan_gather_cost": (C.c_int, [c_vp, C.POINTER(C.c_double)]),
    "icnv_plan_tmp_width": (C.c_int, [c_vp, c_i64p]),
    "icnv_dense_to_csr": (C.c_int, [c_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp, c_vp, c_vp]),
    "icnv_rowabs_csr": (C.c_int, [c_vp, c_vp, C.c_int32, C.c_int64, c_vp, c_vp]),
    "icnv_rowabs_dense": (C.c_int, [c_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, c_vp, c_vp]),
    "icnv_row_corrcoef_f64": (C.c_int, [c_vp, C.c_int64, C.c_int64, C.c_int32, c_vp, C.c_int64, c_vp, c_vp]),
    "icnv_csr_to_dense_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_int32, C.c_int64, C.c_int32, c_vp, C.c_int64, c_vp]),
    "icnv_gram_f32": (C.c_int, [c_vp, C.c_int64, C.c_int64, C.c_int32, c_vp, c_vp]),
    "icnv_project_f32": (C.c_int, [c_vp, C.c_int64, C.c_int64, C.c_int32, c_vp, C.c_int32, c_vp, c_vp, c_vp]),
    "icnv_knn_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "icnv_knn_f32": (C.c_int, [c_vp, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, c_vp, c_vp, C.c_int32, c_vp, c_vp]),
    "icnv_fuzzy_rows": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int32, C.c_int64, C.c_float, c_vp, c_vp, c_vp, c_vp]),
    "icnv_weighted_degree": (C.c_int, [c_vp, c_vp, C.c_int64, c_vp, c_vp]),
    "icnv_community_sweep_work_bytes": (C.c_int64, [C.c_int64]),
    "icnv_community_sweep": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, C.c_double, C.c_double, C.c_int32, c_vp, c_vp, c_vp, c_vp]),
    "icnv_umap_epochs": (C.c_int, [c_vp, c_vp, C.c_int64, c_vp, C.c_int32, c_vp, c_vp, c_vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, c_vp]),
    "icnv_tsne_affinities": (C.c_int, [c_vp, C.c_int32, C.c_int32, C.c_int64, C.c_float, c_vp, c_vp]),
    "icnv_tsne_work_floats": (C.c_int64, [C.c_int32]),
    "icnv_tsne_iterations": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, c_vp]),
    "icnv_host_schedule_gathers": (C.c_double, [c_i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i32p, C.POINTER(C.c_uint8)]),
    "icnv_debug_set_timeline": (C.c_int, [c_vp, C.c_int]),
    "icnv_label_sums": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int32, c_vp, c_vp, c_vp]),
}


class IcnvError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load ``libicnv.so`` (building is the job of ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    path = Path(os.environ.get("ICNV_LIB_PATH", _LIB_PATH))  # developer A/B builds only
    if not path.exists():
        raise IcnvError(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). infercnvpy_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().icnv_last_error()
        raise IcnvError(f"{what or 'libicnv'} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_handle(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
