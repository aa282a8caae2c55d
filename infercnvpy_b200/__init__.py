"""B200-native drop-in for the CNV-inference hot path of infercnvpy.

    import infercnvpy_b200 as cnv
    cnv.tl.infercnv(adata, reference_key="cell_type", reference_cat=["T cell"], window_size=250)
    cnv.tl.cnv_score(adata, groupby="sample")

Reference namespace layout: ``/root/reference/src/infercnvpy/__init__.py:3-8``.
"""

from . import datasets, io, pl, pp, tl
from ._anndata import AnnData
from ._layout import build_layout, shard_rows

__version__ = "0.1.0"
__all__ = ["tl", "pp", "pl", "io", "datasets", "AnnData", "build_layout", "shard_rows"]
