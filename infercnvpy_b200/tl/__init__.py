"""Tools: ``cnv.tl.*`` (reference: ``/root/reference/src/infercnvpy/tl/__init__.py``)."""

from ._embed import tsne, umap
from ._infercnv import infercnv
from ._ith import ithcna, ithgex
from ._leiden import leiden
from ._pca import pca
from ._scores import cnv_score

__all__ = ["infercnv", "cnv_score", "ithcna", "ithgex", "pca", "leiden", "umap", "tsne"]
