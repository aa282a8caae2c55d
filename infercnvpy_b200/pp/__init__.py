"""Preprocessing: ``cnv.pp.neighbors`` (reference: ``/root/reference/src/infercnvpy/pp/__init__.py:8-43``)."""

from ._neighbors import neighbors

__all__ = ["neighbors"]
