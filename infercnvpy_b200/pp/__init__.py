"""Preprocessing: ``cnv.pp.*`` (reference: ``/root/reference/src/infercnvpy/pp/__init__.py``)."""

__all__: list[str] = []
