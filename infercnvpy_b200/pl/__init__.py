"""Group summaries of the CNV matrix (the arithmetic behind the reference's summary heatmap,
``/root/reference/src/infercnvpy/pl/_chromosome_heatmap.py:90-189``).  Rendering itself (matplotlib / scanpy.pl.heatmap) is
out of scope; these functions return the numbers a plot would show."""

from ._summary import chromosome_heatmap_summary, group_means

__all__ = ["chromosome_heatmap_summary", "group_means"]
