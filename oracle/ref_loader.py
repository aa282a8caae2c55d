"""Load the UNMODIFIED reference hot-path modules from ``/root/reference``.

TEST INFRASTRUCTURE ONLY (build container).  ``import infercnvpy`` fails in
this image because ``scanpy``/``anndata``/``matplotlib`` are not installed
(``/root/reference/src/infercnvpy/__init__.py:5``), but the two files that
hold the hot path only need numpy/pandas/scipy/tqdm plus

* ``anndata.AnnData`` as a type hint (``tl/_infercnv.py:10``, ``tl/_scores.py:9``),
* ``scanpy.logging.warning`` (``tl/_infercnv.py:11``),
* ``infercnvpy._util`` (``tl/_infercnv.py:15``).

We register tiny stand-ins for those names in ``sys.modules`` and exec the
reference files from where they lie.  Nothing is copied into this repo.  The
GPU box has no ``/root/reference``: ``available()`` is False there and every
consumer (golden generation, the cross-check tests) skips.
"""

from __future__ import annotations

import importlib.util
import logging
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd

REF_SRC = Path("/root/reference/src/infercnvpy")
_cache: dict = {}


def available() -> bool:
    return (REF_SRC / "tl" / "_infercnv.py").is_file()


class MiniAnnData:
    """Just enough of AnnData for ``infercnv()`` / ``cnv_score()`` to run."""

    def __init__(self, X, obs=None, var=None, layers=None, obsm=None):
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(g).astype(str))
        self.layers = layers if layers is not None else {}
        self.obsm = obsm if obsm is not None else {}
        self.uns = {}

    @property
    def shape(self):
        return self.X.shape

    @property
    def var_names(self):
        return self.var.index

    def __getitem__(self, key):
        rows, cols = key
        if isinstance(rows, slice) and rows == slice(None):  # adata[:, gene_mask]  (_infercnv.py:110)
            cols = np.asarray(cols)
            return MiniAnnData(
                self.X[:, cols],
                obs=self.obs,
                var=self.var.loc[cols],
                layers={k: v[:, cols] for k, v in self.layers.items()},
            )
        assert isinstance(cols, slice) and cols == slice(None)  # adata[cell_mask, :]  (_scores.py:131,204)
        rows = np.asarray(rows)
        return MiniAnnData(
            self.X[rows],
            obs=self.obs.loc[rows],
            var=self.var,
            layers={k: v[rows] for k, v in self.layers.items()},
            obsm={k: v[rows] for k, v in self.obsm.items()},
        )


def _load_file(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return ``(ref_infercnv_module, ref_scores_module)``."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")

    if "anndata" not in sys.modules:
        ad = types.ModuleType("anndata")
        ad.AnnData = MiniAnnData
        sys.modules["anndata"] = ad
    if "scanpy" not in sys.modules:
        sc = types.ModuleType("scanpy")
        lg = types.ModuleType("scanpy.logging")
        _log = logging.getLogger("ref-infercnvpy")
        lg.warning = lambda msg, *a, **k: _log.debug(msg)
        lg.info = lambda msg, *a, **k: _log.debug(msg)
        sc.logging = lg
        sys.modules["scanpy"] = sc
        sys.modules["scanpy.logging"] = lg

    pkg = types.ModuleType("infercnvpy")
    pkg.__path__ = [str(REF_SRC)]
    sys.modules.setdefault("infercnvpy", pkg)
    tl_pkg = types.ModuleType("infercnvpy.tl")
    tl_pkg.__path__ = [str(REF_SRC / "tl")]
    sys.modules.setdefault("infercnvpy.tl", tl_pkg)

    _load_file("infercnvpy._util", REF_SRC / "_util.py")
    m_inf = _load_file("infercnvpy.tl._infercnv", REF_SRC / "tl" / "_infercnv.py")
    m_sco = _load_file("infercnvpy.tl._scores", REF_SRC / "tl" / "_scores.py")
    _cache["mods"] = (m_inf, m_sco)
    return _cache["mods"]
