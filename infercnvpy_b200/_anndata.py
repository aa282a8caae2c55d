"""Minimal AnnData stand-in.

``anndata`` / ``scanpy`` are not installed in this image.  The public functions of
this package only touch the attributes below, so they work on a real
``anndata.AnnData`` when one is passed and on this duck type otherwise (tests,
bench, smoke).
"""

from __future__ import annotations

import numpy as np
import pandas as pd


class AnnData:
    def __init__(self, X=None, obs=None, var=None, layers=None, obsm=None, obsp=None, uns=None):
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(g).astype(str))
        if len(self.obs) != n or len(self.var) != g:
            raise ValueError("obs/var length does not match X")
        self.layers = dict(layers) if layers is not None else {}
        self.obsm = dict(obsm) if obsm is not None else {}
        self.obsp = dict(obsp) if obsp is not None else {}
        self.uns = dict(uns) if uns is not None else {}

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def var_names(self):
        return self.var.index

    def obsm_keys(self):
        return list(self.obsm.keys())

    def uns_keys(self):
        return list(self.uns.keys())

    def copy(self):
        import copy

        def dup(v):  # numpy / scipy / pandas: .copy(); torch: .clone(); anything else (device CSR views) is shared
            return v.copy() if hasattr(v, "copy") else (v.clone() if hasattr(v, "clone") else v)

        return AnnData(
            dup(self.X),
            obs=self.obs.copy(),
            var=self.var.copy(),
            layers={k: v.copy() for k, v in self.layers.items()},
            obsm={k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsm.items()},
            obsp={k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsp.items()},
            uns=copy.deepcopy(self.uns),
        )

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        r = np.arange(self.n_obs)[rows] if not isinstance(rows, slice) or rows != slice(None) else slice(None)
        c = np.arange(self.n_vars)[np.asarray(cols)] if not isinstance(cols, slice) else cols
        X = self.X[r][:, c] if not isinstance(r, slice) else self.X[:, c]
        obs = self.obs.iloc[r] if not isinstance(r, slice) else self.obs
        return AnnData(
            X,
            obs=obs,
            var=self.var.iloc[c],
            layers={k: (v[r][:, c] if not isinstance(r, slice) else v[:, c]) for k, v in self.layers.items()},
            obsm={k: (v[r] if not isinstance(r, slice) else v) for k, v in self.obsm.items()},
        )

    def __repr__(self):
        return f"AnnData(duck) n_obs x n_vars = {self.n_obs} x {self.n_vars}; obsm={list(self.obsm)}; uns={list(self.uns)}"
