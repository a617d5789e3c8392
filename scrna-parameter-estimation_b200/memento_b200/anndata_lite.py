"""Minimal AnnData stand-in (the image has no ``anndata``).

The reference API touches only ``X, obs, var, uns, shape, copy()`` and
``_inplace_subset_var(mask)`` (reference main.py:40,44,47,83,115,124,229,271,359), so tests, the
bench and the golden-vector generator use this class; a real ``anndata.AnnData`` works unchanged.
"""
import copy as _copy

import numpy as np
import pandas as pd


class AnnDataLite:
    def __init__(self, X, obs=None, var=None, uns=None):
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.Index(["g%d" % i for i in range(g)]))
        self.uns = uns if uns is not None else {}

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    def copy(self):
        return AnnDataLite(self.X.copy(), self.obs.copy(), self.var.copy(), _copy.deepcopy(self.uns))

    def _inplace_subset_var(self, mask):
        mask = np.asarray(mask)
        self.X = self.X[:, mask]
        self.var = self.var.iloc[np.flatnonzero(mask)] if mask.dtype == bool else self.var.iloc[mask]
