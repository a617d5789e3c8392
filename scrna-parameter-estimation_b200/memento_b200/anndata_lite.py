"""Minimal AnnData stand-in (the image has no ``anndata``).

The reference API touches only ``X, obs, var, uns, shape, copy()`` and
``_inplace_subset_var(mask)`` (reference main.py:40,44,47,83,115,124,229,271,359), so tests, the
bench and the golden-vector generator use this class; a real ``anndata.AnnData`` works unchanged.
"""
import copy as _copy

import numpy as np
import pandas as pd


class AnnDataLite:
    """``X`` is materialised lazily after ``_inplace_subset_var``: the host column slice of a 500M-nonzero matrix
    costs seconds (scipy fancy indexing) and nothing on the device path reads it."""

    def __init__(self, X, obs=None, var=None, uns=None):
        self._X = X
        self._cols = None          # pending column selection (int index array into self._X)
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.Index(["g%d" % i for i in range(g)]))
        self.uns = uns if uns is not None else {}

    @property
    def X(self):
        if self._cols is not None:
            self._X = self._X[:, self._cols]
            self._cols = None
        return self._X

    @X.setter
    def X(self, value):
        self._X, self._cols = value, None

    @property
    def shape(self):
        return (self._X.shape[0], self._X.shape[1] if self._cols is None else int(self._cols.size))

    @property
    def n_obs(self):
        return self.shape[0]

    @property
    def n_vars(self):
        return self.shape[1]

    def copy(self):
        return AnnDataLite(self.X.copy(), self.obs.copy(), self.var.copy(), _copy.deepcopy(self.uns))

    def _inplace_subset_var(self, mask):
        mask = np.asarray(mask)
        idx = np.flatnonzero(mask) if mask.dtype == bool else mask.astype(np.int64)
        self._cols = idx if self._cols is None else self._cols[idx]
        self.var = self.var.iloc[idx]
