"""Shared helpers for the parity tests: rebuild the golden dataset and compare arrays."""
import os

import numpy as np
import pandas as pd
import scipy.sparse as sp

from memento_b200.anndata_lite import AnnDataLite

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_adata(st=None):
    """The exact input the golden generator used, rebuilt from the arrays stored in stages.npz
    (so the tests do not depend on the RNG stream of the synthetic generator)."""
    st = load("stages.npz") if st is None else st
    X = sp.csr_matrix((st["X_data"], st["X_indices"], st["X_indptr"]), shape=tuple(st["X_shape"]))
    n, g = X.shape
    obs = pd.DataFrame({"stim": st["stim"], "cell": st["cell"], "q": st["q"]},
                       index=pd.Index(["c%d" % i for i in range(n)]))
    var = pd.DataFrame(index=pd.Index(["gene%d" % i for i in range(g)]))
    return AnnDataLite(X, obs, var)


def assert_close(a, b, rtol, atol=0.0, what=""):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), (what, "NaN pattern differs", np.flatnonzero(nan_a != nan_b)[:10])
    ok = ~nan_a
    inf = np.isinf(a) & ok
    assert np.array_equal(a[inf], b[inf]), (what, "inf pattern differs")
    ok &= ~np.isinf(a)
    err = np.abs(a[ok] - b[ok]) - (atol + rtol * np.abs(b[ok]))
    assert (err <= 0).all(), (what, "max excess", err.max(), "at", np.argmax(err),
                               a[ok][np.argmax(err)], b[ok][np.argmax(err)])
