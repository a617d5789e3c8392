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


def gev_battery_vector(i):
    """Coefficient row i of the GEV battery (tests/golden/gev_battery.npz): x[0] = statistic, x[1:] = bootstrap
    values around it.  Normal / Student-t(5) / skewed (centred gamma) nulls of 1000-10000 replicates with effects of
    3-6 null standard deviations in either direction, i.e. the rows that reach the GEV branch of _compute_asl
    (hypothesis_test.py:94-141) at the default num_boot."""
    import numpy as np
    rng = np.random.default_rng(1000 + i)
    n = (1000, 2000, 4000, 10000)[i % 4]
    kind = i % 3
    z = (3.0, 3.5, 4.0, 5.0, 6.0)[i % 5]
    sign = 1.0 if (i // 2) % 2 == 0 else -1.0
    if kind == 0:
        noise = rng.normal(0, 1, n)
    elif kind == 1:
        noise = rng.standard_t(5, n) / np.sqrt(5.0 / 3.0)
    else:
        noise = (rng.gamma(4.0, 1.0, n) - 4.0) / 2.0
    stat = sign * z * 0.1
    return np.concatenate([[stat], stat + 0.1 * noise])
