"""Oracle: compressed bootstrap (TEST INFRASTRUCTURE, see oracle/__init__.py).

numpy restatement of reference ``memento/bootstrap.py``.  The RNG calls are made in the same
order and on the same generators as the reference so that, with the same seeds, the outputs
are identical: the unique-value hash uses the GLOBAL ``np.random`` state
(bootstrap.py:62-65) and every resampling call builds a fresh ``Generator(PCG64(5))``
(bootstrap.py:102, :135).
"""
import numpy as np

BOOT_SEED = 5  # reference: bootstrap.py:102, :135


def unique_table(expr, size_factor):
    """Distinct (count[, count2], size-factor) rows of one (gene[, gene2], group) slice.

    reference: bootstrap.py:62-71.  ``expr`` is a sparse (cells x 1) or (cells x 2) matrix,
    ``size_factor`` the binned size factors of the same cells.  Rows come back in ascending
    order of a random-projection code (i.e. arbitrary but reproducible under np.random.seed).
    Returns (inv_sf (U,1), inv_sf_sq (U,1), values (U,ncol), multiplicity (U,))."""
    code = expr.dot(np.random.random(expr.shape[1]))
    code = code + np.random.random() * size_factor
    _, first, mult = np.unique(code, return_index=True, return_counts=True)
    vals = expr[first].toarray()
    inv = 1.0 / size_factor[first].reshape(-1, 1)
    return inv, inv ** 2, vals, mult


def canonical_table(inv_sf, values, mult):
    """Order-independent view of a unique table: rows sorted by (values..., inv_sf)."""
    cols = [inv_sf.reshape(-1)] + [values[:, k] for k in range(values.shape[1] - 1, -1, -1)]
    order = np.lexsort(cols)
    return inv_sf.reshape(-1)[order], values[order], mult[order]


def draw_counts(n_cells, mult, num_boot):
    """(U x B) multinomial resample counts.  reference: bootstrap.py:102-103, :135-137."""
    gen = np.random.Generator(np.random.PCG64(BOOT_SEED))
    return gen.multinomial(n_cells, mult / mult.sum(), size=num_boot).T


def bootstrap_1d(col, size_factor, q, weighted_estimator, num_boot, precomputed=None):
    """Bootstrap mean / variance replicates of one gene in one group.
    reference: bootstrap.py:74-116."""
    inv_sf, inv_sf_sq, vals, mult = unique_table(col, size_factor) if precomputed is None else precomputed
    if vals.shape[0] <= 1:  # bootstrap.py:97-98
        return np.full(num_boot, np.nan), np.full(num_boot, np.nan)
    n_obs = col.shape[0]
    W = draw_counts(n_obs, mult, num_boot)
    return weighted_estimator(vals, W, n_obs, q, inv_sf, inv_sf_sq)


def bootstrap_2d(cols, size_factor, q, weighted_estimator, weighted_cov, num_boot, precomputed=None):
    """Bootstrap covariance and both variances of a gene pair in one group.
    reference: bootstrap.py:119-157 (no U<=1 guard there either)."""
    n_obs = cols.shape[0]
    inv_sf, inv_sf_sq, vals, mult = unique_table(cols, size_factor) if precomputed is None else precomputed
    W = draw_counts(n_obs, mult, num_boot)
    x = vals[:, 0].reshape(-1, 1)
    y = vals[:, 1].reshape(-1, 1)
    cov = weighted_cov(x, y, W, n_obs, q, inv_sf, inv_sf_sq)
    _, var_1 = weighted_estimator(x, W, n_obs, q, inv_sf, inv_sf_sq)
    _, var_2 = weighted_estimator(y, W, n_obs, q, inv_sf, inv_sf_sq)
    return cov, var_1, var_2
