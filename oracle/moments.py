"""Oracle: method-of-moments estimators (TEST INFRASTRUCTURE, see oracle/__init__.py).

numpy/scipy restatement of reference ``memento/estimator.py`` and of the size-factor binning in
``memento/main.py``.  dtypes flow as in the reference: with a float64 ``X`` everything is float64;
with the usual float32 ``X`` the row totals, the size factors and therefore the sparse-form
moments are float32-accurate only (scipy keeps float32 in ``X.sum`` and ``1/size_factor``), which
is why the product's float64 CUDA path is compared to float32-X reference outputs at 1e-5 and to
float64-X reference outputs at round-off.
"""
import numpy as np
import scipy.sparse as sp
import scipy.stats as stats

SUPPORTED_ESTIMATORS = ("hyper_relative", "mean_only")


def _a1(m):
    return np.asarray(m).reshape(-1)


# --------------------------------------------------------------------------- size factors
def row_totals(X):
    """Raw UMI totals per cell.  reference: estimator.py:64-69 (``total=True`` branch; the
    result is NOT normalised)."""
    return _a1(X.sum(axis=1))  # keeps X's dtype: float32 X gives float32 totals, as in the reference


def masked_size_factor(X, gene_mask, shrinkage):
    """Size factor from a gene subset.  reference: estimator.py:71-76.

    ``Nrc = X[:, mask].sum(1); Nrc += quantile(Nrc, shrinkage); sf = Nrc / mean(Nrc)``."""
    totals = _a1(X.multiply(gene_mask).sum(axis=1))
    totals += np.quantile(totals, shrinkage)  # in place: the dtype of X is kept (reference :74)
    return totals / totals.mean()


def bin_size_factor(size_factor, num_bins):
    """Equal-width binning of the size factors.  reference: main.py:138-148.

    Returns (approx_sf, bin_index) where approx_sf is the per-bin mean and cells equal to the
    global maximum keep the exact maximum; bin_index is 0-based."""
    binmean, _, binnumber = stats.binned_statistic(
        size_factor, size_factor, bins=num_bins, statistic="mean")
    idx = np.clip(binnumber, 1, binmean.shape[0])
    approx = binmean[idx - 1]
    top = size_factor.max()
    approx[size_factor == top] = top
    return approx, idx - 1


# --------------------------------------------------------------------------- 1D moments
def hyper_1d_sparse(X, n_obs, q, size_factor):
    """Hypergeometric mean / variance from a sparse cells x genes matrix.
    reference: estimator.py:175-185 (sparse branch of ``_hyper_1d_relative``)."""
    w = (1.0 / size_factor).reshape(1, -1)
    w2 = (1.0 / size_factor ** 2).reshape(1, -1)
    X = sp.csc_matrix(X) if not sp.issparse(X) else X
    m1 = _a1(X.T.dot(w.T)) / n_obs
    sq = _a1(X.power(2).T.dot(w2.T)) / n_obs
    lin = _a1(X.T.dot(w2.T)) / n_obs
    m2 = sq - (1 - q) * lin
    return [m1, m2 - m1 ** 2]


def hyper_1d_weighted(values, weights, n_obs, q, inv_sf, inv_sf_sq):
    """Tuple form: ``values`` (U x 1) distinct counts, ``weights`` (U x B) resample counts.
    reference: estimator.py:171-174."""
    m1 = (values * weights * inv_sf).sum(axis=0) / n_obs
    m2 = (values ** 2 * weights * inv_sf_sq
          - (1 - q) * values * weights * inv_sf_sq).sum(axis=0) / n_obs
    return [m1, m2 - m1 ** 2]


def mean_only_sparse(X, n_obs, q, size_factor):
    """reference: estimator.py:197-204 (``_mean_only_1p`` sparse branch): mean + 1, variance 10."""
    w = (1.0 / size_factor).reshape(1, -1)
    m1 = _a1(X.T.dot(w.T)) / n_obs
    return [m1 + 1, np.ones(m1.shape) * 10]


def mean_only_weighted(values, weights, n_obs, q, inv_sf, inv_sf_sq):
    """reference: estimator.py:194-196, :202-204."""
    m1 = (values * weights * inv_sf).sum(axis=0) / n_obs
    return [m1 + 1, np.ones(m1.shape) * 10]


def estimator_1d(estimator_type):
    """(sparse_fn, weighted_fn) for an estimator name.  reference: estimator.py:19-32; only the
    names that are actually defined in the reference are accepted."""
    if estimator_type == "hyper_relative":
        return hyper_1d_sparse, hyper_1d_weighted
    if estimator_type == "mean_only":
        return mean_only_sparse, mean_only_weighted
    raise NameError("estimator_type %r is not defined in the reference" % (estimator_type,))


# --------------------------------------------------------------------------- mean-variance trend
def fit_mean_var(mean, var):
    """Quadratic fit of log var on log mean.  reference: estimator.py:84-93."""
    keep = (mean > 0) & (var > 0)
    return np.polyfit(np.log(mean[keep]), np.log(var[keep]), 2)


def residual_variance(mean, var, mv_fit):
    """exp(log var - trend(log mean)); NaN where mean<=0 or var<=0.  reference: estimator.py:103-111."""
    keep = (mean > 0) & (var > 0)
    out = np.full(mean.shape, np.nan)
    with np.errstate(invalid="ignore"):
        out[keep] = np.exp(np.log(var[keep]) - np.polyval(mv_fit, np.log(mean[keep])))
    return out


# --------------------------------------------------------------------------- 2D moments
def hyper_cov_sparse(X, n_obs, size_factor, q, idx1, idx2):
    """Covariance of gene pairs (idx1[k], idx2[k]).  reference: estimator.py:220-233."""
    idx1 = np.asarray(idx1)
    idx2 = np.asarray(idx2)
    same = idx1 == idx2
    w = np.sqrt(1.0 / size_factor ** 2)
    D = sp.diags(w)
    A = (D @ X[:, idx1]).tocsr()
    Bm = (D @ X[:, idx2]).tocsr()
    prod = _a1(A.multiply(Bm).sum(axis=0)) / n_obs
    if same.any():
        D2 = sp.diags(w ** 2)
        prod[same] = prod[same] - (1 - q) * _a1((D2 @ X[:, idx1[same]]).sum(axis=0)) / n_obs
    return prod - _a1(A.mean(axis=0)) * _a1(Bm.mean(axis=0))


def hyper_cov_weighted(x, y, weights, n_obs, q, inv_sf, inv_sf_sq):
    """Tuple form of the covariance.  reference: estimator.py:214-218."""
    m1 = (x * weights * inv_sf).sum(axis=0) / n_obs
    m2 = (y * weights * inv_sf).sum(axis=0) / n_obs
    mx = (x * y * weights * inv_sf_sq).sum(axis=0) / n_obs
    return mx - m1 * m2


def corr_from_cov(cov, var_1, var_2):
    """reference: estimator.py:273-292.  NOTE: mutates var_1 / var_2 in place like the reference
    (non-positive variances become NaN); entries with non-finite sqrt(var_1 var_2) keep the
    sentinel 5.0 which the clip then turns into 1.0."""
    if not isinstance(cov, np.ndarray):
        return cov / np.sqrt(var_1 * var_2)
    corr = np.full(cov.shape, 5.0)
    var_1[var_1 <= 0] = np.nan
    var_2[var_2 <= 0] = np.nan
    denom = np.sqrt(var_1 * var_2)
    ok = np.isfinite(denom)
    corr[ok] = cov[ok] / denom[ok]
    corr[corr > 1] = 1
    corr[corr < -1] = -1
    return corr


def corr_symmetric(X, n_obs, size_factor, q, var):
    """All-by-all correlation matrix.  reference: estimator.py:236-270."""
    G = X.shape[1]
    w = np.sqrt(1.0 / size_factor ** 2)
    Xw = (sp.diags(w) @ X).tocsr()
    prod = (Xw.T @ Xw).toarray() / Xw.shape[0]
    diag = np.arange(G)
    prod[diag, diag] = prod[diag, diag] - (1 - q) * _a1((sp.diags(w ** 2) @ X).sum(axis=0)) / n_obs
    mean = _a1(Xw.mean(axis=0))
    cov = prod - np.outer(mean, mean)
    # the reference NaNs copies of var (estimator.py:259-262) but then uses the ORIGINAL var in
    # the outer product (:263), so two negative variances give a finite denominator.
    with np.errstate(invalid="ignore"):
        denom = np.sqrt(np.outer(var, var))
    ok = np.isfinite(denom)
    corr = np.full(cov.shape, 5.0)
    corr[ok] = cov[ok] / denom[ok]
    near = (corr < 1.05) & (corr > -1.05)
    corr[near] = np.clip(corr[near], -1, 1)
    corr[(corr > 1) | (corr < -1)] = np.nan
    return corr
