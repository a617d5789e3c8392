"""Oracle: per-gene meta-regression and ASL p-values (TEST INFRASTRUCTURE, see oracle/__init__.py).

numpy restatement of reference ``memento/hypothesis_test.py``.  Random imputation and the
hierarchical replicate bootstrap draw from the GLOBAL ``np.random`` state in the same order as
the reference, so seeded runs agree with it exactly.
"""
import warnings

import numpy as np
import scipy.stats as stats
from sklearn.linear_model import LinearRegression

from . import moments, resample

GEV_MAX_EXTREME = 10        # hypothesis_test.py:90
GEV_TAIL_SIZES = tuple(range(300, 50, -30))  # N_exec = 300, 270, ..., 60  (:102-116)
GEV_KS_ALPHA = 0.05         # :110


# --------------------------------------------------------------------------- imputation
def fill_invalid(val, record=None):
    """Replace entries <=0 or NaN by random picks among the valid ones (in place); None when
    nothing is valid.  reference: hypothesis_test.py:23-33.

    ``np.random.choice(valid, n)`` is ``valid[np.random.randint(0, len(valid), n)]`` on the same
    stream (checked in tests); drawing the indices explicitly lets a test record which replicate
    every invalid entry was filled from (``record``: dict that receives ``src``, the source
    replicate index per entry, -1 where the entry was kept)."""
    with np.errstate(invalid="ignore"):
        bad = ~(val > 0)
    n_bad = bad.sum()
    if n_bad == val.shape[0]:
        return None
    valid_pos = np.flatnonzero(~bad)
    pick = np.random.randint(0, valid_pos.shape[0], n_bad)
    if record is not None:
        src = np.full(val.shape[0], -1, dtype=np.int32)
        src[bad] = valid_pos[pick]
        record["src"] = src
    val[bad] = val[valid_pos[pick]]
    return val


def fill_nan(val):
    """NaN-only imputation used for correlations.  reference: hypothesis_test.py:35-40."""
    bad = np.isnan(val)
    val[bad] = np.random.choice(val[~bad], bad.sum())
    return val


# --------------------------------------------------------------------------- ASL
def _gev_tail(tail, n_total, stat_abs, side):
    """One GEV tail attempt ladder.  reference: hypothesis_test.py:101-116 (left), :121-134 (right).
    Returns the tail ASL or None when no tail size passes the KS check."""
    for n_exec in GEV_TAIL_SIZES:
        data = tail[:n_exec] if side == "left" else tail[-n_exec:]
        params = stats.genextreme.fit(data)
        _, ks_p = stats.kstest(data, "genextreme", args=params)
        if ks_p > GEV_KS_ALPHA:
            if side == "left":
                p = stats.genextreme.cdf(-stat_abs, *params)
            else:
                p = stats.genextreme.sf(stat_abs, *params)
            return (n_exec / n_total) * p
    return None


def compute_asl(x, resampling="bootstrap", approx=False):
    """Achieved significance level of x[0] against the resampled x[1:].
    reference: hypothesis_test.py:57-141."""
    if np.all(x == x.mean()):
        return np.nan
    null = x[1:] - x[0] if resampling == "bootstrap" else x[1:]
    null = null[np.isfinite(null)]
    stat = x[0]
    if approx:
        mu, sd = stats.norm.fit(null)
        a = np.abs(stat)
        return stats.norm.sf(a, mu, sd) + stats.norm.cdf(-a, mu, sd)
    a = abs(stat)
    # reference :85-88; for stat == 0 both branches reduce to (null > 0) + (null < 0)
    extreme = (null > a).sum() + (null < -a).sum()
    upper = (extreme + 1) / (null.shape[0] + 1)
    if extreme > GEV_MAX_EXTREME:
        return upper
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            dist = np.sort(null)
            left = _gev_tail(dist, dist.shape[0], a, "left")
            if left is None:
                return upper
            right = _gev_tail(dist, dist.shape[0], a, "right")
            if right is None:
                return upper
            return right + left
        except Exception:
            return upper


# --------------------------------------------------------------------------- regression
def cross_coef(A, Bm, w):
    """Weighted marginal slope of each column of Bm on each column of A.
    reference: hypothesis_test.py:218-228."""
    A0 = A - np.average(A, axis=0, weights=w)
    B0 = Bm - np.average(Bm, axis=0, weights=w)
    ss = np.average(A0 ** 2, axis=0, weights=w)
    return (A0.T * w).dot(B0) / w.sum() / ss[:, None]


def cross_coef_resampled(A, Bm, w):
    """Per-replicate weighted slope for the hierarchical bootstrap (A: R x B x T, Bm: R x B,
    w: R x B).  reference: hypothesis_test.py:231-239."""
    B0 = Bm - np.average(Bm, axis=0, weights=w)
    wsum = w.sum(axis=0)
    A0 = A - (A * w[:, :, None]).sum(axis=0) / wsum[:, None]
    ss = (A0 ** 2 * w[:, :, None]).sum(axis=0) / wsum[:, None]
    return np.einsum("ijk,ij->jk", A0 * w[:, :, None], B0).T / wsum / ss.T


def _residualise(covariate, Y, w):
    """Y minus its weighted least-squares projection on [1, covariate].
    reference: hypothesis_test.py:269-271 (sklearn LinearRegression, fit_intercept=True)."""
    return Y - LinearRegression(n_jobs=1).fit(covariate, Y, w).predict(covariate)


def _replicate_assignment(num_rep, num_boot):
    """reference: hypothesis_test.py:275-278 (global RNG; column 0 is the identity)."""
    rep = np.random.choice(num_rep, size=(num_rep, num_boot))
    rep[:, 0] = np.arange(num_rep)
    it = np.random.choice(num_boot, (num_rep, num_boot)) + 1
    it[:, 0] = 0
    return rep, it


def regress(covariate, treatment, boots, n_cells, resample_rep=False, record=None, **asl_kwargs):
    """Shared body of ``_regress_1d`` / ``_regress_2d``.  ``boots`` is a list of (R x (B+1))
    arrays (1D: [log mean, log res-var]; 2D: [corr]).  Returns one (coef0, se, asl) triple per
    array, or None when every bootstrap column was dropped.
    reference: hypothesis_test.py:242-300, :367-414."""
    keep = np.ones(boots[0].shape[1], dtype=bool)
    for b in boots:
        keep &= ~np.any(~np.isfinite(b), axis=0)
    boots = [b[:, keep] for b in boots]
    num_boot = boots[0].shape[1] - 1
    num_rep = boots[0].shape[0]
    if boots[-1].shape[1] == 0:
        return None
    if (treatment == 1).mean() == 1:  # one-sample test: weighted average over groups
        coefs = [np.average(b, axis=0, weights=n_cells).reshape(1, -1) for b in boots]
    else:
        tildes = [_residualise(covariate, b, n_cells) for b in boots]
        t_tilde = _residualise(covariate, treatment, n_cells)
        if resample_rep:
            rep, it = _replicate_assignment(num_rep, num_boot)
            if record is not None:
                record["rep_assign"], record["iter_assign"] = rep, it
            t_res = t_tilde[rep]
            w_res = n_cells[rep]
            coefs = [cross_coef_resampled(t_res, bt[(rep, it)], w_res) for bt in tildes]
        else:
            coefs = [cross_coef(t_tilde, bt, n_cells) for bt in tildes]
    out = []
    for c in coefs:
        asl = np.apply_along_axis(lambda row: compute_asl(row, **asl_kwargs), 1, c)
        se = np.nanstd(c[:, 1:], axis=1)
        out.append((c[:, 0], se, asl))
    return out


# --------------------------------------------------------------------------- per-gene drivers
def ht_1d_gene(true_mean, true_res_var, cells, approx_sf, covariate, treatment, n_cells,
               num_boot, mv_fit, q, weighted_estimator, return_boot=False, recorder=None, **kwargs):
    """All groups of one gene: bootstrap, residual variance, imputation, regression.
    reference: hypothesis_test.py:144-215.  Returns the 6-tuple the reference returns."""
    R = treatment.shape[0]
    good = np.zeros(R, dtype=bool)
    boot_mean = np.full((R, num_boot + 1), np.nan)
    boot_var = np.full((R, num_boot + 1), np.nan)
    for r in range(len(true_mean)):
        if np.isnan(true_mean[r]) or np.isnan(true_res_var[r]) or true_mean[r] == 0 \
                or true_res_var[r] < 0:
            continue
        with np.errstate(divide="ignore"):
            boot_mean[r, 0], boot_var[r, 0] = np.log(true_mean[r]), np.log(true_res_var[r])
        table = resample.unique_table(cells[r], approx_sf[r])
        mean, var = resample.bootstrap_1d(cells[r], approx_sf[r], q[r], weighted_estimator, num_boot,
                                          precomputed=table)
        res_var = moments.residual_variance(mean, var, mv_fit[r])
        rec_m, rec_v = {}, {}
        filled_mean = fill_invalid(mean, rec_m)
        filled_var = fill_invalid(res_var, rec_v)
        if recorder is not None:
            recorder.append({"group": r, "inv_sf": table[0].reshape(-1), "values": table[2][:, 0],
                             "mult": table[3], "n_cells": cells[r].shape[0],
                             "src_mean": rec_m.get("src"), "src_rv": rec_v.get("src")})
        if filled_mean is None or filled_var is None:
            continue
        boot_mean[r, 1:] = np.log(filled_mean)
        boot_var[r, 1:] = np.log(filled_var)
        good[r] = True
    if return_boot:
        return good, boot_mean, boot_var
    if good.sum() == 0:
        return (np.nan,) * 6
    reg_record = {} if recorder is not None else None
    res = regress(covariate[good, :], treatment[good, :], [boot_mean[good, :], boot_var[good, :]],
                  n_cells[good], record=reg_record, **kwargs)
    if recorder is not None and reg_record:
        recorder.append({"group": -1, **reg_record})
    if res is None:
        # the reference returns a 5-list here and its caller then fails to unpack 6 values
        # (hypothesis_test.py:260, main.py:403); the oracle reports all-NaN instead.
        nan = np.full(treatment.shape[1], np.nan)
        return (nan,) * 6
    (mc, mse, masl), (vc, vse, vasl) = res
    return mc, mse, masl, vc, vse, vasl


def ht_2d_pair(true_corr, cells, approx_sf, covariate, treatment, n_cells, num_boot, q,
               weighted_estimator, weighted_cov, **kwargs):
    """All groups of one gene pair.  reference: hypothesis_test.py:303-364."""
    R = treatment.shape[0]
    good = np.zeros(R, dtype=bool)
    boot_corr = np.full((R, num_boot + 1), np.nan)
    for r in range(R):
        if np.isnan(true_corr[r]) or np.abs(true_corr[r]) == 1:
            continue
        boot_corr[r, 0] = true_corr[r]
        cov, v1, v2 = resample.bootstrap_2d(cells[r], approx_sf[r], q[r], weighted_estimator,
                                            weighted_cov, int(num_boot))
        corr = moments.corr_from_cov(cov, v1, v2)
        vals = fill_nan(corr)
        if np.all(np.isnan(vals)):
            continue
        good[r] = True
        boot_corr[r, 1:] = vals
    if good.sum() == 0:
        return np.nan, np.nan, np.nan
    res = regress(covariate[good, :], treatment[good, :], [boot_corr[good, :]], n_cells[good], **kwargs)
    if res is None:
        nan = np.full(treatment.shape[1], np.nan)
        return nan, nan, nan
    return res[0]
