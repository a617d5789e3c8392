"""Result getters (host-side pandas views over ``uns['memento']``; reference main.py:156-168,
:523-683, util.py:22-29).  Same outputs; ``get_groups`` avoids ``pd.to_numeric(errors='ignore')``, which
pandas 3 removed, ``prepare_to_save`` works (the reference's uses an undefined ``pkl``, main.py:683) and
``fdrcorrect`` needs no statsmodels."""
import itertools
import pickle

import numpy as np
import pandas as pd


def get_groups(adata):
    mem = adata.uns["memento"]
    rows = [g.split(mem["label_delimiter"])[1:] for g in mem["groups"]]
    df = pd.DataFrame(rows, index=mem["groups"], columns=mem["label_columns"])
    for col in df.columns:
        try:
            df[col] = pd.to_numeric(df[col])
        except (ValueError, TypeError):
            pass
    return df


def _groupby_keys(adata, groupby):
    """Keys a group name is matched against (substring match, as the reference does): the distinct values of an
    obs column, or every group for 'ALL'."""
    if groupby == "ALL":
        return ["sg"]
    return list(adata.obs[groupby].astype(str).drop_duplicates().values)


def _weighted_by_cells(columns, counts, keys, prefix):
    """Cell-count weighted mean over the groups matching each key; NaN entries do not count (reference
    main.py:553-582, :610-632).  ``columns``: {group: (values, valid mask)}."""
    out = {}
    for key in keys:
        num, den = 0.0, 0.0
        for group, (val, valid) in columns.items():
            if key in group:
                num = num + np.where(np.isnan(val), 0.0, val) * counts[group]
                den = den + valid * counts[group]
        with np.errstate(divide="ignore", invalid="ignore"):
            out[prefix + key] = num / den
    return out


def get_1d_moments(adata, groupby=None):
    """Log mean and log residual variance per group; with ``groupby`` their cell-count weighted means over the
    groups whose name contains each value of that obs column ('ALL': over all groups).  reference main.py:523-582."""
    mem = adata.uns["memento"]
    genes = adata.var.index.tolist()
    counts = {k: v.shape[0] for k, v in mem["group_cells"].items()}
    logs = {}
    with np.errstate(divide="ignore", invalid="ignore"):
        for group, val in mem["1d_moments"].items():
            if group != "all":
                logs[group] = (np.log(val[0]), np.log(val[2]), val[0] > 0, val[2] > 0)
    if groupby is None:
        mean_df = pd.DataFrame({"gene": genes, **{g: v[0] for g, v in logs.items()}})
        var_df = pd.DataFrame({"gene": genes, **{g: v[1] for g, v in logs.items()}})
        return mean_df, var_df, counts
    keys = _groupby_keys(adata, groupby)
    prefix = groupby + "_"
    mean_df = pd.DataFrame({"gene": genes, **_weighted_by_cells({g: (v[0], v[2]) for g, v in logs.items()}, counts, keys, prefix)})
    var_df = pd.DataFrame({"gene": genes, **_weighted_by_cells({g: (v[1], v[3]) for g, v in logs.items()}, counts, keys, prefix)})
    return mean_df, var_df


def get_2d_moments(adata, groupby=None):
    """Correlation of every gene pair per group (or weighted over groups, as get_1d_moments).  reference main.py:585-632."""
    mem = adata.uns["memento"]
    pairs = pd.DataFrame(mem["2d_moments"]["gene_pairs"], columns=["gene_1", "gene_2"])
    counts = {k: v.shape[0] for k, v in mem["group_cells"].items()}
    corr = {g: v["corr"] for g, v in mem["2d_moments"].items() if isinstance(g, str) and "sg^" in g}
    if groupby is None:
        return pd.concat([pairs, pd.DataFrame(corr)], axis=1), counts
    cols = {g: (c, ~np.isnan(c)) for g, c in corr.items()}
    return pd.concat([pairs, pd.DataFrame(_weighted_by_cells(cols, counts, _groupby_keys(adata, groupby), groupby + "_"))], axis=1)


def get_1d_ht_result(adata):
    ht = adata.uns["memento"]["1d_ht"]
    if "treatment_for_gene" in ht:
        df = pd.concat([pd.DataFrame(itertools.product([g], ht["treatment_for_gene"][g]), columns=["gene", "tx"])
                        for g in adata.var.index])
    else:
        df = pd.DataFrame(itertools.product(adata.var.index, ht["treatment"].columns), columns=["gene", "tx"])
    df["de_coef"], df["de_se"], df["de_pval"] = ht["mean_coef"], ht["mean_se"], ht["mean_asl"]
    df["dv_coef"], df["dv_se"], df["dv_pval"] = ht["var_coef"], ht["var_se"], ht["var_asl"]
    return df


def get_2d_ht_result(adata):
    """reference main.py:658-670."""
    mem = adata.uns["memento"]
    df = pd.DataFrame(mem["2d_moments"]["gene_pairs"], columns=["gene_1", "gene_2"])
    df["corr_coef"], df["corr_se"], df["corr_pval"] = mem["2d_ht"]["corr_coef"], mem["2d_ht"]["corr_se"], mem["2d_ht"]["corr_asl"]
    return df


def prepare_to_save(adata, keep=False):
    """Make ``uns['memento']`` writable with scanpy / h5ad (reference main.py:673-683): the per-group mean-variance
    fits are dropped, or pickled to strings with ``keep=True``; the device state of this implementation and the lazy
    per-group matrix views are dropped as well."""
    mem = adata.uns["memento"]
    mem.pop("_b200", None)
    mem.pop("group_cells", None)
    for group in list(mem.get("mv_regressor", {})):
        if keep:
            mem["mv_regressor"][group] = str(pickle.dumps(mem["mv_regressor"][group]))
        else:
            del mem["mv_regressor"][group]


def fdrcorrect(pvals):
    """Benjamini-Hochberg adjusted p-values; NaN inputs get 1 (reference util.py:22-29, which calls statsmodels'
    fdrcorrection on the non-NaN entries of an array of ones)."""
    p = np.asarray(pvals, dtype=float)
    out = np.ones(p.shape)
    ok = ~np.isnan(p)
    m = int(ok.sum())
    if m == 0:
        return out
    order = np.argsort(p[ok])
    ranked = p[ok][order] * m / np.arange(1, m + 1)
    adj = np.minimum.accumulate(ranked[::-1])[::-1]
    res = np.empty(m)
    res[order] = np.minimum(adj, 1.0)
    out[ok] = res
    return out
