"""Oracle: the six public calls of the hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

numpy restatement of reference ``memento/main.py`` for ``setup_memento``, ``create_groups``,
``compute_1d_moments``, ``ht_1d_moments``, ``compute_2d_moments`` and ``ht_2d_moments``, writing
the same ``adata.uns['memento']`` schema.  ``adata`` is any object with ``X`` (scipy CSR),
``obs`` / ``var`` (pandas), ``uns`` (dict), ``shape``, ``copy()`` and
``_inplace_subset_var(mask)`` -- the only members the reference touches.

``num_cpus > 1`` fans the per-gene work out over a fork pool (the reference uses joblib/loky,
main.py:397, :501); with ``num_cpus == 1`` the genes run in order in this process, which keeps
the global-RNG call sequence identical to the reference's sequential joblib path.
"""
import multiprocessing as mp

import numpy as np
import scipy.sparse as sp

from . import moments, testing


def _mem(adata):
    return adata.uns["memento"]


# --------------------------------------------------------------------------- setup_memento
def setup_memento(adata, q_column, filter_mean_thresh=0.07, trim_percent=0.1, shrinkage=0.5,
                  num_bins=30, estimator_type="hyper_relative"):
    """reference: main.py:26-91."""
    assert adata.obs[q_column].max() < 1
    assert isinstance(adata.X, sp.csr_matrix)
    X = adata.X
    n = adata.shape[0]
    mem = adata.uns["memento"] = {}
    mem["q_column"] = q_column
    mem["all_q"] = adata.obs[q_column].values.mean()
    mem["estimator_type"] = estimator_type
    mem["filter_mean_thresh"] = filter_mean_thresh
    mem["num_bins"] = num_bins

    naive = moments.row_totals(X)                                            # main.py:55-59
    m, v = moments.hyper_1d_sparse(X, n, mem["all_q"], naive)               # :62-66
    m[np.asarray(X.mean(axis=0)).reshape(-1) < filter_mean_thresh] = 0      # :67
    rv = moments.residual_variance(m, v, moments.fit_mean_var(m, v))        # :68
    ulim = np.quantile(rv[np.isfinite(rv)], trim_percent)                   # :71
    rv[~np.isfinite(rv)] = np.inf
    mask = rv < ulim
    mem["least_variable_genes"] = adata.var.index[mask].tolist()
    sf = moments.masked_size_factor(X, mask, shrinkage)                     # :78-82
    adata.obs["memento_size_factor"] = sf
    sparse_fn, _ = moments.estimator_1d(estimator_type)
    mem["all_1d_moments"] = sparse_fn(X, n, mem["all_q"], sf)               # :86-91


# --------------------------------------------------------------------------- create_groups
def create_groups(adata, label_columns, label_delimiter="^"):
    """reference: main.py:94-135 and util.py:8-13."""
    mem = _mem(adata)
    label = "sg" + label_delimiter
    for i, col in enumerate(label_columns):
        label = label + adata.obs[col].astype(str)
        if i != len(label_columns) - 1:
            label = label + label_delimiter
    adata.obs["memento_group"] = label
    mem["label_columns"] = label_columns
    mem["label_delimiter"] = label_delimiter
    mem["groups"] = adata.obs["memento_group"].drop_duplicates().tolist()
    mem["q"] = adata.obs[mem["q_column"]].values
    gcol = adata.obs["memento_group"].values
    mem["group_cells"] = {g: adata.X[gcol == g, :].tocsc() for g in mem["groups"]}
    mem["group_q"] = {g: mem["q"][gcol == g].mean() for g in mem["groups"]}


def _bin_size_factor(adata):
    """reference: main.py:138-153."""
    mem = _mem(adata)
    sf = adata.obs["memento_size_factor"].values
    approx, _ = moments.bin_size_factor(sf, mem["num_bins"])
    gcol = adata.obs["memento_group"].values
    mem["all_approx_size_factor"] = approx
    mem["approx_size_factor"] = {g: approx[gcol == g] for g in mem["groups"]}
    mem["size_factor"] = {g: sf[gcol == g] for g in mem["groups"]}


# --------------------------------------------------------------------------- compute_1d_moments
def compute_1d_moments(adata, min_perc_group=0.7, filter_genes=True, gene_list=None):
    """reference: main.py:171-274."""
    mem = _mem(adata)
    if "size_factor" not in mem:
        _bin_size_factor(adata)
    groups = mem["groups"]
    sparse_fn, _ = moments.estimator_1d(mem["estimator_type"])
    mem["1d_moments"] = {
        g: sparse_fn(mem["group_cells"][g], mem["group_cells"][g].shape[0], mem["group_q"][g],
                     mem["size_factor"][g]) for g in groups}
    mem["gene_filter"], mem["gene_rv_filter"] = {}, {}
    for g in groups:                                                         # main.py:199-207
        cells = mem["group_cells"][g]
        obs_mean = np.asarray(cells.mean(axis=0)).reshape(-1)
        mem["gene_filter"][g] = (obs_mean > mem["filter_mean_thresh"]) & (mem["1d_moments"][g][1] > 0)
        obs_max = np.asarray(cells.max(axis=0).todense()).reshape(-1)
        mem["gene_rv_filter"][g] = obs_max >= 2
    rate = np.vstack([mem["gene_filter"][g] for g in groups]).mean(axis=0)   # :210-212
    overall = rate > min_perc_group
    mem["overall_gene_filter"] = overall
    mem["gene_list"] = adata.var.index[overall].tolist()
    if filter_genes:                                                         # :219-229
        mem["group_cells"] = {g: mem["group_cells"][g][:, overall] for g in groups}
        mem["1d_moments"] = {g: [mem["1d_moments"][g][0][overall], mem["1d_moments"][g][1][overall]]
                             for g in groups}
        mem["gene_rv_filter"] = {g: mem["gene_rv_filter"][g][overall] for g in groups}
        adata._inplace_subset_var(overall)
    mean_cat = np.concatenate([mem["1d_moments"][g][0][mem["gene_rv_filter"][g]] for g in groups])
    var_cat = np.concatenate([mem["1d_moments"][g][1][mem["gene_rv_filter"][g]] for g in groups])
    pooled = moments.fit_mean_var(mean_cat, var_cat)                         # :239
    mem["mv_regressor"] = {"all": pooled}
    for g in groups:                                                         # :242-245 (same pooled fit)
        mem["mv_regressor"][g] = moments.fit_mean_var(mean_cat, var_cat)
    for g in groups:                                                         # :248-255
        mem["1d_moments"][g].append(
            moments.residual_variance(mem["1d_moments"][g][0], mem["1d_moments"][g][1],
                                      mem["mv_regressor"][g]))
    if gene_list is not None:                                                # :258-271
        given = np.isin(adata.var.index.values, gene_list)
        mem["group_cells"] = {g: mem["group_cells"][g][:, given] for g in groups}
        mem["1d_moments"] = {g: [mem["1d_moments"][g][k][given] for k in range(3)] for g in groups}
        adata._inplace_subset_var(given)


# --------------------------------------------------------------------------- ht_1d_moments
def _run(tasks, fn, num_cpus):
    if num_cpus <= 1 or len(tasks) <= 1:
        return [fn(t) for t in tasks]
    global _POOL_FN
    _POOL_FN = fn
    with mp.get_context("fork").Pool(num_cpus) as pool:
        return pool.map(_pool_call, tasks, chunksize=max(1, len(tasks) // (8 * num_cpus)))


_POOL_FN = None


def _pool_call(t):
    return _POOL_FN(t)


def ht_1d_moments(adata, covariate, treatment, treatment_for_gene=None, num_boot=10000,
                  num_cpus=1, gene_subset=None, **kwargs):
    """reference: main.py:341-415.  ``gene_subset`` (oracle-only) restricts the run to the given
    gene positions; the output arrays then cover only those genes (used for bounded timing)."""
    mem = _mem(adata)
    groups = mem["groups"]
    G = adata.shape[1]
    n_cells = np.array([mem["group_cells"][g].shape[0] for g in groups])
    _, weighted_fn = moments.estimator_1d(mem["estimator_type"])
    genes = list(range(G)) if gene_subset is None else list(gene_subset)
    names = adata.var.index

    def one(idx):
        tr = treatment.values if treatment_for_gene is None else \
            treatment[treatment_for_gene[names[idx]]].values
        return testing.ht_1d_gene(
            true_mean=[mem["1d_moments"][g][0][idx] for g in groups],
            true_res_var=[mem["1d_moments"][g][2][idx] for g in groups],
            cells=[mem["group_cells"][g][:, idx] for g in groups],
            approx_sf=[mem["approx_size_factor"][g] for g in groups],
            covariate=covariate.values, treatment=tr, n_cells=n_cells, num_boot=num_boot,
            mv_fit=[mem["mv_regressor"][g] for g in groups],
            q=[mem["group_q"][g] for g in groups], weighted_estimator=weighted_fn, **kwargs)

    results = _run(genes, one, num_cpus)
    nts = [treatment.shape[1] if treatment_for_gene is None else len(treatment_for_gene[names[i]])
           for i in genes]
    out = [np.full(sum(nts), np.nan) for _ in range(6)]
    ci = 0
    for nt, res in zip(nts, results):
        for k in range(6):
            out[k][ci:ci + nt] = res[k]
        ci += nt
    ht = {}
    if treatment_for_gene is not None:
        ht["treatment_for_gene"] = treatment_for_gene
    ht["treatment"], ht["covariate"] = treatment, covariate
    for k, name in enumerate(["mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"]):
        ht[name] = out[k]
    mem["1d_ht"] = ht


# --------------------------------------------------------------------------- 2D
def compute_2d_moments(adata, gene_pairs):
    """reference: main.py:293-338."""
    mem = _mem(adata)
    if "size_factor" not in mem:
        _bin_size_factor(adata)
    pos = dict(zip(adata.var.index.values, np.arange(adata.var.shape[0])))
    i1 = np.array([pos[a] for a, _ in gene_pairs], dtype=int)
    i2 = np.array([pos[b] for _, b in gene_pairs], dtype=int)
    out = {"gene_pairs": gene_pairs, "gene_idx_1": i1, "gene_idx_2": i2}
    for g in mem["groups"]:
        cells = mem["group_cells"][g]
        cov = moments.hyper_cov_sparse(cells, cells.shape[0], mem["size_factor"][g], mem["group_q"][g], i1, i2)
        v1 = mem["1d_moments"][g][1][i1]
        v2 = mem["1d_moments"][g][1][i2]
        out[g] = {"cov": cov, "corr": moments.corr_from_cov(cov, v1, v2), "var_1": v1, "var_2": v2}
    mem["2d_moments"] = out


def ht_2d_moments(adata, covariate, treatment, num_boot=10000, num_cpus=1, **kwargs):
    """reference: main.py:418-520 (``treatment_for_gene`` is broken there, main.py:492, and is
    not restated)."""
    mem = _mem(adata)
    groups = mem["groups"]
    n_cells = np.array([mem["group_cells"][g].shape[0] for g in groups])
    i1, i2 = mem["2d_moments"]["gene_idx_1"], mem["2d_moments"]["gene_idx_2"]
    _, weighted_fn = moments.estimator_1d(mem["estimator_type"])
    out = [np.full(i1.shape[0], np.nan) for _ in range(3)]
    tasks, first, dup = [], {}, {}
    for k in range(i1.shape[0]):                                             # main.py:467-482
        a, b = i1[k], i2[k]
        if a == b:
            continue
        key = frozenset({a, b})
        if key in dup:
            dup[key].append(k)
            continue
        dup[key] = [k]
        first[key] = k
        tasks.append((a, b, k))

    def one(task):
        a, b, k = task
        return testing.ht_2d_pair(
            true_corr=[mem["2d_moments"][g]["corr"][k] for g in groups],
            cells=[mem["group_cells"][g][:, [a, b]] for g in groups],
            approx_sf=[mem["approx_size_factor"][g] for g in groups],
            covariate=covariate.values, treatment=treatment.values, n_cells=n_cells,
            num_boot=num_boot, q=[mem["group_q"][g] for g in groups],
            weighted_estimator=weighted_fn, weighted_cov=moments.hyper_cov_weighted, **kwargs)

    results = _run(tasks, one, num_cpus)
    for (a, b, k), res in zip(tasks, results):
        for kk in dup[frozenset({a, b})]:
            for j in range(3):
                out[j][kk] = np.asarray(res[j]).reshape(-1)[0]               # main.py:509 (T == 1)
    mem["2d_ht"] = {"treatment": treatment, "covariate": covariate,
                    "corr_coef": out[0], "corr_se": out[1], "corr_asl": out[2]}
