"""On-disk formats either side of the hot path (host code, no GPU): 10x matrix.mtx directories in, results out."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))

from memento_b200 import io as mio            # noqa: E402
from memento_b200.anndata_lite import AnnDataLite  # noqa: E402


def _toy(n=40, g=25, seed=0):
    rng = np.random.default_rng(seed)
    X = sp.random(n, g, density=0.2, random_state=seed, data_rvs=lambda k: rng.integers(1, 30, k)).tocsr()
    X = sp.csr_matrix(X, dtype=np.float32)
    obs = pd.DataFrame(index=pd.Index(["AAAC%04d-1" % i for i in range(n)]))
    names = ["G%d" % i for i in range(g)]
    var = pd.DataFrame({"gene_ids": ["ENSG%05d" % i for i in range(g)]}, index=pd.Index(names))
    return AnnDataLite(X, obs, var)


@pytest.mark.parametrize("compress", [False, True])
def test_10x_mtx_round_trip(tmp_path, compress):
    ad = _toy()
    mio.write_10x_mtx(ad, str(tmp_path), compress=compress)
    back = mio.read_10x_mtx(str(tmp_path))
    assert sp.isspmatrix_csr(back.X) and back.X.dtype == np.float32 and back.X.has_sorted_indices
    assert back.shape == ad.shape
    assert (back.X != ad.X).nnz == 0
    assert back.obs.index.tolist() == ad.obs.index.tolist()
    assert back.var.index.tolist() == ad.var.index.tolist()
    assert back.var["gene_ids"].tolist() == ad.var["gene_ids"].tolist()
    assert mio.read_10x_mtx(str(tmp_path), var_names="gene_ids").var.index.tolist() == ad.var["gene_ids"].tolist()


def test_10x_mtx_rejects_inconsistent_directory(tmp_path):
    ad = _toy()
    mio.write_10x_mtx(ad, str(tmp_path))
    with open(os.path.join(str(tmp_path), "barcodes.tsv"), "a") as f:
        f.write("EXTRA-1\n")
    with pytest.raises(ValueError):
        mio.read_10x_mtx(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        mio.read_10x_mtx(os.path.join(str(tmp_path), "missing"))


def test_10x_mtx_duplicate_symbols_are_made_unique(tmp_path):
    ad = _toy(g=4)
    mio.write_10x_mtx(ad, str(tmp_path))
    with open(os.path.join(str(tmp_path), "features.tsv"), "w") as f:
        f.write("E0\tA\tGene Expression\nE1\tB\tGene Expression\nE2\tA\tGene Expression\nE3\tA\tGene Expression\n")
    assert mio.read_10x_mtx(str(tmp_path)).var.index.tolist() == ["A", "B", "A-1", "A-2"]


def test_results_round_trip(tmp_path):
    ad = _toy(g=6)
    groups = ["sg^ctrl^T", "sg^stim^T"]
    rng = np.random.default_rng(1)
    tr = pd.DataFrame({"stim": [0.0, 1.0]}, index=groups)
    cov = pd.DataFrame({"intercept": [1.0, 1.0]}, index=groups)
    ad.uns["memento"] = {
        "q_column": "q", "all_q": 0.07, "estimator_type": "hyper_relative", "filter_mean_thresh": 0.07, "num_bins": 30,
        "label_columns": ["stim", "cell"], "label_delimiter": "^", "groups": groups,
        "1d_moments": {g: [rng.random(6), rng.random(6), np.where(rng.random(6) < 0.3, np.nan, rng.random(6))] for g in groups},
        "mv_regressor": {g: rng.random(3) for g in groups + ["all"]},
        "gene_filter": {g: rng.random(9) < 0.5 for g in groups}, "gene_rv_filter": {g: rng.random(6) < 0.5 for g in groups},
        "overall_gene_filter": rng.random(9) < 0.5,
        "1d_ht": {"treatment": tr, "covariate": cov, **{k: rng.random(6) for k in
                                                        ("mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl")}},
        "2d_moments": {"gene_idx_1": np.array([0, 1]), "gene_idx_2": np.array([2, 3]),
                       **{g: {k: rng.random(2) for k in ("cov", "corr", "var_1", "var_2")} for g in groups}},
        "2d_ht": {"treatment": tr, "covariate": cov, "corr_coef": rng.random(2), "corr_se": rng.random(2),
                  "corr_asl": rng.random(2)},
        "_b200": object(), "group_cells": {g: object() for g in groups},
    }
    path = os.path.join(str(tmp_path), "res.npz")
    mio.save_results(ad, path)
    back = mio.load_results(path)
    mem = ad.uns["memento"]
    assert back["groups"] == groups and back["gene_list"] == ad.var.index.tolist() and back["all_q"] == 0.07
    assert "_b200" not in back and "group_cells" not in back
    for g in groups:
        for k in range(3):
            np.testing.assert_array_equal(back["1d_moments"][g][k], mem["1d_moments"][g][k])
        np.testing.assert_array_equal(back["gene_filter"][g], mem["gene_filter"][g])
        for k in ("cov", "corr", "var_1", "var_2"):
            np.testing.assert_array_equal(back["2d_moments"][g][k], mem["2d_moments"][g][k])
    for g in groups + ["all"]:
        np.testing.assert_array_equal(back["mv_regressor"][g], mem["mv_regressor"][g])
    np.testing.assert_array_equal(back["overall_gene_filter"], mem["overall_gene_filter"])
    for k in ("mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"):
        np.testing.assert_array_equal(back["1d_ht"][k], mem["1d_ht"][k])
    for k in ("corr_coef", "corr_se", "corr_asl"):
        np.testing.assert_array_equal(back["2d_ht"][k], mem["2d_ht"][k])
    pd.testing.assert_frame_equal(back["1d_ht"]["treatment"], tr)
    pd.testing.assert_frame_equal(back["2d_ht"]["covariate"], cov)
    np.testing.assert_array_equal(back["2d_moments"]["gene_idx_2"], [2, 3])
