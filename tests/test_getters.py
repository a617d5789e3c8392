"""Result getters and BH correction (host-side) against outputs of the unmodified
reference (tests/golden/getters.npz, written by tests/golden/make_golden.py:getters)."""
import numpy as np
import pandas as pd
import scipy.sparse as sp

import memento_b200 as memento
from helpers import load


def _adata(g):
    groups = [str(x) for x in g["groups"]]
    n_cells = g["n_cells"]
    G = g["mom"].shape[2]
    obs = pd.DataFrame({"stim": g["stim"], "cell": g["cell"]})
    var = pd.DataFrame(index=pd.Index(["g%d" % i for i in range(G)]))
    ad = memento.AnnDataLite(sp.csr_matrix((int(n_cells.sum()), G)), obs, var)
    pairs = [tuple(p) for p in g["pairs"]]
    ht = {k[3:]: g[k] for k in g.files if k.startswith("ht_")}
    ht["treatment"] = pd.DataFrame({"stim": [0, 0, 1, 1]}, index=groups)
    ad.uns["memento"] = {
        "groups": groups, "label_columns": ["stim", "cell"], "label_delimiter": "^",
        "group_cells": {gr: sp.csr_matrix((int(n), G)) for gr, n in zip(groups, n_cells)},
        "1d_moments": {gr: [g["mom"][i, k].copy() for k in range(3)] for i, gr in enumerate(groups)},
        "2d_moments": {"gene_pairs": pairs, **{gr: {"corr": g["corr"][i].copy()} for i, gr in enumerate(groups)}},
        "1d_ht": ht, "2d_ht": {k[4:]: g[k] for k in g.files if k.startswith("ht2_")},
        "mv_regressor": {gr: np.ones(3) for gr in groups + ["all"]}, "_b200": object()}
    return ad


def _same(a, b):
    np.testing.assert_allclose(np.asarray(a, dtype=float), b, rtol=1e-13, atol=0, equal_nan=True)


def test_getters_match_reference():
    g = load("getters.npz")
    ad = _adata(g)
    with np.errstate(divide="ignore", invalid="ignore"):
        m, v, counts = memento.get_1d_moments(ad)
        assert m.columns[1:].tolist() == [str(c) for c in g["m1_cols"]]
        _same(m.drop(columns="gene").values, g["m1_mean"])
        _same(v.drop(columns="gene").values, g["m1_var"])
        assert counts == {str(k): int(n) for k, n in zip(g["groups"], g["n_cells"])}
        for gb in ("stim", "cell", "ALL"):
            m, v = memento.get_1d_moments(ad, groupby=gb)
            assert m.columns[1:].tolist() == [str(c) for c in g["m1_%s_cols" % gb]]
            _same(m.drop(columns="gene").values, g["m1_%s_mean" % gb])
            _same(v.drop(columns="gene").values, g["m1_%s_var" % gb])
        c, _ = memento.get_2d_moments(ad)
        _same(c.drop(columns=["gene_1", "gene_2"]).values, g["m2_corr"])
        for gb in ("cell", "ALL"):
            c = memento.get_2d_moments(ad, groupby=gb)
            assert c.columns[2:].tolist() == [str(x) for x in g["m2_%s_cols" % gb]]
            _same(c.drop(columns=["gene_1", "gene_2"]).values, g["m2_%s_corr" % gb])
        # unlike the reference, the getter leaves uns['memento']['2d_moments'] untouched
        assert np.isnan(ad.uns["memento"]["2d_moments"]["sg^stim^A"]["corr"]).sum() == 1
    r1 = memento.get_1d_ht_result(ad)
    assert r1["gene"].tolist() == [str(x) for x in g["r1_gene"]] and r1["tx"].tolist() == [str(x) for x in g["r1_tx"]]
    _same(r1[["de_coef", "de_se", "de_pval", "dv_coef", "dv_se", "dv_pval"]].values, g["r1_vals"])
    _same(memento.get_2d_ht_result(ad)[["corr_coef", "corr_se", "corr_pval"]].values, g["r2_vals"])
    _same(memento.fdrcorrect(g["fdr_p"]), g["fdr_q"])
    df = memento.get_groups(ad)
    assert df.index.tolist() == [str(x) for x in g["groups"]] and df["cell"].tolist() == ["A", "B", "A", "B"]


def test_prepare_to_save():
    import pickle
    g = load("getters.npz")
    ad = _adata(g)
    memento.prepare_to_save(ad, keep=True)
    mem = ad.uns["memento"]
    assert "_b200" not in mem and "group_cells" not in mem
    assert all(isinstance(v, str) for v in mem["mv_regressor"].values())
    ad = _adata(g)
    memento.prepare_to_save(ad)
    assert ad.uns["memento"]["mv_regressor"] == {}
    pickle.dumps({k: v for k, v in ad.uns["memento"].items()})       # everything left is plain data
