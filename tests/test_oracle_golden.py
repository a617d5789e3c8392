"""The oracle (numpy restatement) against the golden outputs of the unmodified reference.

CPU only.  Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py from
/root/reference.  Tolerances are float64 round-off (1e-10) except the GEV-tail ASL, which goes
through scipy's Nelder-Mead in both and is compared exactly as well (same scipy, same inputs).
"""
import numpy as np
import pytest

from helpers import assert_close, golden_adata, load
from memento_b200 import synth
from oracle import moments, pipeline, resample, testing

RT = 1e-10


@pytest.fixture(scope="module")
def st():
    return load("stages.npz")


@pytest.fixture(scope="module")
def prepared(st):
    ad = golden_adata(st)
    pipeline.setup_memento(ad, "q")
    pipeline.create_groups(ad, ["stim", "cell"])
    pipeline.compute_1d_moments(ad, min_perc_group=0.7)
    return ad


def test_size_factor_and_naive_moments(st):
    ad = golden_adata(st)
    naive = moments.row_totals(ad.X)
    assert_close(naive, st["naive_sf"], 0, what="naive sf")
    m, v = moments.hyper_1d_sparse(ad.X, ad.shape[0], 0.07, naive)
    assert_close(m, st["naive_mean"], RT)
    assert_close(v, st["naive_var"], RT, atol=1e-18)


def test_setup_memento(st):
    ad = golden_adata(st)
    pipeline.setup_memento(ad, "q")
    mem = ad.uns["memento"]
    assert_close(ad.obs["memento_size_factor"].values, st["size_factor"], RT)
    assert mem["least_variable_genes"] == st["least_variable_genes"].tolist()
    assert_close(mem["all_1d_moments"][0], st["all_mean"], RT)
    assert_close(mem["all_1d_moments"][1], st["all_var"], RT, atol=1e-18)
    assert_close(mem["all_q"], st["all_q"], RT)


def test_groups_and_1d_moments(st, prepared):
    mem = prepared.uns["memento"]
    assert mem["groups"] == st["groups"].tolist()
    assert_close([mem["group_q"][g] for g in mem["groups"]], st["group_q"], RT)
    assert [mem["group_cells"][g].shape[0] for g in mem["groups"]] == st["group_ncells"].tolist()
    assert_close(mem["all_approx_size_factor"], st["approx_sf"], RT)
    assert np.array_equal(mem["overall_gene_filter"], st["overall_gene_filter"])
    assert mem["gene_list"] == st["gene_list"].tolist()
    assert_close(mem["mv_regressor"]["all"], st["mv_regressor"], 1e-9)
    for gi, g in enumerate(mem["groups"]):
        assert_close(mem["1d_moments"][g][0], st["m1d_mean_%d" % gi], RT)
        assert_close(mem["1d_moments"][g][1], st["m1d_var_%d" % gi], RT, atol=1e-18)
        assert_close(mem["1d_moments"][g][2], st["m1d_rv_%d" % gi], 1e-9)
        assert np.array_equal(mem["gene_filter"][g], st["gene_filter_%d" % gi])
        assert np.array_equal(mem["gene_rv_filter"][g], st["gene_rv_filter_%d" % gi])


def test_unique_tables_and_bootstrap(st, prepared):
    mem = prepared.uns["memento"]
    for k, (gene, gi) in enumerate(st["table_picks"]):
        g = mem["groups"][gi]
        col = mem["group_cells"][g][:, gene]
        sf = mem["approx_size_factor"][g]
        np.random.seed(100 + k)
        inv_sf, inv_sf_sq, vals, mult = resample.unique_table(col, sf)
        assert_close(inv_sf, st["tab%d_inv_sf" % k], RT)
        assert_close(vals, st["tab%d_expr" % k], 0)
        assert np.array_equal(mult, st["tab%d_counts" % k])
        assert mult.sum() == col.shape[0]
        W = resample.draw_counts(col.shape[0], mult, 64)
        assert np.array_equal(W, st["tab%d_W" % k])
        np.random.seed(100 + k)
        mean, var = resample.bootstrap_1d(col, sf, mem["group_q"][g], moments.hyper_1d_weighted, 64)
        assert_close(mean, st["tab%d_boot_mean" % k], RT)
        assert_close(var, st["tab%d_boot_var" % k], RT, atol=1e-18)


def test_2d_moments(st, prepared):
    ad = prepared.copy()
    names = ad.var.index.tolist()
    pairs = [(names[i], names[j]) for i, j in zip(st["pairs_idx1"], st["pairs_idx2"])]
    pipeline.compute_2d_moments(ad, pairs)
    mem = ad.uns["memento"]
    for gi, g in enumerate(mem["groups"]):
        d = mem["2d_moments"][g]
        assert_close(d["cov"], st["m2d_cov_%d" % gi], RT, atol=1e-18)
        assert_close(d["corr"], st["m2d_corr_%d" % gi], 1e-9)
        assert_close(d["var_1"], st["m2d_var1_%d" % gi], RT, atol=1e-18)
    g = mem["groups"][1]
    cm = moments.corr_symmetric(mem["group_cells"][g], mem["group_cells"][g].shape[0],
                                mem["size_factor"][g], mem["group_q"][g], mem["1d_moments"][g][1])
    assert_close(cm, st["corr_matrix_g1"], 1e-9, atol=1e-12)
    cols = mem["group_cells"][g][:, [0, 3]]
    np.random.seed(321)
    inv_sf, _, vals, mult = resample.unique_table(cols, mem["approx_size_factor"][g])
    assert_close(inv_sf, st["tab2d_inv_sf"], RT)
    assert_close(vals, st["tab2d_expr"], 0)
    assert np.array_equal(mult, st["tab2d_counts"])
    np.random.seed(321)
    cov, v1, v2 = resample.bootstrap_2d(cols, mem["approx_size_factor"][g], mem["group_q"][g],
                                        moments.hyper_1d_weighted, moments.hyper_cov_weighted, 64)
    assert_close(cov, st["tab2d_cov"], RT, atol=1e-18)
    assert_close(v1, st["tab2d_var1"], RT, atol=1e-18)
    assert_close(v2, st["tab2d_var2"], RT, atol=1e-18)


@pytest.mark.parametrize("variant,kw", [
    ("default", dict(resampling="bootstrap")),
    ("approx", dict(resampling="bootstrap", approx=True)),
    ("resample_rep", dict(resampling="bootstrap", approx=True, resample_rep=True)),
])
def test_ht_1d(prepared, variant, kw):
    ht = load("ht1d.npz")
    ad = prepared.copy()
    cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
    assert_close(cov.values, ht["covariate"], 0)
    assert_close(tr.values, ht["treatment"], 0)
    np.random.seed(2024)
    pipeline.ht_1d_moments(ad, cov, tr, num_boot=int(ht["num_boot"]), num_cpus=1, **kw)
    res = ad.uns["memento"]["1d_ht"]
    for key in ["mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"]:
        assert_close(res[key], ht["%s_%s" % (variant, key)], 1e-8, atol=1e-12, what=(variant, key))


def test_ht_1d_one_sample(prepared):
    import pandas as pd
    ht = load("ht1d.npz")
    ad = prepared.copy()
    groups = ad.uns["memento"]["groups"]
    cov, _ = synth.design_from_groups(groups, ["stim", "cell"])
    ones = pd.DataFrame({"one": np.ones(len(groups))}, index=groups)
    np.random.seed(77)
    pipeline.ht_1d_moments(ad, cov, ones, num_boot=int(ht["num_boot"]), num_cpus=1,
                           resampling="bootstrap", approx=True)
    res = ad.uns["memento"]["1d_ht"]
    for key in ["mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"]:
        assert_close(res[key], ht["onesample_%s" % key], 1e-8, atol=1e-12, what=key)


def test_ht_2d(st, prepared):
    h2 = load("ht2d.npz")
    ht = load("ht1d.npz")
    ad = prepared.copy()
    names = ad.var.index.tolist()
    pairs = [(names[i], names[j]) for i, j in zip(st["pairs_idx1"], st["pairs_idx2"])]
    pipeline.compute_2d_moments(ad, pairs)
    cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
    np.random.seed(99)
    pipeline.ht_2d_moments(ad, cov, tr, num_boot=int(ht["num_boot"]), num_cpus=1,
                           resampling="bootstrap", approx=True)
    res = ad.uns["memento"]["2d_ht"]
    for key in ["corr_coef", "corr_se", "corr_asl"]:
        assert_close(res[key], h2[key], 1e-8, atol=1e-12, what=key)


def test_asl_branches():
    a = load("asl.npz")
    for name in ["count", "gev", "gev_neg", "const", "zero_extreme"]:
        x = a[name + "_x"]
        assert_close(testing.compute_asl(x.copy(), "bootstrap"), a[name + "_asl"], 1e-9, what=name)
        assert_close(testing.compute_asl(x.copy(), "bootstrap", approx=True), a[name + "_asl_approx"], 1e-9,
                     what=name + " approx")


def test_asl_gev_battery():
    """The oracle's compute_asl against the reference's on the first 12 rows of the GEV battery (regenerated rows are
    checked against the fixture's checksums; scipy's fits are deterministic, so the match is to round-off)."""
    from helpers import gev_battery_vector
    g = load("gev_battery.npz")
    for i in range(12):
        x = gev_battery_vector(i)
        np.testing.assert_allclose(np.sum(x * np.arange(1, x.size + 1)), g["checksum"][i], rtol=1e-12)
        assert_close(testing.compute_asl(x.copy(), "bootstrap"), g["asl"][i], 1e-9, what="row %d" % i)


def test_parallel_driver_matches_sequential_point_estimates(prepared):
    """num_cpus>1 uses a fork pool; coefficients (RNG-free, column 0) must not depend on it."""
    ad1, ad2 = prepared.copy(), prepared.copy()
    cov, tr = synth.design_from_groups(ad1.uns["memento"]["groups"], ["stim", "cell"])
    sub = list(range(12))
    pipeline.ht_1d_moments(ad1, cov, tr, num_boot=50, num_cpus=1, gene_subset=sub, resampling="bootstrap", approx=True)
    pipeline.ht_1d_moments(ad2, cov, tr, num_boot=50, num_cpus=2, gene_subset=sub, resampling="bootstrap", approx=True)
    assert_close(ad1.uns["memento"]["1d_ht"]["mean_coef"], ad2.uns["memento"]["1d_ht"]["mean_coef"], 1e-12)
