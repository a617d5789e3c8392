"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference package ``memento`` is imported from /root/reference through the stub shim in
``tests/golden/ref_shim`` (unused third-party imports only).  Inputs are produced by the repo's
own synthetic generator with fixed seeds; every reference call that consumes the global numpy RNG
is preceded by an explicit ``np.random.seed`` recorded in the fixture, so the oracle (and the
CUDA replay mode) can reproduce the exact call sequence.  Outputs: ``stages.npz``, ``ht1d.npz``,
``ht2d.npz``, ``asl.npz``, ``stages_f32.npz``, ``getters.npz``, ``gev_battery.npz``.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.append(os.path.join(ROOT, "scrna-parameter-estimation_b200"))    # after the reference: its `memento` must win

import memento.main as ref_main                      # noqa: E402  (the reference)
import memento.estimator as ref_est                  # noqa: E402
import memento.bootstrap as ref_boot                 # noqa: E402
import memento.hypothesis_test as ref_ht             # noqa: E402
from memento_b200 import synth                       # noqa: E402

assert ref_main.__file__.startswith("/root/reference"), ref_main.__file__
warnings.simplefilter("ignore")

N_CELLS, N_GENES, Q = 800, 150, 0.07
NUM_BOOT = 200


def dataset(dtype=np.float64):
    """float64 X: the reference then computes in float64 throughout (tight parity target).
    float32 X (h5ad convention): the reference's own moments are float32-accurate only."""
    ad = synth.make_counts(N_CELLS, N_GENES, n_conditions=2, n_types=2, q=Q, seed=7)
    ad.X = ad.X.astype(dtype)
    return ad


def stages_f32():
    """Point estimates of the reference on the SAME counts stored as float32."""
    out = {}
    ad = dataset(np.float32)
    ref_main.setup_memento(ad, q_column="q")
    mem = ad.uns["memento"]
    out["size_factor"] = ad.obs["memento_size_factor"].values
    out["least_variable_genes"] = np.asarray(mem["least_variable_genes"], dtype="U")
    out["all_mean"], out["all_var"] = mem["all_1d_moments"]
    ref_main.create_groups(ad, label_columns=["stim", "cell"])
    ref_main.compute_1d_moments(ad, min_perc_group=0.7)
    out["overall_gene_filter"] = mem["overall_gene_filter"]
    out["mv_regressor"] = mem["mv_regressor"]["all"]
    for gi, g in enumerate(mem["groups"]):
        out["m1d_mean_%d" % gi], out["m1d_var_%d" % gi], out["m1d_rv_%d" % gi] = mem["1d_moments"][g]
    np.savez_compressed(os.path.join(HERE, "stages_f32.npz"), **out)
    print("stages_f32.npz:", len(out), "arrays; dtypes", out["size_factor"].dtype, out["all_mean"].dtype)


def csr_parts(X, prefix, out):
    out[prefix + "_data"] = X.data
    out[prefix + "_indices"] = X.indices
    out[prefix + "_indptr"] = X.indptr
    out[prefix + "_shape"] = np.array(X.shape)


def stages():
    out = {}
    ad = dataset()
    csr_parts(ad.X, "X", out)
    out["stim"] = np.asarray(ad.obs["stim"].tolist(), dtype="U")
    out["cell"] = np.asarray(ad.obs["cell"].tolist(), dtype="U")
    out["q"] = ad.obs["q"].values

    # setup_memento internals, stage by stage (reference main.py:55-91)
    naive = ref_est._estimate_size_factor(ad.X, "hyper_relative", total=True, shrinkage=0.0)
    out["naive_sf"] = naive
    m, v = ref_est._hyper_1d_relative(ad.X, ad.shape[0], Q, naive)
    out["naive_mean"], out["naive_var"] = m.copy(), v.copy()

    ref_main.setup_memento(ad, q_column="q")
    mem = ad.uns["memento"]
    out["size_factor"] = ad.obs["memento_size_factor"].values
    out["least_variable_genes"] = np.asarray(mem["least_variable_genes"], dtype="U")
    out["all_mean"], out["all_var"] = mem["all_1d_moments"]
    out["all_q"] = np.array(mem["all_q"])

    ref_main.create_groups(ad, label_columns=["stim", "cell"])
    out["groups"] = np.asarray(mem["groups"], dtype="U")
    out["group_q"] = np.array([mem["group_q"][g] for g in mem["groups"]])
    out["group_ncells"] = np.array([mem["group_cells"][g].shape[0] for g in mem["groups"]])

    ref_main.compute_1d_moments(ad, min_perc_group=0.7)
    out["approx_sf"] = mem["all_approx_size_factor"]
    out["overall_gene_filter"] = mem["overall_gene_filter"]
    out["gene_list"] = np.asarray(mem["gene_list"], dtype="U")
    out["mv_regressor"] = mem["mv_regressor"]["all"]
    for gi, g in enumerate(mem["groups"]):
        out["m1d_mean_%d" % gi], out["m1d_var_%d" % gi], out["m1d_rv_%d" % gi] = mem["1d_moments"][g]
        out["gene_filter_%d" % gi] = mem["gene_filter"][g]
        out["gene_rv_filter_%d" % gi] = mem["gene_rv_filter"][g]

    # unique tables + bootstrap replicates for a few (gene, group) slices
    G = ad.shape[1]
    picks = [(0, 0), (1, 1), (G // 2, 2), (G - 1, 3), (3, 0)]
    out["table_picks"] = np.array(picks)
    for k, (gene, gi) in enumerate(picks):
        g = mem["groups"][gi]
        col = mem["group_cells"][g][:, gene]
        sf = mem["approx_size_factor"][g]
        np.random.seed(100 + k)
        inv_sf, inv_sf_sq, expr, counts = ref_boot._unique_expr(col, sf)
        out["tab%d_inv_sf" % k], out["tab%d_expr" % k], out["tab%d_counts" % k] = inv_sf, expr, counts
        np.random.seed(100 + k)
        mean, var = ref_boot._bootstrap_1d(col, sf, mem["group_q"][g], ref_est._hyper_1d_relative, num_boot=64)
        out["tab%d_boot_mean" % k], out["tab%d_boot_var" % k] = mean, var
        gen = np.random.Generator(np.random.PCG64(5))
        out["tab%d_W" % k] = gen.multinomial(col.shape[0], counts / counts.sum(), size=64).T

    # 2D point estimates
    names = ad.var.index.tolist()
    pairs = [(names[0], names[j]) for j in range(1, 12)] + [(names[5], names[5]), (names[7], names[2]), (names[2], names[7])]
    ref_main.compute_2d_moments(ad, pairs)
    out["pairs_idx1"] = mem["2d_moments"]["gene_idx_1"]
    out["pairs_idx2"] = mem["2d_moments"]["gene_idx_2"]
    for gi, g in enumerate(mem["groups"]):
        d = mem["2d_moments"][g]
        out["m2d_cov_%d" % gi], out["m2d_corr_%d" % gi] = d["cov"], d["corr"]
        out["m2d_var1_%d" % gi], out["m2d_var2_%d" % gi] = d["var_1"], d["var_2"]
    gi = 1
    g = mem["groups"][gi]
    out["corr_matrix_g1"] = ref_main.get_corr_matrix(ad, g)

    # one 2D unique table + bootstrap
    cols = mem["group_cells"][g][:, [0, 3]]
    np.random.seed(321)
    inv_sf, inv_sf_sq, expr, counts = ref_boot._unique_expr(cols, mem["approx_size_factor"][g])
    out["tab2d_inv_sf"], out["tab2d_expr"], out["tab2d_counts"] = inv_sf, expr, counts
    np.random.seed(321)
    cov, v1, v2 = ref_boot._bootstrap_2d(cols, mem["approx_size_factor"][g], mem["group_q"][g],
                                         ref_est._hyper_1d_relative, ref_est._hyper_cov_relative, num_boot=64)
    out["tab2d_cov"], out["tab2d_var1"], out["tab2d_var2"] = cov, v1, v2
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **out)
    print("stages.npz:", len(out), "arrays;", ad.shape[1], "genes pass")
    return ad


def designs(ad):
    groups = ad.uns["memento"]["groups"]
    return synth.design_from_groups(groups, ["stim", "cell"])


def ht1d(ad):
    out = {}
    cov, tr = designs(ad)
    out["covariate"], out["treatment"] = cov.values, tr.values
    variants = {
        "default": dict(resampling="bootstrap"),
        "approx": dict(resampling="bootstrap", approx=True),
        "resample_rep": dict(resampling="bootstrap", approx=True, resample_rep=True),
    }
    for name, kw in variants.items():
        a = ad.copy()
        np.random.seed(2024)
        ref_main.ht_1d_moments(a, covariate=cov, treatment=tr, num_boot=NUM_BOOT, num_cpus=1, verbose=0, **kw)
        for key in ["mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"]:
            out["%s_%s" % (name, key)] = a.uns["memento"]["1d_ht"][key]
    # one-sample test (treatment all ones) on the control groups only is exercised through _regress_1d
    a = ad.copy()
    ones = pd.DataFrame({"one": np.ones(len(a.uns["memento"]["groups"]))}, index=a.uns["memento"]["groups"])
    np.random.seed(77)
    ref_main.ht_1d_moments(a, covariate=cov, treatment=ones, num_boot=NUM_BOOT, num_cpus=1, verbose=0,
                           resampling="bootstrap", approx=True)
    for key in ["mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl"]:
        out["onesample_%s" % key] = a.uns["memento"]["1d_ht"][key]
    # per-gene boot arrays for two genes (debug path of _ht_1d restated: we call the pieces)
    out["num_boot"] = np.array(NUM_BOOT)
    np.savez_compressed(os.path.join(HERE, "ht1d.npz"), **out)
    print("ht1d.npz:", len(out), "arrays")


def ht2d(ad):
    out = {}
    cov, tr = designs(ad)
    a = ad.copy()
    np.random.seed(99)
    ref_main.ht_2d_moments(a, covariate=cov, treatment=tr, num_boot=NUM_BOOT, num_cpus=1, verbose=0,
                           resampling="bootstrap", approx=True)
    for key in ["corr_coef", "corr_se", "corr_asl"]:
        out[key] = a.uns["memento"]["2d_ht"][key]
    np.savez_compressed(os.path.join(HERE, "ht2d.npz"), **out)
    print("ht2d.npz:", len(out), "arrays")


def asl_cases():
    """_compute_asl on fixed vectors: the three branches (approx / count / GEV tails)."""
    out = {}
    rng = np.random.default_rng(3)
    cases = {
        "count": np.concatenate([[0.05], 0.05 + rng.normal(0, 0.1, 2000)]),
        "gev": np.concatenate([[0.45], 0.45 + rng.normal(0, 0.1, 3000)]),
        "gev_neg": np.concatenate([[-0.5], -0.5 + rng.standard_t(6, 4000) * 0.1]),
        "const": np.full(50, 0.3),
        "zero_extreme": np.concatenate([[2.0], 2.0 + rng.normal(0, 0.1, 1500)]),
    }
    for name, x in cases.items():
        out[name + "_x"] = x
        out[name + "_asl"] = np.array(ref_ht._compute_asl(x.copy(), resampling="bootstrap"))
        out[name + "_asl_approx"] = np.array(ref_ht._compute_asl(x.copy(), resampling="bootstrap", approx=True))
    np.savez_compressed(os.path.join(HERE, "asl.npz"), **out)
    print("asl.npz:", {k: float(v) for k, v in out.items() if k.endswith("_asl")})


def gev_battery(n_cases=60):
    """_compute_asl (hypothesis_test.py:57-141) on the rows of tests/helpers.py:gev_battery_vector: the fixture keeps
    the reference's ASL, the extreme count and a checksum of every regenerated row (rows are not stored)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import gev_battery_vector
    asl, cnt, chk, size = [], [], [], []
    for i in range(n_cases):
        x = gev_battery_vector(i)
        null = x[1:] - x[0]
        a = abs(x[0])
        cnt.append(int((null > a).sum() + (null < -a).sum()))
        asl.append(float(ref_ht._compute_asl(x.copy(), resampling="bootstrap")))
        chk.append(float(np.sum(x * np.arange(1, x.size + 1))))
        size.append(x.size)
    np.savez_compressed(os.path.join(HERE, "gev_battery.npz"), asl=np.array(asl), extreme=np.array(cnt),
                        checksum=np.array(chk), size=np.array(size))
    print("gev_battery.npz:", n_cases, "rows,", int((np.array(cnt) <= 10).sum()), "in the GEV branch")


def getters():
    """The reference's result getters (main.py:523-670) and BH correction (util.py:22-29) on a small fabricated
    ``uns['memento']`` (the getters only read the dictionary)."""
    import memento.util as ref_util
    from memento_b200.anndata_lite import AnnDataLite
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    G, n_pairs = 7, 9
    groups = ["sg^ctrl^A", "sg^ctrl^B", "sg^stim^A", "sg^stim^B"]
    n_cells = [30, 50, 20, 40]
    obs = pd.DataFrame({"stim": np.repeat(["ctrl", "ctrl", "stim", "stim"], n_cells), "cell": np.repeat(["A", "B", "A", "B"], n_cells)})
    var = pd.DataFrame(index=pd.Index(["g%d" % i for i in range(G)]))
    ad = AnnDataLite(sp.csr_matrix((sum(n_cells), G)), obs, var)
    mom = rng.gamma(2.0, 1.0, size=(len(groups), 3, G))
    mom[0, 0, 2] = 0.0          # a zero mean (log -> -inf) and a NaN residual variance
    mom[1, 2, 4] = np.nan
    corr = rng.uniform(-1, 1, size=(len(groups), n_pairs))
    corr[2, 3] = np.nan
    pairs = [("g%d" % rng.integers(G), "g%d" % rng.integers(G)) for _ in range(n_pairs)]
    ht = {k: rng.normal(size=G) for k in ("mean_coef", "mean_se", "var_coef", "var_se")}
    ht.update({k: rng.uniform(size=G) for k in ("mean_asl", "var_asl")})
    ht["treatment"] = pd.DataFrame({"stim": [0, 0, 1, 1]}, index=groups)
    ht2 = {"corr_coef": rng.normal(size=n_pairs), "corr_se": rng.uniform(size=n_pairs), "corr_asl": rng.uniform(size=n_pairs)}
    ad.uns["memento"] = {
        "groups": groups, "label_columns": ["stim", "cell"], "label_delimiter": "^",
        "group_cells": {g: sp.csr_matrix((n, G)) for g, n in zip(groups, n_cells)},
        "1d_moments": {g: [mom[i, 0].copy(), mom[i, 1].copy(), mom[i, 2].copy()] for i, g in enumerate(groups)},
        "2d_moments": {"gene_pairs": pairs, **{g: {"corr": corr[i].copy()} for i, g in enumerate(groups)}},
        "1d_ht": ht, "2d_ht": ht2}
    out = {"groups": np.array(groups), "n_cells": np.array(n_cells), "mom": mom, "corr": corr,
           "pairs": np.array(pairs), "stim": np.array(obs["stim"].tolist(), dtype=str), "cell": np.array(obs["cell"].tolist(), dtype=str)}
    for k, v in ht.items():
        if k != "treatment":
            out["ht_" + k] = v
    for k, v in ht2.items():
        out["ht2_" + k] = v
    with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m, v, counts = ref_main.get_1d_moments(ad)
        out["m1_mean"], out["m1_var"] = m.drop(columns="gene").values, v.drop(columns="gene").values
        out["m1_cols"] = np.array(m.columns[1:].tolist())
        for gb in ("stim", "cell", "ALL"):
            m, v = ref_main.get_1d_moments(ad, groupby=gb)
            out["m1_%s_mean" % gb], out["m1_%s_var" % gb] = m.drop(columns="gene").values, v.drop(columns="gene").values
            out["m1_%s_cols" % gb] = np.array(m.columns[1:].tolist())
        c, _ = ref_main.get_2d_moments(ad)
        out["m2_corr"] = c.drop(columns=["gene_1", "gene_2"]).values.astype(float)
        for gb in ("cell", "ALL"):
            # the reference zeroes the NaNs of uns['memento']['2d_moments'][g]['corr'] in place: give it a copy
            ad.uns["memento"]["2d_moments"] = {"gene_pairs": pairs, **{g: {"corr": corr[i].copy()} for i, g in enumerate(groups)}}
            c = ref_main.get_2d_moments(ad, groupby=gb)
            out["m2_%s_corr" % gb] = c.drop(columns=["gene_1", "gene_2"]).values.astype(float)
            out["m2_%s_cols" % gb] = np.array(c.columns[2:].tolist())
        r1 = ref_main.get_1d_ht_result(ad)
        out["r1_gene"], out["r1_tx"] = np.array(r1["gene"].tolist(), dtype=str), np.array(r1["tx"].tolist(), dtype=str)
        out["r1_vals"] = r1[["de_coef", "de_se", "de_pval", "dv_coef", "dv_se", "dv_pval"]].values
        r2 = ref_main.get_2d_ht_result(ad)
        out["r2_vals"] = r2[["corr_coef", "corr_se", "corr_pval"]].values
        p = rng.uniform(size=40) ** 2
        p[[3, 17]] = np.nan
        out["fdr_p"], out["fdr_q"] = p, ref_util._fdrcorrect(p.copy())
    np.savez_compressed(os.path.join(HERE, "getters.npz"), **out)
    print("getters.npz:", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("getters", "gev_battery"):
        {"getters": getters, "gev_battery": gev_battery}[sys.argv[1]]()
        sys.exit(0)
    ad = stages()
    stages_f32()
    ht1d(ad)
    ht2d(ad)
    asl_cases()
    getters()
    gev_battery()
