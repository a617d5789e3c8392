"""World-size-2 gloo tests (CPU) of the gene-sharded multi-GPU host logic: shard planning, the two
all-reduces and the all-gathers that setup_memento / compute_1d_moments perform between kernels.
The per-rank arithmetic that the CUDA kernels do on the GPU box is done here by the oracle, so the
test checks the exchange design: sharded result == single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from helpers import golden_adata, load
from memento_b200 import dist as mdist
from oracle import moments as o_moments


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = mdist.DistContext()
        st = load("stages.npz")
        ad = golden_adata(st)
        X = ad.X
        work = np.diff(X.tocsc().indptr)
        bounds = mdist.shard_plan(work, world)
        Xl, lo, hi = mdist.shard_columns(X, bounds, rank)
        n = X.shape[0]
        # (1) UMI totals: local row sums -> all-reduce
        naive = ctx.all_reduce_sum(o_moments.row_totals(Xl).astype(np.float64))
        # (2) local moments with the global totals, all-gather for the global fit / quantile
        m, v = o_moments.hyper_1d_sparse(Xl, n, 0.07, naive)
        m[np.asarray(Xl.mean(axis=0)).reshape(-1) < 0.07] = 0
        gm, sizes = ctx.all_gather_concat(m)
        gv, _ = ctx.all_gather_concat(v)
        assert sizes == [int(bounds[r + 1] - bounds[r]) for r in range(world)]
        fit = o_moments.fit_mean_var(gm, gv)
        rv = o_moments.residual_variance(m, v, fit)
        grv, _ = ctx.all_gather_concat(rv)
        ulim = np.quantile(grv[np.isfinite(grv)], 0.1)
        rv[~np.isfinite(rv)] = np.inf
        mask = rv < ulim
        gmask, _ = ctx.all_gather_concat(mask)
        assert gmask.dtype == np.bool_
        # (3) masked totals -> all-reduce -> size factor
        tot = ctx.all_reduce_sum(np.asarray(Xl.multiply(mask).sum(axis=1)).reshape(-1).astype(np.float64))
        tot = tot + np.quantile(tot, 0.5)
        sf = tot / tot.mean()
        # ragged 2-D gather (G_local x R)
        two_d, _ = ctx.all_gather_concat(np.arange((hi - lo) * 3, dtype=np.float64).reshape(hi - lo, 3) + 1000 * rank)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), naive=naive, sf=sf, gmask=gmask, fit=fit,
                 two_d=two_d, bounds=bounds)
    finally:
        tdist.destroy_process_group()


def test_sharded_setup_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    st = load("stages.npz")
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    for key in ("naive", "sf", "gmask", "fit", "two_d"):
        assert np.array_equal(r0[key], r1[key]), key          # every rank ends with the same global view
    assert np.allclose(r0["naive"], st["naive_sf"], rtol=0, atol=0)
    assert np.allclose(r0["sf"], st["size_factor"], rtol=1e-12)
    names = np.array(["gene%d" % i for i in range(r0["gmask"].shape[0])])
    assert names[r0["gmask"]].tolist() == st["least_variable_genes"].tolist()
    b = r0["bounds"]
    assert r0["two_d"].shape == (b[-1], 3) and r0["two_d"][b[1], 0] == 1000.0


def test_shard_plan_balances_work():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 1000, size=5000).astype(float)
    for world in (1, 2, 4, 8):
        b = mdist.shard_plan(w, world)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all() and b.shape[0] == world + 1
        loads = np.array([w[b[r]:b[r + 1]].sum() for r in range(world)])
        assert loads.max() <= w.sum() / world + w.max() + 1
    assert mdist.shard_plan(np.zeros(0), 4).tolist() == [0, 0, 0, 0, 0]
    assert mdist.shard_plan(np.ones(3), 8)[-1] == 3


# ----------------------------------------------------------------------------- result gather of ht_1d_moments
def _gather_worker(rank, world, port, out_dir):
    """Every rank holds the results of its own (ragged) gene block; ``main._gather_1d_ht`` must leave ALL genes'
    results, in global gene order, on every rank (reference main.py:399-412 assembles all genes in one place)."""
    import types
    import pandas as pd
    from memento_b200 import main as mmain
    from memento_b200.anndata_lite import AnnDataLite
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_genes = [5, 3][rank]
        lo = [0, 5][rank]
        t_gene = np.array([[2, 1, 0, 3, 1], [1, 2, 2]][rank])          # tests per gene (treatment_for_gene pattern)
        n_tests = int(t_gene.sum())
        ht = {k: 100.0 * j + 10.0 * rank + np.arange(n_tests, dtype=np.float64)
              for j, k in enumerate(mmain._HT_KEYS)}
        ht["mean_asl"][0] = np.nan
        st = types.SimpleNamespace(dist=mdist.DistContext())
        var = pd.DataFrame(index=pd.Index(["g%d" % (lo + i) for i in range(n_genes)]))
        import scipy.sparse as sp
        ad = AnnDataLite(sp.csr_matrix((4, n_genes)), var=var, uns={"memento": {"_b200": st, "1d_ht": ht}})
        for _ in range(2):                      # the second call reuses the cached name list
            mmain._gather_1d_ht(ad, t_gene)
        res = ad.uns["memento"]["1d_ht_all"]
        np.savez(os.path.join(out_dir, "g%d.npz" % rank), gene=np.asarray(res["gene"], dtype="U"),
                 n_tests=res["n_tests"], **{k: res[k] for k in mmain._HT_KEYS})
    finally:
        tdist.destroy_process_group()


def test_result_gather_of_ragged_gene_blocks(tmp_path):
    from memento_b200 import main as mmain
    mp.spawn(_gather_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "g0.npz"), np.load(tmp_path / "g1.npz")
    for key in r0.files:
        assert np.array_equal(r0[key], r1[key], equal_nan=(r0[key].dtype.kind == "f")), key
    assert r0["gene"].tolist() == ["g%d" % i for i in range(8)]
    assert r0["n_tests"].tolist() == [2, 1, 0, 3, 1, 1, 2, 2]
    for j, k in enumerate(mmain._HT_KEYS):
        want = np.concatenate([100.0 * j + np.arange(7.0), 100.0 * j + 10.0 + np.arange(5.0)])
        if k == "mean_asl":
            want[[0, 7]] = np.nan            # every rank's first test
        assert np.array_equal(r0[k], want, equal_nan=True), k
