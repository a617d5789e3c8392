"""Two-rank NCCL test of the gene-sharded pipeline (needs >= 2 GPUs; skipped otherwise): each rank
holds half of the gene columns, and the sharded results must equal the single-GPU run."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as tdist
    import memento_b200 as memento
    from helpers import golden_adata
    from memento_b200 import dist as mdist, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = mdist.DistContext(device=dev)
        ad = golden_adata()
        bounds = mdist.shard_plan(np.diff(ad.X.tocsc().indptr), world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        keep = np.zeros(ad.shape[1], dtype=bool)
        keep[lo:hi] = True
        ad._inplace_subset_var(keep)
        ad.X = ad.X.tocsr()
        memento.setup_memento(ad, "q", dist=ctx, gene_offset=lo)
        memento.create_groups(ad, ["stim", "cell"])
        memento.compute_1d_moments(ad, min_perc_group=0.7)
        cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
        memento.ht_1d_moments(ad, cov, tr, num_boot=300, resampling="bootstrap", approx=True, seed=3)
        mem = ad.uns["memento"]
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), sf=ad.obs["memento_size_factor"].values,
                 genes=np.asarray(ad.var.index.tolist(), dtype="U"), fit=mem["mv_regressor"]["all"],
                 mean0=mem["1d_moments"][mem["groups"][0]][0], rv0=mem["1d_moments"][mem["groups"][0]][2],
                 mean_coef=mem["1d_ht"]["mean_coef"], mean_asl=mem["1d_ht"]["mean_asl"],
                 var_asl=mem["1d_ht"]["var_asl"])
    finally:
        tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_pipeline_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    import memento_b200 as memento
    from helpers import golden_adata
    from memento_b200 import synth
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ad = golden_adata()
    memento.setup_memento(ad, "q")
    memento.create_groups(ad, ["stim", "cell"])
    memento.compute_1d_moments(ad, min_perc_group=0.7)
    cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
    memento.ht_1d_moments(ad, cov, tr, num_boot=300, resampling="bootstrap", approx=True, seed=3)
    mem = ad.uns["memento"]
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(2)]
    assert np.allclose(parts[0]["sf"], ad.obs["memento_size_factor"].values, rtol=1e-13)
    assert np.concatenate([p["genes"] for p in parts]).tolist() == ad.var.index.tolist()
    assert np.allclose(parts[0]["fit"], mem["mv_regressor"]["all"], rtol=1e-9)
    cat = lambda k: np.concatenate([p[k] for p in parts])  # noqa: E731
    g0 = mem["groups"][0]
    assert np.allclose(cat("mean0"), mem["1d_moments"][g0][0], rtol=1e-12)
    assert np.allclose(cat("rv0"), mem["1d_moments"][g0][2], rtol=1e-8, equal_nan=True)
    # global RNG stream ids: the sharded run draws the same replicates as the single-GPU run
    assert np.allclose(cat("mean_coef"), mem["1d_ht"]["mean_coef"], rtol=1e-9, equal_nan=True)
    assert np.allclose(cat("mean_asl"), mem["1d_ht"]["mean_asl"], rtol=1e-6, equal_nan=True)
    assert np.allclose(cat("var_asl"), mem["1d_ht"]["var_asl"], rtol=1e-6, equal_nan=True)
