"""Runs the public API on the larger BASELINE.json shapes (one GPU's gene shard of them) and prints one JSON line
per shape with the wall time of every API call: a scale / limits check (grid limits, int32 offsets, shared-memory
staging bounds, tile planning), not a parity test.
    python scripts/scale_check.py northstar   # 1M cells x 2500 genes (1/8 of 20k), 2 x 20 groups, B = 10k
    python scripts/scale_check.py c4          # 250k cells x 8000 genes, 2 x 1000 groups, B = 10k, approx, 256 genes tested"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
import torch
import memento_b200 as memento
from memento_b200 import synth

which = sys.argv[1] if len(sys.argv) > 1 else "northstar"
cfg = {"northstar": dict(cells=1_000_000, genes=2500, types=20, approx=False, test_genes=None),
       "c4": dict(cells=250_000, genes=8000, types=1000, approx=True, test_genes=256)}[which]
t = {}
def timed(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); t[name] = time.perf_counter() - t0; return r
ad = timed("synth", lambda: synth.make_counts_fast(cfg["cells"], cfg["genes"], n_conditions=2, n_types=cfg["types"], q=0.07, seed=7, device="cuda"))
timed("setup_memento", lambda: memento.setup_memento(ad, "q", profile=True))
timed("create_groups", lambda: memento.create_groups(ad, ["stim", "cell"]))
gl = None
timed("compute_1d_moments", lambda: memento.compute_1d_moments(ad, min_perc_group=0.7))
if cfg["test_genes"]:
    keep = ad.var.index[:cfg["test_genes"]].tolist()
    timed("compute_1d_moments(gene_list)", lambda: memento.compute_1d_moments(ad, min_perc_group=0.7, gene_list=keep))
groups = ad.uns["memento"]["groups"]
cov, tr = synth.design_from_groups(groups, ["stim", "cell"])
if which == "c4":
    # Perturb-seq style: thousands of guide groups, no per-group covariates (a 1000-column dummy design would make the
    # per-mask weighted QR the whole cost, here and in the reference)
    cov = cov.iloc[:, :1] * 0 + 1.0
kw = dict(num_boot=10000, resampling="bootstrap", approx=cfg["approx"])
timed("ht_1d_moments(first)", lambda: memento.ht_1d_moments(ad, cov, tr, seed=1, **kw))
timed("ht_1d_moments", lambda: memento.ht_1d_moments(ad, cov, tr, seed=2, **kw))
st = ad.uns["memento"]["_b200"]
res = ad.uns["memento"]["1d_ht"]
out = {"shape": which, "cells": cfg["cells"], "genes": cfg["genes"], "groups": len(groups), "nnz": int(st.seg.nnz),
       "genes_tested": int(ad.shape[1]), "finite_asl": int(np.isfinite(res["mean_asl"]).sum()),
       "seconds": {k: round(v, 4) for k, v in t.items()},
       "genes_per_s": ad.shape[1] / t["ht_1d_moments"], "stage_ms": {k: round(v, 2) for k, v in st.timer.collect().items()}, "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
