"""mm_seg_moments in the short-segment regime (C5 of BASELINE.json: thousands of groups): one GPU's gene shard,
1.2M cells x 2500 genes x 4000 groups.  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np, torch
import memento_b200 as memento
from memento_b200 import synth
t = {}
def timed(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); t[name] = round(time.perf_counter() - t0, 3); return r
ad = timed("synth", lambda: synth.make_counts_fast(1_200_000, 2500, n_conditions=2, n_types=2000, q=0.1, seed=7, device="cuda"))
timed("setup_memento", lambda: memento.setup_memento(ad, "q"))
timed("create_groups", lambda: memento.create_groups(ad, ["stim", "cell"]))
st = ad.uns["memento"]["_b200"]; seg = st.seg
if getattr(st, "inv_sf_sorted", None) is None:
    from memento_b200 import main as M
    M._bin_size_factor(ad)
sf = st.inv_sf_sorted
res = {}
for kern in ("auto", "tile", "stream"):
    if kern == "auto": os.environ.pop("MM_MOMENTS_KERNEL", None)
    else: os.environ["MM_MOMENTS_KERNEL"] = kern
    for _ in range(2): seg.moments(sf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): seg.moments(sf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[kern] = {"ms": round(ms, 4), "GB/s": round(seg.moments_bytes() / ms / 1e6, 1)}
os.environ.pop("MM_MOMENTS_KERNEL", None)
timed("compute_1d_moments", lambda: memento.compute_1d_moments(ad, min_perc_group=0.05))
print(json.dumps({"shape": "c5 shard", "cells": 1_200_000, "genes": 2500, "groups": seg.R, "nnz": seg.nnz, "n_seg": seg.n_seg,
                  "mean_segment": seg.nnz / seg.n_seg, "algorithmic_bytes": seg.moments_bytes(), "seg_moments": res, "seconds": t}))
