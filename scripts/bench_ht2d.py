"""C3, second half: ht_2d_moments on a 20 000-pair subsample of the 1.5k x 10k block (C2-shaped matrix).
Prints one JSON line with wall times."""
import json, os, sys, time, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np, torch
import memento_b200 as memento
from memento_b200 import synth
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q", profile=True); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
names = ad.var.index.tolist()
rng = np.random.default_rng(0)
tfs = names[:1500]
pairs = [(tfs[i], names[j]) for i, j in zip(rng.integers(0, len(tfs), n_pairs), rng.integers(0, len(names), n_pairs))]
t = {}
def timed(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); t[name] = round(time.perf_counter() - t0, 3); return r
timed("compute_2d_moments", lambda: memento.compute_2d_moments(ad, pairs))
cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
for approx in (True, False):
    timed("ht_2d_moments(approx=%s)" % approx, lambda: memento.ht_2d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", approx=approx, seed=1))
st = ad.uns["memento"]["_b200"]
res = ad.uns["memento"]["2d_ht"]
print(json.dumps({"workload": "configs[2]: ht_2d_moments on %d random pairs of a 1500 x %d block, 25000 cells, 16 groups, num_boot=10000" % (n_pairs, len(names)),
                  "seconds": t, "pairs_per_s_default": n_pairs / t["ht_2d_moments(approx=False)"],
                  "finite_asl": int(np.isfinite(res["corr_asl"]).sum()), "stage_ms": {k: round(v, 1) for k, v in st.timer.collect().items()}}))
