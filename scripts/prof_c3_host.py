"""Host-side profile (cProfile, cumulative) of compute_2d_moments and ht_2d_moments(bootstrap='shared') on the BASELINE
configs[2] block: where the wall clock of the API call goes when the device work is milliseconds.
    python scripts/prof_c3_host.py > gpurun_out/prof_c3_host.txt"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
import torch
import memento_b200 as memento
from memento_b200 import synth

ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, q=0.07, seed=7, device="cuda")
memento.setup_memento(ad, "q")
memento.create_groups(ad, ["stim", "cell"])
memento.compute_1d_moments(ad, min_perc_group=0.7)
names = ad.var.index.to_numpy()
A, B = names[:1500], names
pairs = np.stack([np.repeat(A, B.size), np.tile(B, A.size)], axis=1)
cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])


def show(pr, title, n=22):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(n)
    print("=" * 30, title)
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])


for rep in range(2):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable(); memento.compute_2d_moments(ad, pairs); torch.cuda.synchronize(); pr.disable()
    print("compute_2d_moments wall", time.perf_counter() - t0)
    if rep == 1:
        show(pr, "compute_2d_moments")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
memento.ht_2d_moments(ad, cov, tr, num_boot=8, resampling="bootstrap", approx=True, seed=1, bootstrap="shared")
torch.cuda.synchronize(); pr.disable()
print("ht_2d_moments(shared, B=8) wall", time.perf_counter() - t0)
show(pr, "ht_2d_moments shared")
