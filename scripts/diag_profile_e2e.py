"""cProfile of one host-staged ht_1d_moments call (where does the host spend the time before the first kernel?)."""
import os, sys, time, gc, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import torch
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q", profile=True); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
st = ad.uns["memento"]["_b200"]
for _ in range(3):
    st.offload(); memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=1)
st.offload(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=2)
pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("cumulative").print_stats(45)
