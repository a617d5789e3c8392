"""Step time of ht_1d on the bench workload as a function of the acceptance threshold below which a segment uses the
conditional-binomial chain instead of the Poissonised sampler."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q", profile=True); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
st = ad.uns["memento"]["_b200"]
st.count_modes = True
for ma in (0.2, 0.1, 0.05, 0.02, 0.3, 0.2):
    st.min_accept = ma
    memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=1)
    torch.cuda.synchronize(); st.timer.collect(); st.timer.ms.clear()
    t0 = time.perf_counter()
    memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ms = st.timer.collect()
    print("min_accept", ma, "step %.1f ms" % (dt * 1e3), "bootstrap %.1f ms" % ms.get("bootstrap_1d", 0),
          "poisson segs", st.last_stats.get("poisson_segments"), "chain segs", st.last_stats.get("chain_segments"), "direct segs", st.last_stats.get("direct_segments"))
    st.timer.ms.clear()
