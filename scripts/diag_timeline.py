"""Device timeline of one ht_1d_moments call on the bench workload: start / end of every timed stage relative to the
start of the call (CUDA events), to see where the device idles between stages."""
import os, sys, time
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
    memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=1)
st.timer.collect(); st.timer.ms.clear()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=2)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("wall %.1f ms, device %.1f ms" % ((t1 - t0) * 1e3, e0.elapsed_time(e1)))
rows = sorted((e0.elapsed_time(a), e0.elapsed_time(b), name) for name, a, b in st.timer.pending)
for a, b, name in rows:
    print("%8.2f -> %8.2f  (%7.2f)  %s" % (a, b, b - a, name))
