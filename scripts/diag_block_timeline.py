import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np, torch
import memento_b200 as memento
from memento_b200 import synth
from torch.profiler import profile, ProfilerActivity
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q"); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
st = ad.uns["memento"]["_b200"]; seg = st.seg
idx_a, idx_b = np.arange(1500), np.arange(seg.G)
sums = seg.moments(st.inv_sf_sorted)
for _ in range(2):
    out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums); del out
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print("n kernels", len(ev), "span us", ev[-1].time_range.end - t0, "sum us", sum(e.time_range.end - e.time_range.start for e in ev))
for e in ev[:14] + ev[-8:]:
    print(round(e.time_range.start - t0, 1), round(e.time_range.end - e.time_range.start, 1), e.name[:60])
