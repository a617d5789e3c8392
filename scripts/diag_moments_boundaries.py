import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200')
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device='cuda')
memento.setup_memento(ad,'q'); memento.create_groups(ad,['stim','cell'])
st = ad.uns['memento']['_b200']
w = torch.rand(25000, dtype=torch.float64, device='cuda') + 0.5
for name, sg in (("all-cells (1 segment per gene)", st.seg_all), ("16 groups", st.seg)):
    for _ in range(3): sg.moments(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): sg.moments(w)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(name, "nnz", sg.nnz, "n_seg", sg.n_seg, "mean", sg.nnz // sg.n_seg, "ms %.4f" % ms, "GB/s %.1f" % (sg.moments_bytes()/ms/1e6))
