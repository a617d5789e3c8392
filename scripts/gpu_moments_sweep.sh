cat > /tmp/mom.py <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200')
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device='cuda')
memento.setup_memento(ad,'q'); memento.create_groups(ad,['stim','cell']); memento.compute_1d_moments(ad)
st = ad.uns['memento']['_b200']; seg = st.seg
for name, sg, sf in (("grouped", seg, st.inv_sf_sorted),):
    for _ in range(3): sg.moments(sf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): sg.moments(sf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(os.environ.get("MM_MOMENTS_W","auto"), name, "nnz", sg.nnz, "n_seg", sg.n_seg, "ms", ms, "GB/s", sg.moments_bytes()/ms/1e6)
PY
python /tmp/mom.py 2>&1 | grep GB/s; MM_MOMENTS_NOFLAT=1 MM_MOMENTS_NOSMEM=1 python /tmp/mom.py 2>&1 | grep GB/s

