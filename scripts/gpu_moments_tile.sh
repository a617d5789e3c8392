# correctness of the tile kernel + a sweep of its configurations on the C2 matrix
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moments.py -m gpu -x -q 2>&1 | tail -15
cat > /tmp/mom.py <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200')
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device='cuda')
memento.setup_memento(ad,'q'); memento.create_groups(ad,['stim','cell']); memento.compute_1d_moments(ad)
st = ad.uns['memento']['_b200']; sg = st.seg; sf = st.inv_sf_sorted
for _ in range(3): sg.moments(sf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): sg.moments(sf)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
print("kernel", os.environ.get("MM_MOMENTS_KERNEL","auto"), "threads", os.environ.get("MM_MOMENTS_THREADS","-"), "chunk", os.environ.get("MM_MOMENTS_CHUNK","-"), "notile", os.environ.get("MM_MOMENTS_NOTILE"), "nnz", sg.nnz, "n_seg", sg.n_seg, "ms", round(ms,4), "GB/s", round(sg.moments_bytes()/ms/1e6,1))
PY
for c in 4 8; do for t in 640 896; do MM_MOMENTS_KERNEL=stream MM_MOMENTS_CHUNK=$c MM_MOMENTS_THREADS=$t timeout 300 python /tmp/mom.py 2>&1 | grep GB/s; done; done
