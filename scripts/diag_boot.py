import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200')
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth, engine
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device='cuda')
memento.setup_memento(ad,'q'); memento.create_groups(ad,['stim','cell']); memento.compute_1d_moments(ad)
st = ad.uns['memento']['_b200']; seg = st.seg
mm = ad.uns['memento']
if st.design is None:
    from memento_b200 import main as M; M._refresh_design(ad)
G = 1258
tab = engine.unique_tables(seg, st.design, st.cell_bin, 0, G, 0)
U = tab["seg_U"].clamp(min=0)
ent = tab["entries"].view(-1, 32)
n_field = ent[:, 24:28].contiguous().view(torch.int32).reshape(-1)
lo = (seg.seg_ptr[:G*seg.R] - seg.seg_ptr[0]).long()
# gather n of the valid entries
idx = torch.cat([torch.arange(int(l), int(l)+int(u), device=U.device) for l,u in zip(lo.tolist()[:2000], U.tolist()[:2000])])
nn = n_field[idx].cpu().numpy()
print("segments", 2000, "mean U", U[:2000].float().mean().item(), "frac n<=16", (nn<=16).mean(), "frac n==1", (nn==1).mean(), "frac n<=4", (nn<=4).mean())
for sampler in ("poisson", "chain", "poisson"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.bootstrap_tile(seg, st.design, tab, G, 0, 10000, 5, sampler=sampler)
    e1.record(); torch.cuda.synchronize()
    print(sampler, e0.elapsed_time(e1), "ms")
