python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'scrna-parameter-estimation_b200')
import torch, numpy as np
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(25000, 2000, n_conditions=2, n_types=8, seed=7, device='cuda')
memento.setup_memento(ad,'q', profile=True); memento.create_groups(ad,['stim','cell']); memento.compute_1d_moments(ad)
st = ad.uns['memento']['_b200']; st.count_modes=True
cov,tr = synth.design_from_groups(ad.uns['memento']['groups'],['stim','cell'])
for sampler, ma in (('poisson',0.1),('poisson',0.15),('poisson',0.2),('poisson',0.3),('chain',0)):
    st.min_accept = ma
    memento.ht_1d_moments(ad,cov,tr,num_boot=10000,resampling='bootstrap',approx=True,sampler=sampler)
    st.timer.collect(); st.timer.ms.clear()
    torch.cuda.synchronize(); t=time.time()
    memento.ht_1d_moments(ad,cov,tr,num_boot=10000,resampling='bootstrap',approx=True,sampler=sampler)
    torch.cuda.synchronize(); print(sampler, ma, 'genes', ad.shape[1], 'wall', time.time()-t, st.timer.collect(), {k:v for k,v in st.last_stats.items()})
    st.timer.ms.clear()
PY
