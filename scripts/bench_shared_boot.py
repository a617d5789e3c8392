"""ht_2d_moments(bootstrap='shared') on the BASELINE configs[2] block (C2 data: TFs x targets, 16 groups):
    python scripts/bench_shared_boot.py [n_a] [n_b] [num_boot] > profiles/rNN_bench_shared_boot.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
import torch
import memento_b200 as memento
from memento_b200 import synth, _lib

n_a = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n_b = int(sys.argv[2]) if len(sys.argv) > 2 else 0
num_boot = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, q=0.07, seed=7, device="cuda")
memento.setup_memento(ad, "q", profile=True)
memento.create_groups(ad, ["stim", "cell"])
memento.compute_1d_moments(ad, min_perc_group=0.7)
names = ad.var.index.to_numpy()
A = names[:n_a]
B = names if n_b <= 0 else names[:n_b]
t0 = time.perf_counter()
pairs = np.stack([np.repeat(A, B.size), np.tile(B, A.size)], axis=1)
t_pairs = time.perf_counter() - t0
t0 = time.perf_counter()
memento.compute_2d_moments(ad, pairs)
torch.cuda.synchronize()
t_2d = time.perf_counter() - t0
groups = ad.uns["memento"]["groups"]
cov, tr = synth.design_from_groups(groups, ["stim", "cell"])
st = ad.uns["memento"]["_b200"]
out = {}
for nb in (4, num_boot):          # the first call also builds the B panels; the difference is the per-replicate cost
    l0 = _lib.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    memento.ht_2d_moments(ad, cov, tr, num_boot=nb, resampling="bootstrap", approx=True, seed=1, bootstrap="shared")
    torch.cuda.synchronize()
    out[nb] = (time.perf_counter() - t0, _lib.launch_count() - l0)
stage = st.timer.collect()
# device time of the replicate loop (CUDA events around it, both calls): the wall clock of a call is dominated by the
# host-side bookkeeping of the 11 M-pair result schema (stacking per-group arrays, de-duplication, read-back)
per_rep = stage["shared_block_bootstrap"] * 1e-3 / (4 + num_boot)
ht = ad.uns["memento"]["2d_ht"]
n_pairs = pairs.shape[0]
flops = 3 * 2.0 * A.size * B.size * sum(-(-int(n) // 64) * 64 for n in np.diff(st.seg.group_start_host))
print(json.dumps({
    "block": "%d x %d genes, %d groups, 25000 cells" % (A.size, B.size, len(groups)), "pairs": int(n_pairs),
    "pair_array_s": t_pairs, "compute_2d_moments_s": t_2d,
    "ht_2d_shared_s": {str(k): v[0] for k, v in out.items()}, "launches": {str(k): v[1] for k, v in out.items()},
    "ms_per_replicate": per_rep * 1e3, "tensor_TFLOPs_per_s_incl_everything": flops / per_rep / 1e12,
    "pair_replicates_per_s": n_pairs / per_rep,
    "projected_s_num_boot_10000": per_rep * 1e4,
    "per_pair_path_pairs_per_s_at_B10000": 6800.0,
    "equivalent_pairs_per_s_at_B10000": n_pairs / (per_rep * 1e4),
    "finite_asl": int(np.isfinite(ht["corr_asl"]).sum()), "stage_ms": stage,
    "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
