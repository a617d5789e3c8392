"""Where the roles of block_gemm_kernel wait (MM_BLOCK_DEBUG=9 counters), on the BASELINE configs[2] block:
    MM_BLOCK_DEBUG=9 python scripts/diag_block_waits.py"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
import torch
import memento_b200 as memento
from memento_b200 import synth, _lib

ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q"); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
st = ad.uns["memento"]["_b200"]; seg = st.seg
idx_a, idx_b = np.arange(1500), np.arange(seg.G)
sums = seg.moments(st.inv_sf_sorted)
buf = (ctypes.c_uint64 * 8)()
lib = _lib.load()
out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums); del out
lib.mm_block_debug_counters(0, buf)                # reset after the warm-up
out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums); del out
lib.mm_block_debug_counters(0, buf)
v = [int(x) for x in buf]
ctas = max(1, v[7])
names = ["producer waits for an empty slot", "MMA thread waits for a full slot", "MMA thread waits for a drained accumulator",
         "epilogue warp waits for a finished accumulator", "epilogue warp: TMEM loads + float64 adds", "epilogue warp: stores",
         "kernel"]
print(json.dumps({"ctas": ctas, "cycles_per_cta": {n: round(c / ctas) for n, c in zip(names, v[:7])}}, indent=1))
