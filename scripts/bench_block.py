"""C3 of BASELINE.json: 2D moments of a dense 1.5k x 10k gene-pair block on the C2-shaped matrix, timed on the
device (panels + tcgen05 GEMM, all 16 groups), with the tensor-core roofline next to it.  Prints one JSON line.
    python scripts/bench_block.py [--tfs 1500] [--targets 10000] [--reps 5]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
import torch
import memento_b200 as memento
from memento_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=25000); ap.add_argument("--genes", type=int, default=10000)
ap.add_argument("--tfs", type=int, default=1500); ap.add_argument("--targets", type=int, default=10000)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
ad = synth.make_counts_fast(a.cells, a.genes, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q"); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
st = ad.uns["memento"]["_b200"]; seg = st.seg
G = seg.G
idx_a = np.arange(min(a.tfs, G)); idx_b = np.arange(min(a.targets, G))
sums = seg.moments(st.inv_sf_sorted)
for _ in range(2):
    out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums); del out
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    out = seg.block_cross(idx_a, idx_b, st.inv_sf_sorted, sums); del out
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
n_g = np.diff(seg.group_start_host)
flops = seg.block_flops(len(idx_a), len(idx_b), n_g)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = peaks.get("bf16_tflops", 1590.0)
print(json.dumps({"workload": "configs[2]: 2D moments of a %d x %d gene block, %d cells, %d groups" % (len(idx_a), len(idx_b), seg.n_cells, seg.R),
                  "ms": ms, "pairs_per_s": len(idx_a) * len(idx_b) / (ms * 1e-3),
                  "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                               "frac": flops / (ms * 1e-3) / 1e12 / peak,
                               "note": "3 fp16 products per (pair, cell) incl. padding to 64 cells per group; time includes the "
                                       "panel build and the float64 block write (%.2f GB)" % (len(idx_a) * len(idx_b) * seg.R * 8 / 1e9)}}))
