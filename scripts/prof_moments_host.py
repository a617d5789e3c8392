import cProfile, pstats, sys, os, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import torch
import memento_b200 as memento
from memento_b200 import synth
ad = synth.make_counts_fast(1_000_000, 2500, n_conditions=2, n_types=20, q=0.07, seed=7, device="cuda")
for name, f in (("setup", lambda: memento.setup_memento(ad, "q")), ("groups", lambda: memento.create_groups(ad, ["stim", "cell"])),
                ("moments", lambda: memento.compute_1d_moments(ad, min_perc_group=0.7))):
    pr = cProfile.Profile(); pr.enable(); f(); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
    print("=====", name); print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:22]))
