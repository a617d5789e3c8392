"""A/B of mm_seg_moments tuning variants in one process (mm_reload_tuning between variants):
    python scripts/ab_moments.py [c2 northstar c5] > gpurun_out/ab_moments.json
Every variant's output is compared with the default's (the summation order does not depend on the variant: equal bits
expected) and timed with CUDA events over 20 launches after 3 warm-ups."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import torch
import memento_b200 as memento
from memento_b200 import synth, _lib

SHAPES = {"c2": dict(cells=25_000, genes=10_000, conditions=2, types=8, donors=1, q=0.07, labels=["stim", "cell"]),
          "northstar": dict(cells=1_000_000, genes=2_500, conditions=2, types=20, donors=1, q=0.07, labels=["stim", "cell"]),
          "c5": dict(cells=1_200_000, genes=2_500, conditions=2, types=20, donors=100, q=0.1, labels=["stim", "cell", "donor"])}
peak = 6565.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def entry(ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / peak, 3), "algorithmic_bytes": int(nbytes)}

VARIANTS = [
    {},
    {"MM_MOMENTS_KERNEL": "stream"},
    {"MM_MOMENTS_KERNEL": "stream_l1"},
    {"MM_MOMENTS_KERNEL": "stream", "MM_MOMENTS_CHUNK": "8"},
    {"MM_MOMENTS_KERNEL": "stream", "MM_MOMENTS_CHUNK": "8", "MM_MOMENTS_THREADS": "768"},
    {"MM_MOMENTS_KERNEL": "tile"},
]
KEYS = sorted({k for v in VARIANTS for k in v})


def main():
    out = {}
    for name in (sys.argv[1:] or ["c2"]):
        w = SHAPES[name]
        ad = synth.make_counts_fast(w["cells"], w["genes"], n_conditions=w["conditions"], n_types=w["types"], q=w["q"],
                                    seed=7, n_donors=w["donors"], device="cuda")
        memento.setup_memento(ad, "q")
        memento.create_groups(ad, w["labels"])
        memento.compute_1d_moments(ad, min_perc_group=0.7, filter_genes=False)
        st = ad.uns["memento"]["_b200"]
        seg = st.seg
        rows = []
        ref = None
        for var in VARIANTS:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(var)
            _lib.reload_tuning()
            res = seg.moments(st.inv_sf_sorted).clone()
            torch.cuda.synchronize()
            if ref is None:
                ref = res
            same = bool(torch.equal(res, ref))
            maxrel = float(((res - ref).abs() / ref.abs().clamp_min(1e-300)).max())
            e = entry(timed(lambda: seg.moments(st.inv_sf_sorted), reps=20), seg.moments_bytes())
            e.update(variant=var, bit_equal=same, max_rel_diff=maxrel)
            rows.append(e)
            print(name, var, e["ms"], e["frac"], same, file=sys.stderr, flush=True)
        out[name] = {"mean_segment": seg.nnz / max(1, seg.n_seg), "groups": seg.R, "variants": rows}
        for k in KEYS:
            os.environ.pop(k, None)
        _lib.reload_tuning()
        del ad, st, seg
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
