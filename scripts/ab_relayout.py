"""Kernel-level timing of the ingest re-layout (csrc/relayout.cu) on the BASELINE shapes:
    python scripts/ab_relayout.py [c2 northstar c5] > gpurun_out/ab_relayout.json
The chunk plan and every buffer are prepared once; mm_relayout_count (row scan + chunk prefix) and mm_relayout_fill are
timed separately with CUDA events (10 launches after 2 warm-ups), for the all-cells matrix (one group) and the grouped
one.  Algorithmic bytes: count 4 B per nonzero (indices), fill 16 B (indices + values in, values + rows out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import numpy as np
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
    for _ in range(2):
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


def plan_and_time(csr, order, gs):
    n_cells, n_genes = csr.shape
    dev = csr.device
    R = gs.size - 1
    sizes = np.diff(gs)
    rpc = 256
    per_group = (sizes + rpc - 1) // rpc
    gcl = np.concatenate([[0], np.cumsum(per_group)]).astype(np.int32)
    n_chunks = int(gcl[-1])
    cg = np.repeat(np.arange(R, dtype=np.int32), per_group)
    within = np.arange(n_chunks, dtype=np.int64) - gcl[cg]
    crl = np.concatenate([gs[cg] + within * rpc, [n_cells]]).astype(np.int32)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)  # noqa: E731
    crl, cg, gcl = d(crl), d(cg), d(gcl)
    order_d = None if order is None else d(np.asarray(order, dtype=np.int32))
    n_blocks = (n_genes + 127) // 128
    cnt = torch.empty(max(n_chunks, 1) * n_genes, dtype=torch.int32, device=dev)
    seg_ptr = torch.zeros(n_genes * R + 1, dtype=torch.int64, device=dev)
    seg_len = torch.zeros(n_genes * R, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    bnd = torch.empty((n_blocks + 1) * n_cells, dtype=torch.int32, device=dev)
    vals = torch.empty(csr.nnz, dtype=torch.float32, device=dev)
    rows = torch.empty(csr.nnz, dtype=torch.int32, device=dev)

    def count():
        _lib.call("mm_relayout_count", dev, csr.indptr, csr.indices, order_d, crl, cg, gcl, n_chunks, n_genes, R,
                  cnt, seg_len, 1, err, bnd, n_cells)
    count()
    torch.cumsum(seg_len, 0, out=seg_ptr[1:])

    def fill():
        _lib.call("mm_relayout_fill", dev, csr.indptr, csr.indices, csr.data, order_d, crl, cg, n_chunks, n_genes, R,
                  cnt, seg_ptr, vals, rows, 1, err, bnd, n_cells)
    t_fill = timed(fill)          # cnt holds the prefix after count(): fill can be repeated
    t_count = timed(count)
    assert int(err.item()) == 0
    return {"chunks": n_chunks, "groups": R, "count": entry(t_count, csr.nnz * 4), "fill": entry(t_fill, csr.nnz * 16),
            "total": entry(t_count + t_fill, csr.nnz * 20)}


def main():
    out = {"hbm_peak_GB/s": peak}
    for name in (sys.argv[1:] or ["c2"]):
        w = SHAPES[name]
        ad = synth.make_counts_fast(w["cells"], w["genes"], n_conditions=w["conditions"], n_types=w["types"], q=w["q"],
                                    seed=7, n_donors=w["donors"], device="cuda")
        memento.setup_memento(ad, "q")
        st = ad.uns["memento"]["_b200"]
        csr = st.csr
        res = {"cells": w["cells"], "genes": w["genes"], "nnz": int(csr.nnz)}
        for cfg in os.environ.get("AB_RELAYOUT_CFGS", "2").split(","):
            for thr in os.environ.get("AB_RELAYOUT_SCAN_THREADS", "256").split(","):
                os.environ["MM_RELAYOUT_CFG"] = cfg
                os.environ["MM_RELAYOUT_SCAN_THREADS"] = thr
                _lib.reload_tuning()
                tag = "all_cells_cfg%s_scan%s" % (cfg, thr)
                res[tag] = plan_and_time(csr, None, np.asarray([0, w["cells"]], dtype=np.int64))
                print(name, "cfg", cfg, "scan threads", thr, json.dumps(res[tag]), file=sys.stderr, flush=True)
        os.environ.pop("MM_RELAYOUT_CFG"); os.environ.pop("MM_RELAYOUT_SCAN_THREADS")
        _lib.reload_tuning()
        R = w["conditions"] * w["types"] * w["donors"]
        codes = np.random.default_rng(0).integers(0, R, size=w["cells"]).astype(np.int32)
        order = np.argsort(codes, kind="stable")
        gs = np.concatenate([[0], np.cumsum(np.bincount(codes, minlength=R))]).astype(np.int64)
        res["grouped"] = plan_and_time(csr, order, gs)
        out[name] = res
        print(name, json.dumps(res), file=sys.stderr, flush=True)
        del ad, st, csr
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
