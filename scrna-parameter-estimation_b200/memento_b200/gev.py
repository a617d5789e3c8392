"""GEV tail refinement of small ASLs (reference hypothesis_test.py:94-141) -- device stage."""
import torch

from . import _lib

MAX_EXTREME = 10   # reference hypothesis_test.py:90


def refine_tail_asl(res, device, timer, stats):
    """Tests with <= 10 extreme bootstrap replicates get their ASL from fitted GEV tails
    (mm_gev_tail_asl); the others keep (c + 1) / (n + 1).  ``res`` is the dict returned by
    engine.regress_tile with ``coef_rows`` present; ``res['asl']`` is updated in place.  The tests are
    selected on the device, so this enqueues one launch and never waits for the GPU (the caller may run it on a
    side stream, concurrently with the next tile's bootstrap)."""
    ext = res["extreme"].reshape(-1)
    asl = res["asl"].reshape(-1)
    n_rows = int(asl.numel())
    if n_rows == 0:
        return
    rows = res["coef_rows"]
    num_boot = rows.shape[-1] - 1
    status = torch.empty(n_rows, dtype=torch.int32, device=device)
    ev = timer.start()
    _lib.call("mm_gev_tail_asl", device, rows, None, n_rows, num_boot, asl, status, ext, MAX_EXTREME)
    timer.stop("gev_tail_asl", ev)
    stats["launches"] = stats.get("launches", 0) + 1
    res["gev_status"] = status          # -1 not a tail test, 0 empirical bound kept, 1 GEV tails used
    stats.setdefault("_gev_status", []).append(status)


def count_tail_tests(stats):
    """Number of tests that went through the GEV stage (reads the device status vectors: call after the work
    has been waited for)."""
    stats["gev_tests"] = stats.get("gev_tests", 0) + sum(int((s >= 0).sum().item()) for s in stats.pop("_gev_status", []))
