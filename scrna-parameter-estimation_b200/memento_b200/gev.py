"""GEV tail refinement of small ASLs (reference hypothesis_test.py:94-141) -- device stage."""
import torch

from . import _lib

MAX_EXTREME = 10   # reference hypothesis_test.py:90


def refine_tail_asl(res, device, timer, stats):
    """Tests with <= 10 extreme bootstrap replicates get their ASL from fitted GEV tails
    (mm_gev_tail_asl); the others keep (c + 1) / (n + 1).  ``res`` is the dict returned by
    engine.regress_tile with ``coef_rows`` present; ``res['asl']`` is updated in place."""
    ext = res["extreme"].reshape(-1)
    asl = res["asl"].reshape(-1)
    flag = (ext >= 0) & (ext <= MAX_EXTREME) & torch.isfinite(asl)
    idx = flag.nonzero().reshape(-1).to(torch.int32)
    n = int(idx.numel())
    stats["gev_tests"] = stats.get("gev_tests", 0) + n
    if n == 0:
        return
    rows = res["coef_rows"]
    num_boot = rows.shape[-1] - 1
    status = torch.zeros(n, dtype=torch.int32, device=device)
    ev = timer.start()
    _lib.call("mm_gev_tail_asl", device, rows, idx, n, num_boot, asl, status)
    timer.stop("gev_tail_asl", ev)
    stats["launches"] = stats.get("launches", 0) + 1
    res["gev_status"] = status
    res["gev_rows"] = idx
