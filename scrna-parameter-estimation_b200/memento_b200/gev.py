"""GEV tail refinement of small ASLs (reference hypothesis_test.py:94-141) -- device stage."""


def refine_tail_asl(res, device, timer, stats):
    """Placeholder until the device GEV kernel lands: tests with <= 10 extreme replicates keep the
    empirical bound (c + 1) / (n + 1), which is also the reference's own fallback (:119, :136, :141)."""
    stats["gev_pending"] = stats.get("gev_pending", 0) + int(((res["extreme"] >= 0) & (res["extreme"] <= 10)).sum().item())
