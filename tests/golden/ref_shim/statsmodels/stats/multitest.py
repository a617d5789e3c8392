import numpy as np


def fdrcorrection(pvals, alpha=0.05):
    """Benjamini-Hochberg, enough for the reference's util._fdrcorrect (off the hot path)."""
    p = np.asarray(pvals, dtype=float)
    n = p.size
    order = np.argsort(p)
    ranked = p[order] * n / np.arange(1, n + 1)
    ranked = np.minimum.accumulate(ranked[::-1])[::-1]
    out = np.empty(n)
    out[order] = np.minimum(ranked, 1.0)
    return out <= alpha, out
