"""mm_seg_unique (csrc/unique.cu) on segments that reach every tier of the kernel: one warp (<= 768 nonzeros), the
2048-slot shared-memory table of a CTA, its overflow into the 8192-slot table, and the global-memory table for
segments with more than 6144 distinct (count, bin) keys -- against np.unique over the same device arrays
(reference bootstrap.py:40-71, _unique_expr)."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp
import torch

import memento_b200 as memento
from memento_b200 import engine
from memento_b200.anndata_lite import AnnDataLite

pytestmark = pytest.mark.gpu

# (nonzeros, largest count): distinct keys ~ min(nonzeros, largest count x size-factor bins present)
SPECS = [(500, 5), (768, 2000), (769, 3), (5000, 4), (5000, 3000), (6144, 9), (6145, 9), (20000, 6), (20000, 120),
         (20000, 600), (30000, 4000), (29999, 2)]


def test_unique_tiers_vs_numpy():
    rng = np.random.default_rng(0)
    n = 30000
    cols = []
    for nnz, hi in SPECS:
        x = np.zeros(n, dtype=np.float32)
        x[rng.choice(n, nnz, replace=False)] = rng.integers(1, hi + 1, nnz)
        cols.append(x)
    filler = rng.poisson(rng.gamma(2.0, 1.0, size=60)[None, :] * rng.uniform(0.5, 2.0, size=n)[:, None]).astype(np.float32)
    X = sp.csr_matrix(np.concatenate([np.stack(cols, axis=1), filler], axis=1))
    obs = pd.DataFrame({"q": np.full(n, 0.07), "g": ["a"] * n})
    var = pd.DataFrame(index=pd.Index(["g%d" % i for i in range(X.shape[1])]))
    ad = AnnDataLite(X, obs, var)
    memento.setup_memento(ad, "q", filter_mean_thresh=0.0)
    memento.create_groups(ad, ["g"])
    memento.compute_1d_moments(ad, min_perc_group=0.0, filter_genes=False)
    st = ad.uns["memento"]["_b200"]
    seg = st.seg
    assert seg.R == 1 and seg.G == X.shape[1]
    tab = engine.unique_tables(seg, st.design, st.cell_bin, 0, seg.G, 0, want_raw=True)
    torch.cuda.synchronize()
    ptr = seg.seg_ptr.cpu().numpy()
    vals, rows = seg.vals.cpu().numpy(), seg.rows.cpu().numpy()
    bins = st.cell_bin.cpu().numpy()
    key = tab["raw_key"].cpu().numpy().view(np.uint32)
    cnt = tab["raw_cnt"].cpu().numpy()
    seg_U = tab["seg_U"].cpu().numpy()
    seen = set()
    for gene in range(len(SPECS)):
        lo, hi = ptr[gene], ptr[gene + 1]
        assert hi - lo == SPECS[gene][0]
        want_k, want_c = np.unique((vals[lo:hi].astype(np.uint32) << 8) | bins[rows[lo:hi]].astype(np.uint32), return_counts=True)
        U = int(seg_U[gene])
        assert U == want_k.size, (gene, SPECS[gene], U, want_k.size)
        assert np.array_equal(key[lo:lo + U], want_k), (gene, SPECS[gene])       # sorted by key
        assert np.array_equal(cnt[lo:lo + U], want_c), (gene, SPECS[gene])
        seen.add("warp" if hi - lo <= 768 else "small" if U <= 1536 else "full" if U <= 6144 else "global")
    assert seen == {"warp", "small", "full", "global"}, seen
