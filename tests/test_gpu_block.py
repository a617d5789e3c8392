"""Dense gene-block covariance on the tensor cores (csrc/block.cu: mm_block_panels + mm_block_gemm, tcgen05 /
TMEM / TMA) against float64 numpy on the same inputs, and through compute_2d_moments against the per-pair
float64 path and the oracle.  Tolerance: the north star asks 1e-5 relative "in float" for point estimates;
for a covariance that may be arbitrarily close to zero the scale is sqrt(var_a var_b), i.e. correlation
units.  Held here: 5e-6 (fp16 hi/lo split keeps 22 bits; fp32 accumulation in tensor memory is limited to 256-cell chunks, float64 beyond)."""
import itertools

import numpy as np
import pytest
import torch

import memento_b200 as memento
from memento_b200 import device as dev_mod
from memento_b200 import main as mm_main
from helpers import golden_adata

pytestmark = pytest.mark.gpu


def _random_matrix(n_cells_per_group, n_genes, seed, density=0.3):
    rng = np.random.default_rng(seed)
    gs = np.concatenate([[0], np.cumsum(n_cells_per_group)]).astype(np.int64)
    n_cells, R = int(gs[-1]), len(n_cells_per_group)
    dense = rng.poisson(rng.gamma(0.6, 3.0, size=(1, n_genes)), size=(n_cells, n_genes)) * (rng.random((n_cells, n_genes)) < density)
    dense[:, 3] = 0                                   # a gene that is zero everywhere
    dense[gs[1]:gs[2], 5] = 0                         # ... and one that is zero in one group
    dense[:, 7] = 4                                   # constant (zero variance)
    vals, rows, seg_ptr = [], [], [0]
    for g in range(n_genes):
        for r in range(R):
            col = dense[gs[r]:gs[r + 1], g]
            nz = np.flatnonzero(col)
            vals.append(col[nz].astype(np.float32)); rows.append((nz + gs[r]).astype(np.int32))
            seg_ptr.append(seg_ptr[-1] + nz.size)
    d = torch.device("cuda", 0)
    seg = dev_mod.SegMatrix(torch.as_tensor(np.concatenate(vals), device=d), torch.as_tensor(np.concatenate(rows), device=d),
                            torch.as_tensor(np.asarray(seg_ptr, dtype=np.int64), device=d), n_genes, R, n_cells, gs)
    sf = rng.uniform(0.4, 2.5, n_cells)
    return seg, dense.astype(np.float64), sf, gs


@pytest.mark.parametrize("sizes,n_genes,na,nb", [([70, 200, 129], 150, 150, 150), ([64, 1], 40, 13, 40),
                                                   ([300, 517, 90, 1000], 300, 129, 257)])
def test_block_cross_vs_numpy(sizes, n_genes, na, nb):
    seg, dense, sf, gs = _random_matrix(sizes, n_genes, seed=len(sizes))
    d = seg.device
    inv_sf = torch.as_tensor(1.0 / sf, device=d)
    sums = seg.moments(inv_sf)
    rng = np.random.default_rng(0)
    idx_a = np.sort(rng.choice(n_genes, na, replace=False))
    idx_b = idx_a if na == nb else np.sort(rng.choice(n_genes, nb, replace=False))
    got = seg.block_cross(idx_a, idx_b, inv_sf, sums).cpu().numpy()
    for r in range(len(sizes)):
        u = dense[gs[r]:gs[r + 1]] / sf[gs[r]:gs[r + 1], None]
        z = u - u.mean(0, keepdims=True)
        want = z[:, idx_a].T @ z[:, idx_b]
        scale = np.sqrt(np.outer((z[:, idx_a] ** 2).sum(0), (z[:, idx_b] ** 2).sum(0)))
        err = np.abs(got[r] - want)
        assert (err <= 5e-6 * scale + 1e-12).all(), (r, float((err / (scale + 1e-300)).max()))


def test_compute_2d_moments_dense_block_vs_pair_path(monkeypatch):
    calls = []
    orig = dev_mod.SegMatrix.block_cross
    monkeypatch.setattr(dev_mod.SegMatrix, "block_cross", lambda self, *a, **k: (calls.append(1), orig(self, *a, **k))[1])
    ad = golden_adata()
    memento.setup_memento(ad, "q")
    memento.create_groups(ad, ["stim", "cell"])
    memento.compute_1d_moments(ad, min_perc_group=0.7)
    names = ad.var.index.tolist()
    rng = np.random.default_rng(1)
    monkeypatch.setattr(mm_main, "DENSE_BLOCK_MIN_PAIRS", 64)       # the golden matrix is small
    A = [names[i] for i in rng.choice(len(names), min(40, len(names)), replace=False)]
    B = [names[i] for i in rng.permutation(len(names))]             # every gene: includes the same-gene pairs
    pairs = list(itertools.product(A, B))
    shuffled = [pairs[i] for i in rng.permutation(len(pairs))]      # any order of the block is recognised
    for plist in (pairs, shuffled):
        dense_ad, pair_ad = ad.copy(), ad.copy()
        n_before = len(calls)
        memento.compute_2d_moments(dense_ad, plist)
        assert len(calls) == n_before + 1          # the tensor-core path ran
        mm_main.DENSE_BLOCK_MIN_PAIRS = 1 << 62
        try:
            memento.compute_2d_moments(pair_ad, plist)
        finally:
            mm_main.DENSE_BLOCK_MIN_PAIRS = 64
        assert len(calls) == n_before + 1          # ... and the per-pair path did not use it
        st = dense_ad.uns["memento"]["_b200"]
        sums = st.seg.moments(st.inv_sf_sorted).cpu().numpy()                  # (5, G, R)
        n_g = np.diff(st.group_start).astype(float)
        plain_var = sums[4] / n_g - (sums[2] / n_g) ** 2                       # second moment of x / sf about its mean
        i1, i2 = dense_ad.uns["memento"]["2d_moments"]["gene_idx_1"], dense_ad.uns["memento"]["2d_moments"]["gene_idx_2"]
        for r, g in enumerate(ad.uns["memento"]["groups"]):
            a, b = dense_ad.uns["memento"]["2d_moments"][g], pair_ad.uns["memento"]["2d_moments"][g]
            scale = np.sqrt(plain_var[i1, r] * plain_var[i2, r])              # what the GEMM error is relative to
            assert (np.abs(a["cov"] - b["cov"]) <= 5e-6 * scale + 1e-14).all(), g
            # the estimator's variances subtract the sampling noise, so the correlation amplifies the error for
            # genes whose corrected variance is a small part of the plain one
            ok = np.isfinite(b["corr"])
            assert np.array_equal(np.isfinite(a["corr"]), ok)
            amp = scale[ok] / np.sqrt(np.abs(b["var_1"][ok] * b["var_2"][ok]))
            clipped = np.abs(b["corr"][ok]) == 1
            assert (np.abs(a["corr"][ok] - b["corr"][ok])[~clipped] <= 5e-6 * amp[~clipped] + 1e-12).all(), g
