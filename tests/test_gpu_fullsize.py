"""Size-independent properties at the FULL size of BASELINE.json's configs[1] (25 000 cells x 10 000 genes, 16
groups, ~59M nonzeros; the oracle would need hours here): checksums of the re-layout and of the moment pass,
agreement of the three moment kernels, and invariance of the test results to how the genes are tiled."""
import numpy as np
import pytest
import torch

import memento_b200 as memento
from memento_b200 import device as dev_mod
from memento_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
    X = ad.X
    memento.setup_memento(ad, "q")
    memento.create_groups(ad, ["stim", "cell"])
    st = ad.uns["memento"]["_b200"]
    seg_before_filter = st.seg
    memento.compute_1d_moments(ad, min_perc_group=0.7)
    return ad, X, seg_before_filter


def test_relayout_checksums_full_size(full):
    ad, X, seg = full
    st = ad.uns["memento"]["_b200"]
    assert seg.nnz == X.nnz and int(seg.seg_ptr[-1]) == X.nnz
    assert float(seg.vals.double().sum()) == float(X.data.astype(np.float64).sum())        # integers: exact
    # rows ascend inside every segment and stay inside the segment's group
    seg_of = torch.repeat_interleave(torch.arange(seg.n_seg, device=seg.device), seg.seg_ptr[1:] - seg.seg_ptr[:-1])
    same = seg_of[1:] == seg_of[:-1]
    assert bool((seg.rows[1:][same] > seg.rows[:-1][same]).all())
    gs = torch.as_tensor(st.group_start, device=seg.device)
    grp = seg_of % seg.R
    assert bool(((seg.rows >= gs[grp]) & (seg.rows < gs[grp + 1])).all())
    # per-gene totals equal the column sums of the input
    col = torch.zeros(seg.G, dtype=torch.float64, device=seg.device).index_add_(0, seg_of // seg.R, seg.vals.double())
    np.testing.assert_array_equal(col.cpu().numpy(), np.asarray(X.sum(axis=0)).ravel().astype(np.float64))


def test_moment_kernels_agree_full_size(full, monkeypatch):
    ad, X, seg = full
    st = ad.uns["memento"]["_b200"]
    outs = {}
    for kern in ("stream", "stream_l1", "tile"):
        monkeypatch.setenv("MM_MOMENTS_KERNEL", kern)
        a = seg.moments(st.inv_sf_sorted).cpu().numpy()
        b = seg.moments(st.inv_sf_sorted).cpu().numpy()
        np.testing.assert_array_equal(a, b)                  # deterministic
        outs[kern] = a
    monkeypatch.delenv("MM_MOMENTS_KERNEL")
    big = torch.zeros(seg.nnz // 4096 + 2, dtype=torch.int32, device=seg.device)
    out = torch.empty(5 * seg.n_seg, dtype=torch.float64, device=seg.device)
    dev_mod._lib.call("mm_seg_moments", seg.device, seg.vals, seg.rows, seg.seg_ptr, seg.n_seg, seg.nnz,
                      st.inv_sf_sorted, int(st.inv_sf_sorted.numel()), out, big, None, None)
    outs["per_segment"] = out.view(5, seg.G, seg.R).cpu().numpy()
    ref = outs["per_segment"]
    assert ref[0].sum() == float(X.data.astype(np.float64).sum())      # checksum of checksums (exact: integers)
    for kern, a in outs.items():
        np.testing.assert_array_equal(a[0], ref[0], err_msg=kern)      # sum x
        np.testing.assert_array_equal(a[1], ref[1], err_msg=kern)      # max x
        np.testing.assert_allclose(a[2:], ref[2:], rtol=1e-12, atol=0, err_msg=kern)


def test_ht_1d_invariant_to_gene_tiling_full_size(full):
    """The same genes tested in one tile or in many small tiles give bit-identical results (RNG streams are
    keyed by global gene id and group, statistics are per gene): 300 genes, all 25 000 cells, B = 2000."""
    ad, _, _ = full
    sub = ad.copy()
    genes = sub.var.index[:300].tolist()
    memento.compute_1d_moments(sub, min_perc_group=0.7, gene_list=genes)
    cov, tr = synth.design_from_groups(sub.uns["memento"]["groups"], ["stim", "cell"])
    res = []
    for ws in (6 << 30, 64 << 20):
        memento.ht_1d_moments(sub, cov, tr, num_boot=2000, resampling="bootstrap", seed=11, workspace_bytes=ws)
        res.append({k: sub.uns["memento"]["1d_ht"][k].copy() for k in ("mean_coef", "mean_se", "mean_asl", "var_coef", "var_se", "var_asl")})
    for k in res[0]:
        # coefficients are per-gene sums in a fixed order: bit-identical.  SE / normal-tail ASL add the replicate
        # columns in an order that depends on how many CTAs share a gene (mm_regress_asl splits small tiles): round-off
        if k.endswith("coef"):
            np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)
        else:
            np.testing.assert_allclose(res[0][k], res[1][k], rtol=1e-9, atol=0, equal_nan=True, err_msg=k)
    assert np.isfinite(res[0]["mean_asl"]).mean() > 0.9
