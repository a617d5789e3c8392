"""Ingest re-layout (csrc/relayout.cu: stable counting transposition CSR -> group-sorted CSC) against the
stable-sort construction it replaces: identical arrays, bit for bit, on matrices with empty rows, empty columns,
empty groups, one-cell groups and one group."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from memento_b200 import device as dev_mod

pytestmark = pytest.mark.gpu


def _case(n_cells, n_genes, n_groups, density, seed, empty_groups=()):
    rng = np.random.default_rng(seed)
    X = sp.random(n_cells, n_genes, density=density, format="csr", random_state=seed, dtype=np.float32)
    X.data = np.ceil(X.data * 9).astype(np.float32)
    if n_cells > 5:
        X = X.tolil(); X[3, :] = 0; X[:, 2] = 0; X = X.tocsr(); X.eliminate_zeros()
    X.sort_indices()
    pool = [g for g in range(n_groups) if g not in empty_groups]
    codes = rng.choice(pool, n_cells).astype(np.int32)
    return X, codes


def _shuffle_rows(X, seed):
    """Same matrix with the entries of every row in random order (legal, non-canonical CSR)."""
    rng = np.random.default_rng(seed)
    indices, data = X.indices.copy(), X.data.copy()
    for r in range(X.shape[0]):
        lo, hi = X.indptr[r], X.indptr[r + 1]
        p = rng.permutation(hi - lo)
        indices[lo:hi], data[lo:hi] = indices[lo:hi][p], data[lo:hi][p]
    return sp.csr_matrix((data, indices, X.indptr.copy()), shape=X.shape)


# (cells, genes, groups, density, empty groups): the tiled path (sorted rows; csrc/relayout.cu relayout_tile_kernel)
# sees gene counts below / across / far above its 256-gene blocks, chunks of exactly 256 rows and ragged ones, blocks
# denser than its 12288-element staging buffer (several passes), groups smaller than a chunk, one group
SHAPES = [(700, 90, 5, 0.2, ()), (5000, 300, 16, 0.05, (3, 15)), (64, 10, 1, 0.5, ()),
          (3000, 40, 400, 0.3, (7,)), (40000, 64, 3, 0.1, ()), (1024, 256, 1, 0.97, ()), (2100, 1300, 4, 0.6, (2,)),
          (513, 2050, 2, 0.02, ()), (9000, 700, 7, 0.25, ()),
          # more than 24 576 genes: the row scan packs its counters as 16-bit halves; 160 000 rows = 625 chunks: one CTA
          # per chunk instead of a cluster
          (600, 30000, 3, 0.01, ()), (160000, 48, 5, 0.1, (1,))]


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_relayout_equals_stable_sort(shape, canonical):
    n_cells, n_genes, R, density, empty = shape
    X, codes = _case(n_cells, n_genes, R, density, seed=n_cells, empty_groups=empty)
    if not canonical:
        if n_cells > 5000 or n_genes > 5000:
            pytest.skip("host-side row shuffle of the large case")
        X = _shuffle_rows(X, seed=1)
    d = torch.device("cuda", 0)
    csr = dev_mod.CsrOnDevice(X, d)
    assert csr.sorted_rows == canonical
    order = np.argsort(codes, kind="stable")
    rank = np.empty_like(order); rank[order] = np.arange(order.size)
    gs = np.concatenate([[0], np.cumsum(np.bincount(codes, minlength=R))]).astype(np.int64)
    new = dev_mod.SegMatrix.from_csr_grouped(csr, order, gs)
    old = dev_mod.SegMatrix.from_csr(csr, torch.as_tensor(codes, device=d), R, torch.as_tensor(rank.astype(np.int32), device=d))
    assert torch.equal(new.seg_ptr, old.seg_ptr)
    assert torch.equal(new.rows, old.rows)
    assert torch.equal(new.vals, old.vals)
    assert np.array_equal(new.group_start_host, old.group_start_host)
    # one group, original order (the all-cells matrix of setup_memento)
    new1 = dev_mod.SegMatrix.from_csr_grouped(csr)
    old1 = dev_mod.SegMatrix.from_csr(csr)
    assert torch.equal(new1.seg_ptr, old1.seg_ptr) and torch.equal(new1.rows, old1.rows) and torch.equal(new1.vals, old1.vals)


def test_threaded_upload_equals_plain_copy():
    """mm_upload (host threads -> pinned ring -> cudaMemcpyAsync on side streams) against torch's own copy: sizes
    below / at / above the chunk size and not a multiple of it, back-to-back calls reusing the ring, and ordering
    with work queued on the caller's stream before and after."""
    d = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    for n in (1, 1000, (8 << 20) // 4, 5 * (8 << 20) // 4 + 123, 40_000_003):
        a = rng.integers(0, 1 << 30, size=n, dtype=np.int32)
        out = torch.full((n,), -1, dtype=torch.int32, device=d)          # queued before: must not land after the copy
        dev_mod._lib.call("mm_upload", d, out, int(a.ctypes.data), a.nbytes, 3)
        doubled = out.to(torch.int64) * 2                                    # queued after: must see the data
        np.testing.assert_array_equal(out.cpu().numpy(), a)
        np.testing.assert_array_equal(doubled.cpu().numpy(), a.astype(np.int64) * 2)
    big = rng.random(9_000_000)
    t = dev_mod.to_device(big, d, np.float64)                               # 72 MB: goes through mm_upload
    np.testing.assert_array_equal(t.cpu().numpy(), big)
    dev_mod._lib.load().mm_upload_release()
    t2 = dev_mod.to_device(big, d, np.float32)                              # ring re-created
    np.testing.assert_array_equal(t2.cpu().numpy(), big.astype(np.float32))


def test_csr_canonical_check_on_device():
    X, _ = _case(500, 60, 3, 0.2, seed=1)
    d = torch.device("cuda", 0)
    assert dev_mod.CsrOnDevice(X, d).sorted_rows
    assert not dev_mod.CsrOnDevice(_shuffle_rows(X, 2), d).sorted_rows
    dup = sp.csr_matrix((np.ones(4, np.float32), np.array([0, 3, 3, 5], np.int32), np.array([0, 4], np.int64)), shape=(1, 8))
    assert not dev_mod.CsrOnDevice(dup, d).sorted_rows                      # duplicate column entries are not canonical
