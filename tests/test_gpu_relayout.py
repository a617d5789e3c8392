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
          (513, 2050, 2, 0.02, ()), (9000, 700, 7, 0.25, ())]


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_relayout_equals_stable_sort(shape, canonical):
    n_cells, n_genes, R, density, empty = shape
    X, codes = _case(n_cells, n_genes, R, density, seed=n_cells, empty_groups=empty)
    if not canonical:
        if n_cells > 5000:
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
