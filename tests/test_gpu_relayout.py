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


@pytest.mark.parametrize("shape", [(700, 90, 5, 0.2, ()), (5000, 300, 16, 0.05, (3, 15)), (64, 10, 1, 0.5, ()),
                                   (3000, 40, 400, 0.3, (7,)), (40000, 64, 3, 0.1, ())])
def test_relayout_equals_stable_sort(shape):
    n_cells, n_genes, R, density, empty = shape
    X, codes = _case(n_cells, n_genes, R, density, seed=n_cells, empty_groups=empty)
    d = torch.device("cuda", 0)
    csr = dev_mod.CsrOnDevice(X, d)
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
