"""Host-side helpers of the test engine that need no GPU: validity-mask partition, gene-tile planning."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))

from memento_b200 import engine  # noqa: E402


@pytest.mark.parametrize("n,R,p", [(1, 1, 0.5), (7, 3, 0.5), (3782, 16, 0.97), (32, 2001, 0.9), (500, 64, 0.99),
                                   (500, 65, 0.99), (100, 8, 1.0), (100, 8, 0.0)])
def test_distinct_masks_is_the_partition_of_np_unique(n, R, p):
    """The packed-bit partition of the per-gene validity masks = np.unique(axis=0)'s (reference
    hypothesis_test.py:249-251 drops groups per gene; genes with the same surviving groups share one weighted
    least-squares functional)."""
    rng = np.random.default_rng(n * 1000 + R)
    good = (rng.random((n, R)) < p).astype(np.uint8)
    first, inv = engine.distinct_masks(good)
    masks = good[first]
    want_masks, want_inv = np.unique(good, axis=0, return_inverse=True)
    assert masks.shape == want_masks.shape and masks.dtype == np.uint8
    assert inv.shape == (n,)
    np.testing.assert_array_equal(masks[inv], good)                       # every gene maps to its own mask
    assert len({m.tobytes() for m in masks}) == masks.shape[0]            # and the masks are distinct
    # same partition: genes share a mask here iff they share one there
    pairs = set(zip(inv.tolist(), np.asarray(want_inv).reshape(-1).tolist()))
    assert len(pairs) == masks.shape[0]


def test_distinct_masks_with_a_second_key_component():
    """With treatment_for_gene a design is (validity mask, id of the gene's treatment-column set): genes share a
    design iff both agree (reference main.py:368-373, :392 slices the treatment frame per gene)."""
    rng = np.random.default_rng(3)
    good = (rng.random((400, 6)) < 0.8).astype(np.uint8)
    extra = rng.integers(0, 5, size=400)
    first, inv = engine.distinct_masks(good, extra)
    np.testing.assert_array_equal(good[first][inv], good)
    np.testing.assert_array_equal(extra[first][inv], extra)
    keys = {(good[i].tobytes(), int(extra[i])) for i in range(400)}
    assert len(keys) == first.size


def test_tile_plan_respects_grid_and_workspace_limits():
    B = 10000
    per_seg = 32 * (B + 1)
    assert engine.tile_plan_groups(16, B, 6 << 30) == (6 << 30) // per_seg // 16          # workspace-bound
    assert engine.tile_plan_groups(16, B, 1 << 40) == 65535 // 16                         # grid-bound (65535 rows)
    assert engine.tile_plan_groups(4000, B, 24 << 30) == 65535 // 4000
    assert engine.tile_plan_groups(70000, B, 24 << 30) == 1                               # one gene always fits the plan
    assert engine.tile_plan_groups(16, B, 1) == 1


def test_trend_fit_from_power_sums_equals_polyfit():
    """main._fit_mv_sums (3 + 8 all-reduced numbers in the gene-sharded runs) against np.polyfit on the pairs
    (reference estimator.py:84-93), including entries the reference drops (mean or variance <= 0)."""
    from memento_b200 import main as mmain
    rng = np.random.default_rng(4)
    mean = np.exp(rng.normal(-1.0, 1.8, 50000))
    var = np.exp(1.3 * np.log(mean) + 0.04 * np.log(mean) ** 2 + 0.2 + rng.normal(0, 0.5, mean.size))
    mean[::97] = 0.0
    var[::89] = -1.0
    want = mmain._fit_mv(mean, var)
    got = mmain._fit_mv_sums(mean, var)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    # sharded: the sums of the parts are the sums of the whole
    class TwoRanks:
        def __init__(self, other):
            self.other, self.calls = other, 0
        def all_reduce_sum(self, a):
            self.calls += 1
            return a + self.other[self.calls - 1]
    keep = (mean > 0) & (var > 0)
    x, y = np.log(mean[keep]), np.log(var[keep])
    h = x.size // 3
    xa, ya, xb, yb = x[:h], y[:h], x[h:], y[h:]
    n, xbar = x.size, x.mean()
    scale = np.sqrt((x * x).mean() - xbar * xbar)
    zb = (xb - xbar) / scale
    other = [np.array([xb.size, xb.sum(), (xb * xb).sum()]),
             np.array([zb.sum(), (zb ** 2).sum(), (zb ** 3).sum(), (zb ** 4).sum(), yb.sum(), (yb * zb).sum(), (yb * zb ** 2).sum()])]
    part = mmain._fit_mv_sums(np.exp(xa), np.exp(ya), TwoRanks(other))
    np.testing.assert_allclose(part, want, rtol=1e-9, atol=1e-11)


def test_pair_index_helpers_equal_the_reference_loops():
    """main._pair_indices / _first_unordered (vectorised) against the reference's per-pair Python loops
    (main.py:310-318 name look-up, :467-482 frozenset de-duplication)."""
    import pandas as pd
    from memento_b200 import main as mmain
    rng = np.random.default_rng(8)
    names = pd.Index(["g%d" % i for i in range(50)])
    pairs = [("g%d" % a, "g%d" % b) for a, b in rng.integers(0, 50, size=(3000, 2))]
    i1, i2 = mmain._pair_indices(names, pairs)
    assert i1.tolist() == [int(a[1:]) for a, _ in pairs] and i2.tolist() == [int(b[1:]) for _, b in pairs]
    j1, j2 = mmain._pair_indices(names, np.asarray(pairs, dtype=object))
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2)
    with pytest.raises(KeyError):
        mmain._pair_indices(names, [("g1", "nope")])
    owner, uniq = mmain._first_unordered(i1, i2)
    first, want = {}, np.full(len(pairs), -1)
    for k, (a, b) in enumerate(zip(i1, i2)):
        if a == b:
            continue
        key = frozenset((int(a), int(b)))
        first.setdefault(key, k)
        want[k] = first[key]
    assert np.array_equal(owner, want)
    assert uniq.tolist() == sorted(set(first.values()))
    assert mmain._pair_indices(names, [])[0].size == 0


def test_pair_indices_product_fast_path_and_lazy_arrays():
    """itertools.product(A, B) as an array takes the |A| + |B| look-up; near-products (one entry changed, a ragged
    tail) fall back to the per-name look-up with the same answer.  LazyArrays materialises on access only."""
    import copy
    import itertools
    import pandas as pd
    from memento_b200 import main as mmain
    names = pd.Index(["g%d" % i for i in range(300)])
    A, B = ["g%d" % i for i in range(5, 75)], ["g%d" % i for i in range(100, 170)]
    prod = np.array(list(itertools.product(A, B)))
    assert prod.shape[0] >= mmain.DENSE_BLOCK_MIN_PAIRS
    want1 = np.array([int(a[1:]) for a, _ in prod]); want2 = np.array([int(b[1:]) for _, b in prod])
    for arr in (prod, prod.astype(object)):
        i1, i2 = mmain._pair_indices(names, arr)
        assert np.array_equal(i1, want1) and np.array_equal(i2, want2)
    broken = prod.copy(); broken[777, 1] = "g3"; want2b = want2.copy(); want2b[777] = 3
    i1, i2 = mmain._pair_indices(names, broken)
    assert np.array_equal(i1, want1) and np.array_equal(i2, want2b)
    i1, i2 = mmain._pair_indices(names, prod[:-3])
    assert np.array_equal(i1, want1[:-3]) and np.array_equal(i2, want2[:-3])
    with pytest.raises(KeyError):
        bad = prod.copy(); bad[:, 1][bad[:, 1] == "g100"] = "zzz"
        mmain._pair_indices(names, bad)
    calls = []
    lz = mmain.LazyArrays({"cov": lambda: calls.append("cov") or np.arange(3.0), "corr": lambda: calls.append("corr") or np.ones(3)})
    assert set(lz.keys()) == {"cov", "corr"} and calls == []
    assert lz["cov"].tolist() == [0.0, 1.0, 2.0] and lz["cov"] is lz["cov"] and calls == ["cov"]
    cp = copy.deepcopy(lz)
    assert calls == ["cov"] and cp["cov"] is not lz["cov"]
    assert [k for k, _ in cp.items()] == ["cov", "corr"] and calls == ["cov", "corr"]
    assert lz.get("nope", 7) == 7


@pytest.mark.parametrize("shuffled", [False, True])
def test_first_unordered_block_equals_the_sorting_version(shuffled):
    """The analytic de-duplication of a full block A x B (A and B overlapping, so mirror pairs exist) against
    _first_unordered, in product order and in a shuffled pair order."""
    from memento_b200 import main as mmain
    rng = np.random.default_rng(3)
    ga = np.sort(rng.choice(400, 70, replace=False)); gb = np.sort(np.concatenate([ga[::2], rng.choice(np.arange(400, 500), 40, replace=False)]))
    idx1, idx2 = np.repeat(ga, gb.size), np.tile(gb, ga.size)
    if shuffled:
        p = rng.permutation(idx1.size)
        idx1, idx2 = idx1[p], idx2[p]
    blk = mmain._as_dense_block(idx1, idx2)
    assert blk is not None and (blk[2] is None) == (not shuffled)
    want_owner, want_uniq = mmain._first_unordered(idx1, idx2)
    owner, uniq = mmain._first_unordered_block(*blk)
    assert np.array_equal(owner, want_owner) and np.array_equal(uniq, want_uniq)
