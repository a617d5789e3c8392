"""mm_seg_moments (register-streaming span kernel, TMA-staged tile kernel and the global-memory fallback) against a numpy restatement of
memento/estimator.py:175-185 on adversarial segment structures: empty segments at every position
(leading, trailing, at tile edges), segments spanning several 4096-nonzero tiles, segments that start or end
exactly on a tile edge, a ragged array end and an empty matrix."""
import numpy as np
import pytest
import torch

from memento_b200 import device as dev_mod
from memento_b200 import _lib

pytestmark = pytest.mark.gpu


def _reference(vals, rows, seg_ptr, inv_sf):
    n_seg = seg_ptr.size - 1
    seg_of = np.repeat(np.arange(n_seg), np.diff(seg_ptr))
    x = vals.astype(np.float64)
    w = inv_sf[rows]
    out = np.zeros((5, n_seg))
    out[0] = np.bincount(seg_of, x, n_seg)
    np.maximum.at(out[1], seg_of, x)
    out[2] = np.bincount(seg_of, x * w, n_seg)
    out[3] = np.bincount(seg_of, x * w * w, n_seg)
    out[4] = np.bincount(seg_of, x * x * w * w, n_seg)
    return out


def _run(lens, n_cells=5000, seed=0, use_tiles=True):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int64)
    seg_ptr = np.concatenate([[0], np.cumsum(lens)])
    nnz = int(seg_ptr[-1])
    vals = rng.integers(1, 40, nnz).astype(np.float32)
    rows = np.concatenate([np.sort(rng.choice(n_cells, min(int(l), n_cells), replace=False)) if l <= n_cells
                           else np.sort(rng.integers(0, n_cells, int(l))) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    inv_sf = 1.0 / rng.uniform(0.3, 3.0, n_cells)
    d = torch.device("cuda", 0)
    seg = dev_mod.SegMatrix(torch.as_tensor(vals, device=d), torch.as_tensor(rows, device=d),
                            torch.as_tensor(seg_ptr, device=d), lens.size, 1, n_cells)
    if not use_tiles:
        out = torch.empty(5 * seg.n_seg, dtype=torch.float64, device=d)
        big = torch.zeros(nnz // 4096 + 2, dtype=torch.int32, device=d)
        _lib.call("mm_seg_moments", d, seg.vals, seg.rows, seg.seg_ptr, seg.n_seg, seg.nnz,
                  torch.as_tensor(inv_sf, device=d), n_cells, out, big, None, None)
        got = out.view(5, -1).cpu().numpy()
    else:
        got = seg.moments(torch.as_tensor(inv_sf, device=d)).cpu().numpy()[:, :, 0]
    want = _reference(vals, rows, seg_ptr, inv_sf)
    np.testing.assert_array_equal(got[0], want[0])      # integer sums are exact
    np.testing.assert_array_equal(got[1], want[1])
    np.testing.assert_allclose(got[2:], want[2:], rtol=1e-12, atol=0)
    return got


CASES = {
    "tiny": [3, 0, 5, 1],
    "leading_trailing_empty": [0, 0, 0, 7, 0, 0, 9, 0, 0],
    "exact_tile_edges": [4096, 4096, 0, 0, 8192, 1, 4095, 0],
    "exact_span_edges": [512, 512, 0, 1024, 1, 511, 0, 0, 513, 511, 0, 512 * 3, 0],
    "window_of_31": [0] * 30 + [2] + [0] * 31 + [3, 0] + [1] * 70 + [0] * 33 + [600],
    "exact_chunk_edges": [2048, 2048, 0, 4096, 1, 2047, 0, 0, 2049, 2047, 512, 1536, 0],
    "multi_tile_segments": [10, 20000, 3, 0, 12289, 4093, 5, 0, 0, 30000, 2],
    "ragged_end": [4096 * 2 + 1],
    "ragged_end3": [5000, 4096 * 3 - 5000 + 3],
    "one_long_piece_per_tile": [1500, 1100, 1496, 1025, 1024, 1023, 2048, 40],
    "many_small": list(np.random.default_rng(1).integers(0, 12, 6000)),
    "more_than_staged_boundaries": [0] * 700 + [5] + [0] * 1300 + [4090, 3] + [0] * 600 + [1, 1],
    "mixed": list(np.random.default_rng(2).choice([0, 1, 7, 60, 400, 1300, 5000], 400)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_seg_moments_tile_kernel(name):
    a = _run(CASES[name], seed=3)
    b = _run(CASES[name], seed=3)
    np.testing.assert_array_equal(a, b)          # deterministic (no atomics)


@pytest.mark.parametrize("name", ["multi_tile_segments", "many_small", "mixed"])
def test_seg_moments_fallback_kernel(name):
    _run(CASES[name], seed=4, use_tiles=False)


@pytest.mark.parametrize("kernel", ["stream", "stream_l1", "tile"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_seg_moments_every_kernel(kernel, name, monkeypatch):
    """Each streaming kernel on every structure, whatever the size-based choice would have been."""
    monkeypatch.setenv("MM_MOMENTS_KERNEL", kernel)
    _run(CASES[name], seed=5)


@pytest.mark.parametrize("regime", ["0", "1"])
def test_seg_moments_tile_regimes(regime, monkeypatch):
    monkeypatch.setenv("MM_MOMENTS_KERNEL", "tile")
    monkeypatch.setenv("MM_MOMENTS_REGIME", regime)
    for name in ("mixed", "exact_tile_edges", "multi_tile_segments", "many_small"):
        _run(CASES[name], seed=6)


@pytest.mark.parametrize("name", ["exact_span_edges", "window_of_31", "multi_tile_segments", "ragged_end3", "mixed"])
def test_seg_moments_stream_single_span_chunks(name, monkeypatch):
    monkeypatch.setenv("MM_MOMENTS_KERNEL", "stream")
    monkeypatch.setenv("MM_MOMENTS_CHUNK", "1")
    _run(CASES[name], seed=8)


@pytest.mark.parametrize("name", ["exact_span_edges", "exact_chunk_edges", "window_of_31", "ragged_end3", "mixed"])
def test_seg_moments_stream_short_spans(name, monkeypatch):
    monkeypatch.setenv("MM_MOMENTS_KERNEL", "stream")
    monkeypatch.setenv("MM_MOMENTS_CHUNK", "8")
    monkeypatch.setenv("MM_MOMENTS_THREADS", "896")
    _run(CASES[name], seed=9)


def test_seg_moments_table_too_large_for_smem(monkeypatch):
    monkeypatch.setenv("MM_MOMENTS_KERNEL", "stream")
    _run(CASES["mixed"], n_cells=40000, seed=7)


@pytest.mark.parametrize("w", ["8", "16", "32"])
def test_seg_moments_lane_widths(w, monkeypatch):
    monkeypatch.setenv("MM_MOMENTS_W", w)
    _run(CASES["mixed"], seed=5, use_tiles=False)


def test_seg_moments_empty_matrix():
    got = _run([0, 0, 0, 0])
    assert not got.any()
