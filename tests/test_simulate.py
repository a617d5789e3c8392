"""memento_b200.simulate (device-side workload generation with the reference's simulate.py semantics) on the CPU
torch backend: negative-binomial moments, exactness properties of the hypergeometric capture step and its
distribution against numpy's multivariate_hypergeometric (what reference simulate.py:105-110 calls per cell)."""
import numpy as np
import scipy.stats as stats
import torch

from memento_b200 import simulate


def test_nb_transcriptomes_have_the_requested_moments():
    means = np.array([0.5, 3.0, 20.0, 100.0])
    variances = means + np.array([0.3, 0.8, 0.2, 0.05]) * means ** 2
    x = simulate.simulate_transcriptomes(200000, means, variances, seed=3).numpy().astype(np.float64)
    assert x.shape == (200000, 4) and (x >= 0).all()
    np.testing.assert_allclose(x.mean(axis=0), means, rtol=0.02)
    np.testing.assert_allclose(x.var(axis=0), variances, rtol=0.05)
    # variance below the mean: the dispersion is floored (simulate.py:60-61), i.e. essentially Poisson
    y = simulate.simulate_transcriptomes(100000, np.array([5.0]), np.array([2.0]), seed=1).numpy()
    assert abs(y.var() / y.mean() - 1.0) < 0.03


def test_hypergeometric_capture_is_exact_sampling_without_replacement():
    rng = np.random.default_rng(0)
    t = torch.as_tensor(rng.poisson(rng.lognormal(1.0, 1.2, size=(1, 40)), size=(300, 40)))
    t[5] = 0                                             # an empty cell
    qs, c = simulate.capture_sampling(t, 0.07, process="hyper", seed=9)
    c, tn = c.numpy(), t.numpy()
    assert (c >= 0).all() and (c <= tn).all()
    np.testing.assert_array_equal(c.sum(axis=1), np.round(0.07 * tn.sum(axis=1)).astype(int))
    assert np.allclose(qs.numpy(), 0.07)
    # chunked and unchunked runs draw different keys but obey the same constraints
    _, c2 = simulate.capture_sampling(t, 0.07, process="hyper", seed=9, chunk_molecules=1 << 10)
    np.testing.assert_array_equal(c2.numpy().sum(axis=1), c.sum(axis=1))


def test_hypergeometric_capture_distribution_matches_numpy():
    colors = np.array([30, 5, 0, 120, 1, 44])
    n_rep, n_sample = 20000, 40                          # q = 0.2 of 200 molecules
    t = torch.as_tensor(np.tile(colors, (n_rep, 1)))
    _, c = simulate.capture_sampling(t, 0.2, process="hyper", seed=4)
    c = c.numpy()
    ref = np.random.Generator(np.random.PCG64(42343)).multivariate_hypergeometric(colors, n_sample, size=n_rep)
    assert (c.sum(axis=1) == n_sample).all() and (c[:, 2] == 0).all()
    for j in (0, 1, 3, 5):
        assert stats.ks_2samp(c[:, j], ref[:, j]).pvalue > 1e-3, j
    # the joint matters too: covariance of two genes' captured counts is negative, -n K1 K2 (N - n) / (N^2 (N - 1))
    N = colors.sum()
    want = -n_sample * colors[0] * colors[3] * (N - n_sample) / (N ** 2 * (N - 1.0))
    assert abs(np.cov(c[:, 0], c[:, 3])[0, 1] - want) < 0.15 * abs(want)


def test_poisson_capture_and_beta_rates():
    t = torch.full((50000, 3), 100)
    qs, c = simulate.capture_sampling(t, 0.1, q_sq=0.012, process="poisson", seed=2)
    qs, c = qs.numpy(), c.numpy()
    assert abs(qs.mean() - 0.1) < 2e-3 and abs((qs ** 2).mean() - 0.012) < 3e-4      # Beta with the given two moments
    np.testing.assert_allclose(c.mean(axis=0), 10.0, rtol=0.02)


def test_extract_parameters_matches_the_formulas():
    import scipy.sparse as sp
    rng = np.random.default_rng(1)
    X = sp.csr_matrix(rng.poisson(0.8, size=(400, 30)).astype(np.float64))
    (xm, xv), (zm, zv), Nc, good = simulate.extract_parameters(X, q=0.1)
    tot = np.asarray(X.sum(axis=1)).ravel()
    D = X.toarray()
    m1 = (D / tot[:, None]).mean(axis=0)
    m2 = ((D / tot[:, None]) ** 2).mean(axis=0) - 0.9 * (D / tot[:, None] ** 2).mean(axis=0)
    np.testing.assert_allclose(xm, m1[good], rtol=1e-12)
    np.testing.assert_allclose(xv, (m2 - m1 ** 2)[good], rtol=1e-9)
    np.testing.assert_allclose(Nc, tot / 0.1)
    np.testing.assert_allclose(zm, xm * Nc.mean(), rtol=1e-12)
