"""Workload generation with the reference's ``memento.simulate`` semantics, on the device.

Same function names and argument meaning as reference memento/simulate.py: negative-binomial transcript counts
(``simulate_transcriptomes``, :52-67 -- the ``norm_cov`` string form, i.e. independent genes) followed by capture
sampling (``capture_sampling``, :91-115): ``process='hyper'`` draws, for every cell, ``round(q * total)`` of its
molecules WITHOUT replacement (multivariate hypergeometric over the genes), ``process='poisson'`` draws
Poisson(q * count).  The reference does the hypergeometric draw in a Python loop over the cells (:105-110) with numpy's
sampler, which cannot produce a 1 M-cell data set; here every molecule gets a uniform random key and the
``n_sample`` smallest keys of a cell are the captured molecules -- the same distribution, as a sort, a k-th value and a
segmented count in torch (any device; the bench uses the GPU).  This is data generation, not the hot path: plain
torch, no hand-written kernels.  The copula form (a covariance matrix as ``norm_cov``) is not provided.
"""
import numpy as np
import torch


def gamma_params_from_moments(m, v):
    """reference simulate.py:37-39."""
    return m ** 2 / v, v / m


def convert_params_nb(mu, theta):
    """Mean / inverse-dispersion -> (r, p) of scipy's nbinom.  reference simulate.py:42-50."""
    r = theta
    var = mu + 1 / r * mu ** 2
    p = (var - mu) / var
    return r, 1 - p


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def simulate_transcriptomes(n_cells, means, variances, Nc=None, norm_cov="indep", device=None, seed=0, chunk=8192):
    """(n_cells, n_genes) int64 transcript counts, NB with the given per-gene means / variances (dispersions
    ``(v - m) / m^2``, floored at 1e-5 as the reference does).  reference simulate.py:52-67.  ``Nc`` is unused in this
    form there too.  Returns a torch tensor on ``device``."""
    if not isinstance(norm_cov, str):
        raise NotImplementedError("the Gaussian-copula form (simulate.py:69-88) is not provided on the device")
    dev = torch.device(device) if device is not None else torch.device("cpu")
    means = np.asarray(means, dtype=np.float64)
    disp = (np.asarray(variances, dtype=np.float64) - means) / means ** 2
    disp[disp < 0] = 1e-5
    theta = torch.as_tensor(1.0 / disp, dtype=torch.float64, device=dev)
    mu = torch.as_tensor(means, dtype=torch.float64, device=dev)
    g = _gen(dev, seed)
    out = torch.empty((n_cells, means.shape[0]), dtype=torch.int64, device=dev)
    for lo in range(0, n_cells, chunk):
        hi = min(n_cells, lo + chunk)
        # NB(theta, p) = Poisson(Gamma(shape theta, scale mu / theta))
        lam = torch._standard_gamma(theta[None, :].expand(hi - lo, -1).contiguous(), generator=g) * (mu / theta)[None, :]
        out[lo:hi] = torch.poisson(lam, generator=g).to(torch.int64)
    return out


def _hyper_rows(counts, n_sample, g):
    """Multivariate hypergeometric draw per row: ``n_sample[i]`` of row i's ``counts[i].sum()`` molecules, without
    replacement.  counts (n, G) int64, n_sample (n,) int64."""
    n, G = counts.shape
    tot = counts.sum(dim=1)
    n_mol = int(tot.max().item()) if n else 0
    if n_mol == 0:
        return torch.zeros_like(counts)
    # molecule m of row i belongs to gene searchsorted(cumsum_i, m, right); keys ~ U(0, 1), padding = +inf
    keys = torch.rand((n, n_mol), generator=g, device=counts.device, dtype=torch.float32)
    col = torch.arange(n_mol, device=counts.device)[None, :]
    keys = torch.where(col < tot[:, None], keys, torch.full_like(keys, float("inf")))
    n_sample = torch.minimum(n_sample, tot)
    order = torch.argsort(keys, dim=1)                       # exact selection of the n_sample smallest keys (ties: none
    rank = torch.empty_like(order)                           # matter -- float keys, and any tie-break is uniform)
    rank.scatter_(1, order, col.expand(n, -1).contiguous())
    taken = (rank < n_sample[:, None]).to(torch.int64)
    csum = torch.cumsum(taken, dim=1)
    edges = torch.cumsum(counts, dim=1)                      # end (exclusive) of every gene's molecules
    csum = torch.cat([torch.zeros((n, 1), dtype=torch.int64, device=counts.device), csum], dim=1)
    upto = torch.gather(csum, 1, edges)
    return torch.diff(upto, dim=1, prepend=torch.zeros((n, 1), dtype=torch.int64, device=counts.device))


def capture_sampling(transcriptomes, q, q_sq=None, process="hyper", seed=42343, chunk_molecules=1 << 27):
    """``(qs, captured)``: per-cell capture rates and the captured counts.  reference simulate.py:91-115.
    ``q_sq`` (second moment of the capture rate) makes the rates Beta distributed as there.  ``transcriptomes``: torch
    tensor or array (n_cells, n_genes) of integer counts; the result lives on its device."""
    t = transcriptomes if isinstance(transcriptomes, torch.Tensor) else torch.as_tensor(np.asarray(transcriptomes))
    t = t.to(torch.int64)
    dev = t.device
    n = t.shape[0]
    g = _gen(dev, seed)
    if q_sq is None:
        qs = torch.full((n,), float(q), dtype=torch.float64, device=dev)
    else:
        m, v = float(q), float(q_sq) - float(q) ** 2
        alpha, beta = m * (m * (1 - m) / v - 1), (1 - m) * (m * (1 - m) / v - 1)
        ga = torch._standard_gamma(torch.full((n,), alpha, dtype=torch.float64, device=dev), generator=g)
        gb = torch._standard_gamma(torch.full((n,), beta, dtype=torch.float64, device=dev), generator=g)
        qs = ga / (ga + gb)
    if process != "hyper":
        return qs, torch.poisson(t.to(torch.float64) * qs[:, None], generator=g).to(torch.int64)
    tot = t.sum(dim=1)
    n_sample = torch.round(qs * tot.to(torch.float64)).to(torch.int64)
    out = torch.empty_like(t)
    rows = max(1, int(chunk_molecules // max(1, int(tot.max().item()) if n else 1)))
    for lo in range(0, n, rows):
        out[lo:lo + rows] = _hyper_rows(t[lo:lo + rows], n_sample[lo:lo + rows], g)
    return qs, out


def extract_parameters(data, q=0.1, min_mean=0.001):
    """Moments of a real data set for the simulator.  reference simulate.py:13-34 (hypergeometric estimator with the raw
    UMI totals as size factors).  ``data``: scipy sparse (cells x genes).  Host numpy: a one-off over a template."""
    import scipy.sparse as sp
    data = sp.csr_matrix(data)
    n = data.shape[0]
    total = np.asarray(data.sum(axis=1)).ravel()
    w = sp.diags(1.0 / total)
    m1 = np.asarray((w @ data).sum(axis=0)).ravel() / n
    m2 = np.asarray((w.power(2) @ data.power(2)).sum(axis=0)).ravel() / n \
        - (1 - q) * np.asarray((w.power(2) @ data).sum(axis=0)).ravel() / n
    x_mean, x_var = m1, m2 - m1 ** 2
    good = np.where(np.asarray(data.mean(axis=0)).ravel() > min_mean)[0]
    Nc = total / q
    z_mean = x_mean * Nc.mean()
    z_var = (x_var + x_mean ** 2) * (Nc ** 2).mean() - x_mean ** 2 * Nc.mean() ** 2
    return (x_mean[good], x_var[good]), (z_mean[good], z_var[good]), Nc, good
