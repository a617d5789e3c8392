"""Synthetic scRNA-seq count generator for tests and the bench.

Negative-binomial transcript counts thinned by a capture rate q: a gamma-Poisson draw with mean
``q * mu_g * s_c * effect`` and gene dispersion ``phi_g`` (binomial thinning of an NB keeps the
dispersion), cell scale ``s_c ~ LogNormal(0, 0.3)``, ``mu_g ~ LogNormal(1.0, 1.4)``,
``phi_g ~ LogNormal(-0.5, 0.5)``.  The first ``de_frac`` of the genes get ``log_fc`` in the
treated condition; cell types shift a random fifth of the genes.  Host-side numpy only: this
is workload generation, not part of the hot path.
"""
import numpy as np
import pandas as pd
import scipy.sparse as sp

from .anndata_lite import AnnDataLite


def make_counts(n_cells, n_genes, n_conditions=2, n_types=1, q=0.07, de_frac=0.1, log_fc=0.5,
                seed=7, chunk=2048, n_donors=1):
    rng = np.random.default_rng(seed)
    mu = rng.lognormal(1.0, 1.4, n_genes)
    phi = rng.lognormal(-0.5, 0.5, n_genes)
    n_de = int(de_frac * n_genes)
    cond_eff = np.zeros((n_conditions, n_genes))
    for c in range(1, n_conditions):
        cond_eff[c, :n_de] = log_fc * c
    type_eff = np.zeros((n_types, n_genes))
    for t in range(1, n_types):
        sel = rng.random(n_genes) < 0.2
        type_eff[t, sel] = rng.normal(0, 0.7, sel.sum())
    cond = rng.integers(0, n_conditions, n_cells)
    ctype = rng.integers(0, n_types, n_cells)
    donor = rng.integers(0, n_donors, n_cells)
    scale = rng.lognormal(0.0, 0.3, n_cells)
    blocks = []
    shape = 1.0 / phi
    for lo in range(0, n_cells, chunk):
        hi = min(lo + chunk, n_cells)
        lam = q * mu[None, :] * scale[lo:hi, None] * np.exp(cond_eff[cond[lo:hi]] + type_eff[ctype[lo:hi]])
        rate = rng.gamma(shape[None, :], lam / shape[None, :])
        blocks.append(sp.csr_matrix(rng.poisson(rate).astype(np.float32)))
    X = sp.vstack(blocks).tocsr()
    X.sort_indices()
    obs = pd.DataFrame({
        "stim": np.where(cond == 0, "ctrl", "stim") if n_conditions == 2 else cond.astype(str),
        "cell": np.array(["ct%d" % t for t in ctype]),
        "donor": np.array(["d%d" % d for d in donor]),
        "q": np.full(n_cells, q),
    }, index=pd.Index(["c%d" % i for i in range(n_cells)]))
    var = pd.DataFrame(index=pd.Index(["gene%d" % i for i in range(n_genes)]))
    return AnnDataLite(X, obs, var)


def design_from_groups(groups, label_columns, delimiter="^", treatment_col="stim", treated="stim",
                       covariate_cols=None):
    """Group-level design frames in the order of ``uns['memento']['groups']``:
    treatment = indicator of ``treated``; covariate = one-hot of the other label columns (drop
    first) or a column of ones when there are none."""
    rows = [g.split(delimiter)[1:] for g in groups]
    df = pd.DataFrame(rows, columns=label_columns, index=groups)
    treatment = pd.DataFrame({treatment_col: (df[treatment_col] == treated).astype(float)}, index=groups)
    covariate_cols = [c for c in label_columns if c != treatment_col] if covariate_cols is None else covariate_cols
    if covariate_cols:
        cov = pd.get_dummies(df[covariate_cols], drop_first=True).astype(float)
        if cov.shape[1] == 0:
            cov = pd.DataFrame({"intercept": np.ones(len(groups))}, index=groups)
    else:
        cov = pd.DataFrame({"intercept": np.ones(len(groups))}, index=groups)
    return cov, treatment


def make_counts_fast(n_cells, n_genes, n_conditions=2, n_types=1, q=0.07, de_frac=0.1, log_fc=0.5, seed=7,
                     chunk=4096, n_donors=1, device=None, shard=0):
    """Same model as :func:`make_counts`, with the gamma-Poisson draws done by torch (on ``device``
    when given, e.g. the GPU for the bench's 25k x 10k matrix).  Cell-level parameters depend on
    ``seed`` only, gene-level parameters and the counts on (``seed``, ``shard``): different shards are
    different gene blocks of the SAME cells (used for the gene-sharded multi-GPU runs)."""
    import torch
    rng = np.random.default_rng([seed, 1 + shard])
    mu = rng.lognormal(1.0, 1.4, n_genes)
    phi = rng.lognormal(-0.5, 0.5, n_genes)
    n_de = int(de_frac * n_genes)
    cond_eff = np.zeros((n_conditions, n_genes))
    for c in range(1, n_conditions):
        cond_eff[c, :n_de] = log_fc * c
    type_eff = np.zeros((n_types, n_genes))
    for t in range(1, n_types):
        sel = rng.random(n_genes) < 0.2
        type_eff[t, sel] = rng.normal(0, 0.7, sel.sum())
    crng = np.random.default_rng([seed, 0])
    cond = crng.integers(0, n_conditions, n_cells)
    ctype = crng.integers(0, n_types, n_cells)
    donor = crng.integers(0, n_donors, n_cells)
    scale = crng.lognormal(0.0, 0.3, n_cells)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed * 1000003 + shard)
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)  # noqa: E731
    mu_t, shape_t = t(q * mu), t(1.0 / phi)
    ce, te, sc = t(np.exp(cond_eff)), t(np.exp(type_eff)), t(scale)
    cond_t, type_t = torch.as_tensor(cond, device=dev), torch.as_tensor(ctype, device=dev)
    data, indices, counts_per_row = [], [], []
    for lo in range(0, n_cells, chunk):
        hi = min(lo + chunk, n_cells)
        lam = mu_t[None, :] * sc[lo:hi, None] * ce[cond_t[lo:hi]] * te[type_t[lo:hi]]
        g = torch._standard_gamma(shape_t[None, :].expand(hi - lo, -1).contiguous(), generator=gen)
        x = torch.poisson(g * lam / shape_t[None, :], generator=gen)
        nz = x.nonzero(as_tuple=False)
        data.append(x[nz[:, 0], nz[:, 1]].cpu().numpy().astype(np.float32))
        indices.append(nz[:, 1].to(torch.int32).cpu().numpy())
        counts_per_row.append(torch.bincount(nz[:, 0], minlength=hi - lo).cpu().numpy())
    indptr = np.concatenate([[0], np.cumsum(np.concatenate(counts_per_row))]).astype(np.int64)
    X = sp.csr_matrix((np.concatenate(data), np.concatenate(indices), indptr), shape=(n_cells, n_genes))
    obs = pd.DataFrame({
        "stim": np.where(cond == 0, "ctrl", "stim") if n_conditions == 2 else cond.astype(str),
        "cell": np.array(["ct%d" % v for v in ctype]),
        "donor": np.array(["d%d" % d for d in donor]),
        "q": np.full(n_cells, q),
    }, index=pd.Index(["c%d" % i for i in range(n_cells)]))
    var = pd.DataFrame(index=pd.Index(["gene%d" % (i + shard * n_genes) for i in range(n_genes)]))
    return AnnDataLite(X, obs, var)
