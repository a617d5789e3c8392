"""On-disk formats either side of the hot path.

Input: a 10x Genomics ``matrix.mtx`` directory (Matrix Market genes x cells + barcodes / features TSV, plain or
gzipped) -> the cells x genes float32 CSR ``AnnDataLite`` that ``setup_memento`` takes (the reference's callers read
``.h5ad`` with scanpy, reference requirements.txt:1; this image has neither ``anndata`` nor ``h5py``).
Output: the numeric results of ``uns['memento']`` as one ``.npz`` (what the reference's ``prepare_to_save``,
main.py:673-683, prepares for scanpy's h5ad writer -- minus the pickled regressor strings, kept as arrays here).
Host code only: nothing here touches the device.
"""
import gzip
import json
import os

import numpy as np
import pandas as pd
import scipy.io
import scipy.sparse as sp

from .anndata_lite import AnnDataLite


def _open_maybe_gz(path):
    for p in (path, path + ".gz"):
        if os.path.exists(p):
            return gzip.open(p, "rt") if p.endswith(".gz") else open(p, "rt")
    raise FileNotFoundError(path)


def _existing(directory, names):
    for n in names:
        for p in (os.path.join(directory, n), os.path.join(directory, n + ".gz")):
            if os.path.exists(p):
                return p
    raise FileNotFoundError("none of %s in %s" % (names, directory))


def read_10x_mtx(directory, var_names="gene_symbols"):
    """``directory``/matrix.mtx[.gz] (genes x cells), barcodes.tsv[.gz], features.tsv[.gz] or genes.tsv[.gz].
    Counts must be exactly representable in float32 (integers below 2^24), as everywhere on the device path."""
    m = scipy.io.mmread(_existing(directory, ["matrix.mtx"]))
    X = sp.csr_matrix(m.T)                                   # cells x genes
    data32 = X.data.astype(np.float32)
    if not np.array_equal(data32.astype(X.data.dtype), X.data):
        raise ValueError("counts are not exactly representable in float32")
    X = sp.csr_matrix((data32, X.indices, X.indptr), shape=X.shape)      # index arrays are converted at upload
    X.sum_duplicates()
    X.sort_indices()
    with _open_maybe_gz(_existing(directory, ["barcodes.tsv"]).removesuffix(".gz")) as f:
        barcodes = [ln.rstrip("\n").split("\t")[0] for ln in f if ln.strip()]
    with _open_maybe_gz(_existing(directory, ["features.tsv", "genes.tsv"]).removesuffix(".gz")) as f:
        feats = [ln.rstrip("\n").split("\t") for ln in f if ln.strip()]
    if len(barcodes) != X.shape[0] or len(feats) != X.shape[1]:
        raise ValueError("matrix is %d genes x %d cells, but %d features and %d barcodes"
                         % (X.shape[1], X.shape[0], len(feats), len(barcodes)))
    ids = [r[0] for r in feats]
    symbols = [r[1] if len(r) > 1 else r[0] for r in feats]
    names = symbols if var_names == "gene_symbols" else ids
    if len(set(names)) != len(names):                        # scanpy's var_names_make_unique convention
        seen, uniq = {}, []
        for n in names:
            k = seen.get(n, 0)
            uniq.append(n if k == 0 else "%s-%d" % (n, k))
            seen[n] = k + 1
        names = uniq
    var = pd.DataFrame({"gene_ids": ids, "gene_symbols": symbols}, index=pd.Index(names))
    obs = pd.DataFrame(index=pd.Index(barcodes))
    return AnnDataLite(X, obs, var)


def write_10x_mtx(adata, directory, compress=False):
    """The inverse of :func:`read_10x_mtx` (integer Matrix Market, genes x cells)."""
    os.makedirs(directory, exist_ok=True)
    X = sp.csr_matrix(adata.X)
    counts = sp.coo_matrix((X.data.astype(np.int64), X.nonzero()), shape=X.shape).T
    path = os.path.join(directory, "matrix.mtx")
    scipy.io.mmwrite(path, counts, field="integer")
    opener = (lambda p: gzip.open(p + ".gz", "wt")) if compress else (lambda p: open(p, "wt"))
    with opener(os.path.join(directory, "barcodes.tsv")) as f:
        f.write("".join("%s\n" % b for b in adata.obs.index))
    ids = adata.var["gene_ids"] if "gene_ids" in adata.var else adata.var.index
    with opener(os.path.join(directory, "features.tsv")) as f:
        f.write("".join("%s\t%s\tGene Expression\n" % (i, s) for i, s in zip(ids, adata.var.index)))
    if compress:
        with open(path, "rb") as src, gzip.open(path + ".gz", "wb") as dst:
            dst.write(src.read())
        os.remove(path)


_RESULT_KEYS = ("1d_ht", "2d_ht")


def save_results(adata, path):
    """Numeric contents of ``uns['memento']`` -> ``path`` (.npz): per-group 1D moments, the mean-variance fits, gene
    filters, the hypothesis-test arrays and the design frames; scalars and names in a JSON header.  Device state
    and the per-group matrix views are not written (cf. ``prepare_to_save``)."""
    mem = adata.uns["memento"]
    arrays, meta = {}, {"groups": list(mem.get("groups", [])), "gene_list": list(map(str, adata.var.index)),
                        "scalars": {k: mem[k] for k in ("q_column", "all_q", "estimator_type", "filter_mean_thresh",
                                                         "num_bins", "label_delimiter") if k in mem},
                        "label_columns": list(mem.get("label_columns", []))}
    for g, trio in mem.get("1d_moments", {}).items():
        for name, a in zip(("mean", "var", "res_var"), trio):
            arrays["1d_moments/%s/%s" % (g, name)] = np.asarray(a)
    for g, fit in mem.get("mv_regressor", {}).items():
        if not isinstance(fit, str):
            arrays["mv_regressor/%s" % g] = np.asarray(fit, dtype=np.float64)
    for key in ("gene_filter", "gene_rv_filter"):
        for g, a in mem.get(key, {}).items():
            arrays["%s/%s" % (key, g)] = np.asarray(a)
    if "overall_gene_filter" in mem:
        arrays["overall_gene_filter"] = np.asarray(mem["overall_gene_filter"])
    for key in _RESULT_KEYS:
        for name, a in mem.get(key, {}).items():
            if isinstance(a, pd.DataFrame):
                arrays["%s/%s/values" % (key, name)] = a.values.astype(np.float64)
                meta.setdefault("frames", {})["%s/%s" % (key, name)] = {"index": list(map(str, a.index)),
                                                                        "columns": list(map(str, a.columns))}
            elif isinstance(a, np.ndarray):
                arrays["%s/%s" % (key, name)] = a
    if "2d_moments" in mem:
        for name in ("gene_idx_1", "gene_idx_2"):
            arrays["2d_moments/%s" % name] = np.asarray(mem["2d_moments"][name])
        for g in meta["groups"]:
            for name, a in mem["2d_moments"].get(g, {}).items():
                arrays["2d_moments/%s/%s" % (g, name)] = np.asarray(a)
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_results(path):
    """Inverse of :func:`save_results`: a nested dict in the layout of ``uns['memento']`` (numpy arrays, DataFrames
    for the design frames) plus ``gene_list``."""
    z = np.load(path, allow_pickle=False)
    meta = json.loads(bytes(z["__meta__"]).decode())
    out = dict(meta["scalars"])
    out.update(groups=meta["groups"], gene_list=meta["gene_list"], label_columns=meta["label_columns"])
    for key in z.files:
        if key == "__meta__":
            continue
        parts = key.split("/")
        a = z[key]
        if parts[0] == "1d_moments":
            slot = out.setdefault("1d_moments", {}).setdefault(parts[1], [None, None, None])
            slot[("mean", "var", "res_var").index(parts[2])] = a
        elif parts[0] in _RESULT_KEYS and parts[-1] == "values":
            fr = meta["frames"]["/".join(parts[:2])]
            out.setdefault(parts[0], {})[parts[1]] = pd.DataFrame(a, index=fr["index"], columns=fr["columns"])
        elif len(parts) == 1:
            out[key] = a
        elif len(parts) == 2:
            out.setdefault(parts[0], {})[parts[1]] = a
        else:
            out.setdefault(parts[0], {}).setdefault(parts[1], {})[parts[2]] = a
    return out
