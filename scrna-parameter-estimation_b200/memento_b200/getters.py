"""Result getters (host-side pandas views over ``uns['memento']``; reference main.py:156-168,
:523-582, :635-655).  Same outputs; ``get_groups`` avoids ``pd.to_numeric(errors='ignore')``, which
pandas 3 removed."""
import itertools

import numpy as np
import pandas as pd


def get_groups(adata):
    mem = adata.uns["memento"]
    rows = [g.split(mem["label_delimiter"])[1:] for g in mem["groups"]]
    df = pd.DataFrame(rows, index=mem["groups"], columns=mem["label_columns"])
    for col in df.columns:
        try:
            df[col] = pd.to_numeric(df[col])
        except (ValueError, TypeError):
            pass
    return df


def get_1d_moments(adata):
    mem = adata.uns["memento"]
    mean_df = pd.DataFrame({"gene": adata.var.index.tolist()})
    var_df = pd.DataFrame({"gene": adata.var.index.tolist()})
    counts = {k: v.shape[0] for k, v in mem["group_cells"].items()}
    with np.errstate(divide="ignore", invalid="ignore"):
        for group, val in mem["1d_moments"].items():
            mean_df[group] = np.log(val[0])
            var_df[group] = np.log(val[2])
    return mean_df, var_df, counts


def get_1d_ht_result(adata):
    ht = adata.uns["memento"]["1d_ht"]
    if "treatment_for_gene" in ht:
        df = pd.concat([pd.DataFrame(itertools.product([g], ht["treatment_for_gene"][g]), columns=["gene", "tx"])
                        for g in adata.var.index])
    else:
        df = pd.DataFrame(itertools.product(adata.var.index, ht["treatment"].columns), columns=["gene", "tx"])
    df["de_coef"], df["de_se"], df["de_pval"] = ht["mean_coef"], ht["mean_se"], ht["mean_asl"]
    df["dv_coef"], df["dv_se"], df["dv_pval"] = ht["var_coef"], ht["var_se"], ht["var_asl"]
    return df
