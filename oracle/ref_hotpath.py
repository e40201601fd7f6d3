"""TEST / BENCH INFRASTRUCTURE ONLY - drives the reference's OWN hot path (unmodified modules, loaded by
`oracle/ref_shim.py`) on a prepared workload and times its stages.  This is the block the engine replaces:

    liana/method/sc/_liana_pipe.py:402-425
        perms = _get_means_perms(adata, n_perms, seed, agg_fun=np.mean, norm_factor=None, n_jobs, verbose)
        ligand_idx, receptor_idx, source_idx, target_idx = _get_mat_idx(adata, lr_res)
        perm_stats = np.stack((perms[:, source_idx, ligand_idx], perms[:, target_idx, receptor_idx]), axis=0)
        scores = _score.fun(x=lr_res, perm_stats=perm_stats)          # _cpdb_score, _cellphonedb.py:9-30

Every stage but `_get_mat_idx` is linear in the number of permutations, `_get_mat_idx` is linear in the number of rows
of `lr_res` (a Python dictionary comprehension over the rows with an `np.where` over `var_names` each,
`_get_mean_perms.py:90-100`); a bounded sample (a chunk of permutations, a slice of rows) is timed and scaled.
The reference's own parallelism is `n_jobs` (joblib / loky over permutations, `_get_mean_perms.py:78-81`); its worker
processes receive `_permute_and_aggregate` by value (cloudpickle), because they cannot import the reference's package
`__init__` (scanpy, plotnine, ... are not installed) - the function itself is the reference's, unmodified.
"""
from __future__ import annotations

import time

import numpy as np
import pandas as pd

from . import ref_shim


def available() -> bool:
    return ref_shim.reference_available()


_NS = None


def _ref():
    global _NS
    if _NS is None:
        _NS = ref_shim.load_reference()
        import cloudpickle
        cloudpickle.register_pickle_by_value(_NS.get_mean_perms)
    return _NS


def make_inputs(X, labels, n_clusters, var_names, tri, cat_names=None):
    """(adata stand-in holding the gene-subset matrix + `@label`, lr_res frame with the columns the block reads)."""
    ns = _ref()
    cats = [f"c{i:02d}" for i in range(n_clusters)] if cat_names is None else list(cat_names)
    obs = pd.DataFrame({"@label": pd.Categorical.from_codes(np.asarray(labels, dtype=np.int64), categories=cats)})
    adata = ns.AnnData(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(np.asarray(var_names, dtype=str))))
    names = np.asarray(var_names, dtype=str)
    cat_arr = np.asarray(cats, dtype=object)
    lr_res = pd.DataFrame({
        "ligand": names[tri.lig_gene], "receptor": names[tri.rec_gene],
        "source": cat_arr[tri.src], "target": cat_arr[tri.tgt],
        "ligand_means": tri.ligand_means.copy(), "receptor_means": tri.receptor_means.copy(),
    })
    return adata, lr_res


def run_block(adata, lr_res, n_perms, seed, n_jobs=1, idx_rows=None, idx_reuse=None):
    """One pass of `_liana_pipe.py:402-425` for cellphonedb.  `idx_rows`: time `_get_mat_idx` on the first `idx_rows`
    rows only (its cost is linear in the rows; the index arrays of the whole frame, which the gather needs, then come
    from a vectorised lookup that is NOT timed).  `idx_reuse` = (seconds, rows) of an earlier measurement of
    `_get_mat_idx` on the same frame: it does not depend on the permutations and is not repeated in every bench step.
    Returns (lr_means, pvals, seconds per stage)."""
    ns = _ref()
    g = ns.get_mean_perms
    t = {}
    t0 = time.perf_counter()
    perms = g._get_means_perms(adata=adata, n_perms=n_perms, seed=seed, agg_fun=np.mean, norm_factor=None,
                               n_jobs=n_jobs, verbose=False)
    t["get_means_perms"] = time.perf_counter() - t0
    T = len(lr_res)
    t0 = time.perf_counter()
    if idx_reuse is None and (idx_rows is None or idx_rows >= T):
        ligand_idx, receptor_idx, source_idx, target_idx = g._get_mat_idx(adata, lr_res)
        t["get_mat_idx"] = time.perf_counter() - t0
        t["get_mat_idx_rows"] = T
    else:
        if idx_reuse is None:
            g._get_mat_idx(adata, lr_res.iloc[:idx_rows])
            t["get_mat_idx"] = time.perf_counter() - t0
            t["get_mat_idx_rows"] = int(idx_rows)
        else:
            t["get_mat_idx"], t["get_mat_idx_rows"] = float(idx_reuse[0]), int(idx_reuse[1])
        var = pd.Index(adata.var_names)
        cats = pd.Index(adata.obs["@label"].cat.categories)
        ligand_idx, receptor_idx = var.get_indexer(lr_res["ligand"]), var.get_indexer(lr_res["receptor"])
        source_idx, target_idx = cats.get_indexer(lr_res["source"]), cats.get_indexer(lr_res["target"])
    t0 = time.perf_counter()
    ligand_stat_perms = perms[:, source_idx, ligand_idx]
    receptor_stat_perms = perms[:, target_idx, receptor_idx]
    perm_stats = np.stack((ligand_stat_perms, receptor_stat_perms), axis=0)
    t["gather"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    lr_means, pvals = ns.cellphonedb_mod._cpdb_score(x=lr_res, perm_stats=perm_stats)
    t["score"] = time.perf_counter() - t0
    return lr_means, pvals, t


def full_run_seconds(t, n_perms_sample, n_perms, n_rows):
    """Extrapolation of the stage times of `run_block` to the full workload (linear in P / in T, see the header)."""
    per_perm = (t["get_means_perms"] + t["gather"] + t["score"]) / max(n_perms_sample, 1)
    idx = t["get_mat_idx"] * (n_rows / max(t["get_mat_idx_rows"], 1))
    return per_perm * n_perms + idx
