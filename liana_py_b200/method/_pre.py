"""Input contract of the hot path: what the reference's `prep_check_adata`
(`liana/method/_pipe_utils/_pre.py:65-169`) and the gene subset in `liana_pipe`
(`liana/method/sc/_liana_pipe.py:149-164`) hand to the permutation code.

Instead of building intermediate AnnData copies, the matrix, gene names and labels are pulled
out once into a `PreparedInput` (CSR float32, names sorted alphabetically, cluster codes in
category order with unused categories dropped).  Same checks, same exception types.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass

import numpy as np
import pandas as pd
from scipy import sparse

_log = logging.getLogger("liana_py_b200")


def _logg(msg, level="info", verbose=False):
    if verbose:
        (_log.warning if level == "warn" else _log.info)(msg)


@dataclass
class PreparedInput:
    X: sparse.csr_matrix          # cells x genes, float32, canonical CSR
    var_names: np.ndarray         # sorted gene names (str)
    labels: np.ndarray            # int32 cluster code per cell
    categories: np.ndarray        # cluster names, code order
    obs_index: pd.Index           # cell names kept
    mat_mean: float = 0.0         # mean over ALL entries of the prepared matrix (SingleCellSignalR, `_liana_pipe.py:139-140`)


def choose_mtx_rep(adata, use_raw=False, layer=None, verbose=False):
    """Matrix choice (.raw.X | layers[layer] | X) -> CSR; errors as `_pre.py:227-264`."""
    is_layer = layer is not None
    if is_layer and use_raw:
        raise ValueError("Cannot specify `layer` and have `use_raw=True`.")
    if is_layer:
        _logg(f"Using the `{layer}` layer!", verbose=verbose)
        X = adata.layers[layer]
    elif use_raw:
        if getattr(adata, "raw", None) is None:
            raise ValueError("`.raw` is not initialized!")
        _logg("Using `.raw`!", verbose=verbose)
        X = adata.raw.X
    else:
        _logg("Using `.X`!", verbose=verbose)
        X = adata.X
    if not sparse.isspmatrix_csr(X):
        _logg("Converting to sparse csr matrix!", verbose=verbose)
        X = sparse.csr_matrix(X)
    return X


def _make_unique(names: pd.Index) -> pd.Index:
    names = pd.Index(names.astype(str))
    if names.is_unique:
        return names
    seen, counts, out = set(names), {}, []
    dup = names.duplicated(keep="first")
    for n, d in zip(names, dup):
        if not d:
            out.append(n)
            continue
        k = counts.get(n, 0) + 1
        while f"{n}-{k}" in seen:
            k += 1
        counts[n] = k
        seen.add(f"{n}-{k}")
        out.append(f"{n}-{k}")
    return pd.Index(out)


def prepare_input(adata, groupby, min_cells, groupby_subset=None, use_raw=False, layer=None,
                  complex_sep="_", verbose=False) -> PreparedInput:
    if hasattr(adata, "mod"):
        raise NotImplementedError("MuData input is outside the scope of the B200 permutation engine; "
                                  "convert to AnnData first.")
    X = choose_mtx_rep(adata, use_raw=use_raw, layer=layer, verbose=verbose)
    if use_raw and layer is None:
        var_names = pd.Index(adata.raw.var_names)
    else:
        var_names = pd.Index(adata.var_names)
    X = X.astype(np.float32)
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    var_names = _make_unique(var_names)
    obs = adata.obs

    # empty features are removed (sum == 0), empty cells only reported (`_pre.py:122-133`)
    col_sums = np.asarray(X.sum(axis=0)).ravel()
    empty = col_sums == 0
    if empty.any():
        _logg(f"{int(empty.sum())} features of mat are empty, they will be removed.", level="warn", verbose=verbose)
        keep = np.flatnonzero(~empty)
        X = X[:, keep]
        var_names = var_names[keep]
    n_empty_cells = int((np.asarray(X.sum(axis=1)).ravel() == 0).sum())
    if n_empty_cells:
        _logg(f"{n_empty_cells} samples of mat are empty, they will be removed.", level="warn", verbose=verbose)
    head = float(np.sum(X.data[:100]))
    if head == np.floor(head):
        _logg("Make sure that normalized counts are passed!", level="warn", verbose=verbose)
    if np.any(~np.isfinite(X.data)):
        raise ValueError("mat contains non finite values (nan or inf), please set them to 0 or remove them.")

    if groupby not in obs.columns:
        raise AssertionError(f"`{groupby}` not found in `adata.obs.columns`.")
    lab = obs[groupby]
    if not isinstance(lab.dtype, pd.CategoricalDtype):
        _logg(f"Converting `{groupby}` to categorical!", level="warn", verbose=verbose)
        lab = lab.astype("category")
        obs[groupby] = lab       # the reference converts in place too (`_pre.py:269-271`)
    rows = np.arange(X.shape[0])
    if groupby_subset is not None:
        rows = rows[np.asarray(lab.isin(groupby_subset))]
    lab_sub = lab.iloc[rows].cat.remove_unused_categories()
    counts = lab_sub.value_counts(sort=False)
    low = [c for c in lab_sub.cat.categories if counts[c] < min_cells]
    if low:
        keep_mask = ~np.asarray(lab_sub.isin(low))
        rows = rows[keep_mask]
        lab_sub = lab_sub[keep_mask].cat.remove_unused_categories()
        _logg("The following cell identities were excluded: {0}".format(", ".join(map(str, low))),
              level="warn", verbose=verbose)
    if len(rows) != X.shape[0]:
        X = X[rows]
    if complex_sep is not None and verbose:
        issues = [n for n in var_names if complex_sep in n]
        if issues:
            _logg(f"{issues} contain `{complex_sep}`. Consider replacing those!", level="warn", verbose=True)
    # genes in alphabetical order (`_pre.py:168`)
    names = np.asarray(var_names, dtype=str)
    order = np.argsort(names, kind="stable")
    if not np.array_equal(order, np.arange(len(order))):
        X = X[:, order]
        names = names[order]
    X = sparse.csr_matrix(X)
    X.sort_indices()
    n_all = X.shape[0] * X.shape[1]
    mat_mean = float(X.data.sum(dtype=np.float64) / n_all) if n_all else 0.0
    return PreparedInput(X=X, var_names=names, labels=np.asarray(lab_sub.cat.codes, dtype=np.int32),
                         categories=np.asarray(lab_sub.cat.categories), obs_index=obs.index[rows], mat_mean=mat_mean)


def assert_covered(subset, superset, prop_missing_allowed=0.98, verbose=False):
    """Raise ValueError if too few resource entities are in the data (`_pre.py:15-62`)."""
    subset = np.asarray(subset)
    is_missing = ~np.isin(subset, superset)
    if subset.size == 0:
        prop_missing, x_missing = 1.0, "values in interactions argument"
    else:
        prop_missing = float(np.sum(is_missing)) / len(subset)
        x_missing = ", ".join(map(str, subset[is_missing]))
    if prop_missing > prop_missing_allowed:
        raise ValueError(
            "Please check if appropriate organism/ID type was provided! "
            f"Allowed proportion ({prop_missing_allowed}) of missing resource elements exceeded "
            f"({prop_missing:.2f}). Too few features from the resource were found in the data."
            f" [{x_missing}] missing from var_names")
    _logg(f"{prop_missing:.2f} of entities in the resource are missing from the data.",
          verbose=verbose and prop_missing > 0)
