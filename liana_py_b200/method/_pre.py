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
    """What `prep_check_adata` hands on, without the matrix copies: the caller's matrix stays staged on the
    device (`Engine.stage_matrix`) until `_pipeline.prepare` commits it with the resource's column map."""
    var_names: np.ndarray         # sorted (alphabetical) names of the non-empty genes (str)
    var_cols: np.ndarray          # column of the staged matrix holding each of var_names
    n_genes_in: int               # columns of the staged matrix
    labels: np.ndarray            # int32 cluster code per KEPT cell
    categories: np.ndarray        # cluster names, code order
    obs_index: pd.Index           # cell names kept
    n_cells: int                  # kept cells
    mat_mean: float = 0.0         # mean over ALL entries of the prepared matrix (SingleCellSignalR, `_liana_pipe.py:139-140`)
    mat_max: np.float32 | None = None   # `adata.X.max()` of the prepared matrix (CellChat, `_liana_pipe.py:143-146`)
    cluster_stats: tuple | None = None  # (mean [L], std [L]) over ALL entries of a cluster's rows (scSeqComm, `:466-476`)
    n_cells_local: int | None = None    # multi-rank staging: kept cells of the row block THIS rank staged (None = all)


def choose_mtx_rep(adata, use_raw=False, layer=None, verbose=False, keep_dense=False):
    """Matrix choice (.raw.X | layers[layer] | X) -> CSR; errors as `_pre.py:227-264`.  `keep_dense`: a dense 2-D
    array is handed on as it is (the engine converts it on the device, `Engine.stage_matrix_dense`)."""
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
        if keep_dense and _is_dense(X):
            return X
        X = sparse.csr_matrix(X)
    return X


def _is_dense(X) -> bool:
    """A plain 2-D NumPy array (not np.matrix) or a 2-D CUDA torch tensor."""
    if type(X) is np.ndarray:
        return X.ndim == 2
    return type(X).__module__.startswith("torch") and hasattr(X, "data_ptr") and X.dim() == 2 and X.is_cuda


def is_mudata(obj) -> bool:
    """MuData (or a look-alike): a `.mod` mapping of modality name -> AnnData (`_liana_pipe.py:126`)."""
    return hasattr(obj, "mod") and hasattr(obj.mod, "keys")


class ModalityPair:
    """What `mdata_to_anndata` returns, reduced to what the pipeline reads: X (CSR), var_names, obs."""
    raw = None
    isbacked = False

    def __init__(self, X, var_names, obs, uns):
        self.X, self.var_names, self.obs, self.uns, self.layers = X, pd.Index(var_names), obs, uns, {}

    @property
    def shape(self):
        return self.X.shape


def _handle_mod(mdata, mod, use_raw, layer, transform, verbose):
    """One modality's matrix and names (`liana/utils/mdata_to_anndata.py:60-75`)."""
    if mod not in mdata.mod.keys():
        raise ValueError(f"`{mod}` is not in the mdata!")
    md = mdata.mod[mod]
    if use_raw:
        if getattr(md, "raw", None) is None:
            raise ValueError("`.raw` is not initialized!")
        X, names = md.raw.X, md.raw.var_names
    else:
        X, names = choose_mtx_rep(md, use_raw=False, layer=layer, verbose=verbose), md.var_names
    if transform:
        _logg(f"Transforming {mod} using {getattr(transform, '__name__', transform)}", verbose=verbose)
        X = transform(X)
    return sparse.csr_matrix(X), pd.Index(names), pd.Index(md.obs_names)


def mdata_to_anndata(mdata, x_mod, y_mod, x_layer=None, y_layer=None, x_use_raw=False, y_use_raw=False,
                     x_transform=None, y_transform=None, verbose=True) -> ModalityPair:
    """Two modalities of a MuData side by side as one matrix (`liana/utils/mdata_to_anndata.py:6-58`:
    `anndata.concat([x, y], axis=1)` = inner join on the cell names, x's order; obs / uns from the MuData)."""
    if x_mod is None or y_mod is None:
        raise ValueError("Both `x_mod` and `y_mod` must be provided!")
    xX, xn, xo = _handle_mod(mdata, x_mod, x_use_raw, x_layer, x_transform, verbose)
    yX, yn, yo = _handle_mod(mdata, y_mod, y_use_raw, y_layer, y_transform, verbose)
    if not xo.equals(yo):
        common = xo.intersection(yo, sort=False)
        xX, yX = xX[xo.get_indexer(common)], yX[yo.get_indexer(common)]
        xo = common
    if xX.shape[0] != mdata.obs.shape[0]:
        raise ValueError(f"Length of passed value for obs_names is {mdata.obs.shape[0]}, but this AnnData has shape: "
                         f"({xX.shape[0]}, {xX.shape[1] + yX.shape[1]})")
    X = sparse.hstack([xX, yX], format="csr")
    return ModalityPair(X, xn.append(yn), mdata.obs.copy(), dict(mdata.uns))


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


class CSRBlock:
    """Rows [r0, r1) of a CSR matrix as VIEWS of its arrays (what `Engine.stage_matrix` reads: indptr / indices / data /
    shape).  Not a SciPy matrix on purpose: SciPy's constructor copies a view that is much smaller than its base array
    (`_prune_array`), i.e. every rank but the first would copy its whole block on the host."""

    def __init__(self, X, r0, r1):
        e0, e1 = int(X.indptr[r0]), int(X.indptr[r1])
        self.indptr = X.indptr[r0:r1 + 1] - X.indptr[r0]
        self.indices, self.data = X.indices[e0:e1], X.data[e0:e1]
        self.shape = (r1 - r0, X.shape[1])

    def tocsr(self):
        return sparse.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)


def _canonical_copy(X):
    """A copy of a SciPy CSR matrix in canonical format (indices sorted within a row, duplicate entries summed); the
    cached format flags of the copy are dropped first, so a stale "is canonical" cannot turn `sum_duplicates` into a no-op."""
    X = X.copy()
    for flag in ("_has_canonical_format", "_has_sorted_indices"):
        X.__dict__.pop(flag, None)
    X.sum_duplicates()
    return X


def _row_block(X, dense, r0, r1):
    """Rows [r0, r1) of the caller's matrix WITHOUT copying its entries (views of the CSR arrays / of the dense rows)."""
    return X[r0:r1] if dense else CSRBlock(X, r0, r1)


def prepare_input(adata, groupby, min_cells, engine, groupby_subset=None, use_raw=False, layer=None,
                  complex_sep="_", verbose=False, want_max=False, want_cluster_stats=False, dist=None) -> PreparedInput:
    """Same checks, order and exception types as `prep_check_adata` (`_pre.py:65-169`); the passes over the
    matrix (feature sums, sample sums, finiteness, later the row / column subset) run on the GPU.

    `dist` (a DistContext with several ranks, every rank holding the same host object): each rank stages only ITS
    contiguous block of rows - 1/world of the bytes crosses each PCIe link - and the statistics are summed over ranks;
    `_pipeline.prepare` then commits the blocks and all-gathers the (small) committed matrix over NVLink."""
    if is_mudata(adata):
        raise ValueError("MuData input: pass `mdata_kwargs` (x_mod, y_mod, ...) through the method call")
    X = choose_mtx_rep(adata, use_raw=use_raw, layer=layer, verbose=verbose, keep_dense=True)
    if use_raw and layer is None:
        var_names = pd.Index(adata.raw.var_names)
    else:
        var_names = pd.Index(adata.var_names)
    dense = _is_dense(X)
    if dense:
        if type(X) is np.ndarray and X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float32)
    else:
        if X.dtype != np.float32:
            X = X.astype(np.float32)
        # duplicate entries would break "one entry per (cell, gene)".  SciPy answers `has_canonical_format` with a pass
        # over every index (0.5 ms per million entries) unless the matrix carries a cached answer; only a cached "no" is
        # acted on here - otherwise the staging pass counts the out-of-order entries on the device (`n_unordered`, below)
        if getattr(X, "_has_canonical_format", None) is False or getattr(X, "_has_sorted_indices", None) is False:
            X = _canonical_copy(X)
    var_names = _make_unique(var_names)
    obs = adata.obs

    # labels first (no matrix needed): which cells survive groupby_subset / min_cells (`_pre.py:143-162`)
    if groupby not in obs.columns:
        raise AssertionError(f"`{groupby}` not found in `adata.obs.columns`.")
    lab = obs[groupby]
    if not isinstance(lab.dtype, pd.CategoricalDtype):
        _logg(f"Converting `{groupby}` to categorical!", level="warn", verbose=verbose)
        # a local Series only: the reference converts inside the copy `prep_check_adata` made of `obs`
        # (`_pre.py:112-119, 269-271`), the caller's column keeps its dtype
        lab = lab.astype("category")
    # on the category CODES (one bincount) instead of pandas categorical ops over every cell
    codes = np.asarray(lab.cat.codes, dtype=np.int64)
    cats = lab.cat.categories
    keep = np.ones(X.shape[0], dtype=bool)
    if groupby_subset is not None:
        in_subset = np.append(np.asarray(cats.isin(groupby_subset)), False)     # code -1 (missing label) -> False
        keep &= in_subset[codes]
    counts = np.bincount(codes[keep & (codes >= 0)], minlength=len(cats))
    low_mask = (counts > 0) & (counts < min_cells)
    low = [c for c, m in zip(cats, low_mask) if m]
    if low:
        keep &= ~np.append(low_mask, False)[codes]
    rows = np.flatnonzero(keep)
    used = (counts > 0) & ~low_mask
    remap = np.cumsum(used) - 1                                                  # old code -> code among the kept categories
    kept_codes = codes[rows]
    new_codes = np.where(kept_codes >= 0, remap[np.maximum(kept_codes, 0)], -1).astype(np.int32)
    kept_categories = np.asarray(cats[used])
    if len(kept_categories) == 0:
        # the reference fails later with an IndexError on its empty category list; say what happened instead
        raise ValueError(f"No cells left: every identity in `{groupby}` was excluded (min_cells={min_cells}"
                         + (", groupby_pairs" if groupby_subset is not None else "") + ").")
    row_keep = None
    if len(rows) != X.shape[0]:
        row_keep = keep.astype(np.uint8)

    sharded = dist is not None and dist.world_size > 1

    def stage(X):
        """Stage this rank's rows [r0, r1) of X (all of them with one rank); statistics summed over the ranks."""
        r0, r1 = dist.row_block(X.shape[0], None if dense else X.indptr) if sharded else (0, X.shape[0])
        Xs = _row_block(X, dense, r0, r1) if sharded else X
        keep_s = None if row_keep is None else row_keep[r0:r1]
        stats = engine.stage_matrix_dense(Xs, keep_s) if dense else engine.stage_matrix(Xs, keep_s)
        if sharded:
            packed = dist.all_reduce_np(np.concatenate([stats["col_sums"], [stats["kept_total"], float(stats["n_nonfinite"]),
                                                                            float(stats["n_empty_rows"]), float(stats["n_unordered"])]]))
            stats = dict(col_sums=packed[:-4], kept_total=float(packed[-4]), n_nonfinite=int(packed[-3]),
                         n_empty_rows=int(packed[-2]), n_unordered=int(packed[-1]))
        return stats, r0, r1

    stats, r0, r1 = stage(X)
    if not dense and stats["n_unordered"]:
        # some row is not strictly increasing (unsorted indices or duplicate entries; the count is summed over the ranks, so
        # every rank takes this branch): SciPy's canonical form - duplicates summed, which is how the reference's SciPy
        # arithmetic treats them - staged again
        X = _canonical_copy(X)
        stats, r0, r1 = stage(X)
    # empty features are removed (sum == 0 over ALL cells), empty cells only reported (`_pre.py:122-133`)
    empty = stats["col_sums"] == 0
    if empty.any():
        _logg(f"{int(empty.sum())} features of mat are empty, they will be removed.", level="warn", verbose=verbose)
    if stats["n_empty_rows"]:
        _logg(f"{stats['n_empty_rows']} samples of mat are empty, they will be removed.", level="warn", verbose=verbose)
    if verbose:                              # sum of the first 100 stored values (`_pre.py:136-138`)
        if dense:
            first = X[:max(1, 100 // max(X.shape[1], 1) + 1) * 8].reshape(-1)
            first = np.asarray(first.cpu() if hasattr(first, "cpu") else first, dtype=np.float32)
            head = float(np.sum(first[first != 0][:100]))
        else:
            head = float(np.sum(X.data[:100]))
        if head == np.floor(head):
            _logg("Make sure that normalized counts are passed!", level="warn", verbose=verbose)
    if stats["n_nonfinite"]:
        raise ValueError("mat contains non finite values (nan or inf), please set them to 0 or remove them.")
    if low:
        _logg("The following cell identities were excluded: {0}".format(", ".join(map(str, low))),
              level="warn", verbose=verbose)
    keep_cols = np.flatnonzero(~empty)
    names = np.asarray(var_names, dtype=str)[keep_cols]
    if complex_sep is not None and verbose:
        issues = [n for n in names if complex_sep in n]
        if issues:
            _logg(f"{issues} contain `{complex_sep}`. Consider replacing those!", level="warn", verbose=True)
    order = np.argsort(names, kind="stable")            # genes in alphabetical order (`_pre.py:168`)
    n_all = len(rows) * len(keep_cols)
    mat_max = None
    if want_max:
        # scipy's sparse .max(): the implicit zeros take part (any prepared matrix that is not fully dense has one)
        stored_max, n_stored = engine.staged_max()
        if sharded:
            stored_max = np.float32(dist.all_reduce_np(np.array([float(stored_max)]), "max")[0])
            n_stored = int(dist.all_reduce_np(np.array([n_stored], dtype=np.int64))[0])
        mat_max = stored_max if (n_stored >= n_all and n_stored > 0) else np.float32(max(stored_max, np.float32(0)))
    cluster_stats = None
    if want_cluster_stats:
        # `temp.X.mean()` / `np.std(temp.X.toarray())` per label over the prepared matrix (all non-empty genes, zeros
        # included): from per-cluster sum x, sum x^2 of the staged matrix
        L = len(kept_categories)
        row_label = np.full(X.shape[0], -1, dtype=np.int32)
        row_label[rows] = new_codes
        sx, sq = engine.staged_cluster_stats(row_label[r0:r1], L)
        if sharded:
            both = dist.all_reduce_np(np.concatenate([sx, sq]))
            sx, sq = both[:L], both[L:]
        n_entries = np.bincount(new_codes[new_codes >= 0], minlength=L).astype(np.float64) * len(keep_cols)   # unlabeled cells: no cluster
        mean = sx / n_entries
        cluster_stats = (mean, np.sqrt(np.maximum(sq / n_entries - mean ** 2, 0.0)))
    return PreparedInput(mat_max=mat_max, cluster_stats=cluster_stats, var_names=names[order], var_cols=keep_cols[order].astype(np.int64), n_genes_in=X.shape[1],
                         labels=new_codes, categories=kept_categories, obs_index=obs.index[rows], n_cells=len(rows),
                         n_cells_local=int(keep[r0:r1].sum()) if sharded else None,
                         mat_mean=float(stats["kept_total"] / n_all) if n_all else 0.0)


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
