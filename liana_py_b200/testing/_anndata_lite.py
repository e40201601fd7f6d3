"""A small AnnData look-alike for environments without `anndata`.

The engine's drop-in methods only need the handful of attributes the reference
pipeline touches: ``X`` (SciPy CSR), ``obs``/``var``
(pandas), ``var_names``/``obs_names``, ``raw``, ``layers``, ``uns``, ``obsp``,
``obsm``, ``shape``, ``isbacked``, ``copy()`` and ``__getitem__`` with row /
column selectors.  Row subsetting drops unused categories from categorical obs
columns, which is the anndata behaviour the reference relies on
(`liana/method/_pipe_utils/_get_mean_perms.py:47`, `_liana_pipe.py:256`).

This is test / demo infrastructure; with real ``anndata`` installed the drop-in
methods accept ``anndata.AnnData`` directly.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy import sparse


class _Raw:
    def __init__(self, X, var):
        self.X = X
        self.var = var

    @property
    def var_names(self):
        return self.var.index


class AnnDataLite:
    def __init__(self, X=None, obs=None, var=None, dtype=None, obsp=None, uns=None,
                 obsm=None, layers=None, raw=None):
        if X is not None and not sparse.issparse(X):
            X = np.asarray(X)
        if dtype is not None and X is not None:
            X = X.astype(dtype)
        self.X = X
        n_obs, n_var = X.shape
        if obs is None:
            obs = pd.DataFrame(index=pd.RangeIndex(n_obs).astype(str))
        if var is None:
            var = pd.DataFrame(index=pd.RangeIndex(n_var).astype(str))
        assert obs.shape[0] == n_obs, "obs rows must match X rows"
        assert var.shape[0] == n_var, "var rows must match X columns"
        self.obs = obs
        self.var = var
        self.obsp = {} if obsp is None else obsp
        self.uns = {} if uns is None else uns
        self.obsm = {} if obsm is None else obsm
        self.layers = {} if layers is None else layers
        self.raw = raw
        self.isbacked = False
        self.is_view = False                 # subsetting copies (anndata hands out views)

    # -- names / shape -----------------------------------------------------------------
    @property
    def var_names(self):
        return self.var.index

    @var_names.setter
    def var_names(self, names):
        self.var.index = pd.Index(names)

    @property
    def obs_names(self):
        return self.obs.index

    @obs_names.setter
    def obs_names(self, names):
        self.obs.index = pd.Index(names)

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    def uns_keys(self):
        return list(self.uns.keys())

    def var_names_make_unique(self):
        names = pd.Series(self.var.index.astype(str))
        dup = names.duplicated(keep='first')
        if dup.any():
            counts = {}
            out = []
            seen = set(names)
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
            self.var.index = pd.Index(out)

    def set_raw(self):
        """Freeze the current X/var as ``.raw`` (like ``adata.raw = adata``)."""
        self.raw = _Raw(self.X.copy(), self.var.copy())

    # -- copies / subsetting ------------------------------------------------------------
    def copy(self):
        return AnnDataLite(X=self.X.copy(), obs=self.obs.copy(), var=self.var.copy(),
                           obsp=dict(self.obsp), uns=dict(self.uns), obsm=dict(self.obsm),
                           layers={k: v.copy() for k, v in self.layers.items()},
                           raw=self.raw)

    def to_memory(self):
        return self

    @staticmethod
    def _row_selector(sel, n):
        if isinstance(sel, slice):
            return np.arange(n)[sel]
        if isinstance(sel, (pd.Series, pd.DataFrame)):
            sel = sel.values
        sel = np.asarray(sel)
        if sel.ndim == 2:
            sel = sel.reshape(-1)
        if sel.dtype == bool:
            return np.flatnonzero(sel)
        return sel.astype(np.int64)

    def _col_selector(self, sel):
        n = self.X.shape[1]
        if isinstance(sel, slice):
            return np.arange(n)[sel]
        if isinstance(sel, (pd.Series, pd.Index)):
            sel = sel.values
        sel = np.asarray(sel)
        if sel.dtype == bool:
            return np.flatnonzero(sel)
        if sel.dtype.kind in 'iu':
            return sel.astype(np.int64)
        pos = self.var.index.get_indexer(sel)
        if (pos < 0).any():
            raise KeyError("unknown variable names in column selector")
        return pos

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rsel, csel = key
        else:
            rsel, csel = key, slice(None)
        rows = self._row_selector(rsel, self.X.shape[0])
        cols = self._col_selector(csel)
        all_rows = isinstance(rsel, slice) and rsel == slice(None)
        all_cols = isinstance(csel, slice) and csel == slice(None)
        X = self.X
        if not all_rows:
            X = X[rows]
        if not all_cols:
            X = X[:, cols]
        obs = self.obs if all_rows else self.obs.iloc[rows].copy()
        if not all_rows:
            for c in obs.columns:
                if isinstance(obs[c].dtype, pd.CategoricalDtype):
                    obs[c] = obs[c].cat.remove_unused_categories()
        var = self.var if all_cols else self.var.iloc[cols].copy()
        layers = {}
        for k, v in self.layers.items():
            lv = v if all_rows else v[rows]
            layers[k] = lv if all_cols else lv[:, cols]
        out = AnnDataLite(X=X, obs=obs.copy() if all_rows else obs, var=var.copy() if all_cols else var,
                          obsp={}, uns=dict(self.uns), obsm=dict(self.obsm) if all_rows else {},
                          layers=layers, raw=self.raw if all_rows else None)
        return out

    def __repr__(self):
        return f"AnnDataLite with n_obs x n_vars = {self.shape[0]} x {self.shape[1]}"


class MuDataLite:
    """A MuData look-alike: `.mod` (name -> AnnData-like) plus the shared obs / obsp / obsm / uns the
    reference's `mdata_to_anndata` copies (`liana/utils/mdata_to_anndata.py:53-56`)."""

    def __init__(self, mod: dict, obs=None):
        self.mod = dict(mod)
        first = next(iter(self.mod.values()))
        self.obs = first.obs.copy() if obs is None else obs
        self.obsp, self.obsm, self.uns = {}, {}, {}

    @property
    def shape(self):
        return (self.obs.shape[0], sum(m.shape[1] for m in self.mod.values()))

    def __getitem__(self, key):
        """Row subsetting (what `by_sample` does with a MuData): every modality and the shared obs."""
        rows = AnnDataLite._row_selector(key, self.obs.shape[0])
        obs = self.obs.iloc[rows].copy()
        for c in obs.columns:
            if isinstance(obs[c].dtype, pd.CategoricalDtype):
                obs[c] = obs[c].cat.remove_unused_categories()
        out = MuDataLite({k: m[rows] for k, m in self.mod.items()}, obs=obs)
        out.uns = dict(self.uns)
        return out

    def copy(self):
        out = MuDataLite({k: m.copy() for k, m in self.mod.items()}, obs=self.obs.copy())
        out.obsp, out.obsm, out.uns = dict(self.obsp), dict(self.obsm), dict(self.uns)
        return out

    def __repr__(self):
        return f"MuDataLite with {len(self.mod)} modalities, n_obs = {self.obs.shape[0]}"
