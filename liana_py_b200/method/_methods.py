"""Drop-in `li.mt.cellphonedb` / `li.mt.geometric_mean` / `li.mt.cellchat` (+ `.by_sample`) backed by the CUDA engine.

Same call signatures, defaults, result columns, dtypes and sort order as the reference's
`Method.__call__` / `MethodMeta.by_sample` (`liana/method/sc/_Method.py:92-267`).  Engine-only
options are keyword-only with defaults:

  perm_source  'philox'    (default) permutations generated on device from (seed, perm id): p-values statistically
                           equivalent to liana's, NOT the same numbers for the same seed
               'reference' the reference's own NumPy PCG64 draws -> results identical to liana
               None        = the environment variable LIANA_B200_PERM_SOURCE, else 'philox'
  order        None (auto: 'exact' with perm_source='reference', 'fast' with 'philox') | 'exact' | 'fast'
  gmean_mode   'product' (exact comparison) | 'literal' (float64 exp/log like scipy.stats.gmean)
  device       CUDA device index;  dist: a DistContext to shard permutations / samples over ranks (torchrun,
               one process per GPU)
  devices      list of CUDA device indices: ONE call, one process, several GPUs - the input is prepared on
               devices[0], the committed matrix is copied to the others over NVLink, the permutations (or the
               samples of `by_sample`) are split over them by threads; results are identical to one GPU's
`n_jobs` is accepted and ignored (parallelism is on the GPU).
"""
from __future__ import annotations

import weakref
from typing import Optional

import numpy as np
import pandas as pd

from .. import _constants as K
from . import _pipeline as PL
from . import _tables
from ._dist import DistContext

V = K.DefaultValues


class NullCounts:
    """What the engine hands a score function instead of the reference's perm_stats[2,P,T] cube:
    exceedance counts per row and the number of permutations."""

    def __init__(self, counts: np.ndarray, n_perms: int):
        self.counts, self.n_perms = counts, n_perms

    def pvals(self) -> np.ndarray:
        return self.counts / self.n_perms          # float64, `_get_mean_perms.py:137`


class MethodMeta:
    """Method metadata record; same fields as `liana/method/sc/_Method.py:16-69`."""
    instances = []

    def __init__(self, method_name, complex_cols, add_cols, fun, magnitude, magnitude_ascending,
                 specificity, specificity_ascending, permute, reference, engine_score=None):
        self.__class__.instances.append(weakref.proxy(self))
        self.method_name = method_name
        self.complex_cols = complex_cols
        self.add_cols = add_cols
        self.fun = fun
        self.magnitude = magnitude
        self.magnitude_ascending = magnitude_ascending
        self.specificity = specificity
        self.specificity_ascending = specificity_ascending
        self.permute = permute
        self.reference = reference
        self.engine_score = engine_score        # 'cpdb' | 'gmean' | None: which null the engine counts

    def describe(self):
        print(f"{self.method_name} uses `{self.magnitude}` and `{self.specificity}`"
              f" as measures of expression strength and interaction specificity, respectively")

    def get_meta(self):
        return pd.DataFrame([{"Method Name": self.method_name, "Magnitude Score": self.magnitude,
                              "Specificity Score": self.specificity, "Reference": self.reference}])

    # ------------------------------------------------------------------------------------
    def by_sample(self, adata, sample_key: str, key_added: str = K.Keys.uns_key, inplace: bool = V.inplace,
                  verbose=V.verbose, **kwargs):
        """Run the method per sample (`_Method.py:92-159`).  Samples are independent; with a
        DistContext (kwarg `dist`) contiguous blocks of samples are sharded over ranks and the numeric result
        columns are all-gathered (`_by_sample_blocks`) - no data-path collective."""
        if sample_key not in adata.obs:
            raise ValueError(f"{sample_key} was not found in `adata.obs`.")
        if not isinstance(adata.obs[sample_key].dtype, pd.CategoricalDtype):
            adata.obs[sample_key] = adata.obs[sample_key].astype("category")
        full_verbose = verbose == 'full'
        samples = list(adata.obs[sample_key].cat.categories)
        dist: Optional[DistContext] = kwargs.pop("dist", None)
        # engine-only option (ranks): gather=True - every rank returns the frame of ALL samples; gather=<rank> (an int) -
        # only that rank does, the others return None: the 28 M-row frame of 96 samples is then assembled once, with the
        # whole host to itself, instead of `world` times side by side; gather=False - every rank keeps the frame of ITS
        # block of samples (to be written out per rank)
        gather = kwargs.pop("gather", True)
        devices = kwargs.pop("devices", None)
        # category CODES, not the names: `obs[sample_key] == name` over an object array costs ~10 ns per cell and sample
        sample_col = np.asarray(adata.obs[sample_key].cat.codes)
        sizes = np.bincount(sample_col[sample_col >= 0], minlength=len(samples)).tolist()
        mine = range(len(samples)) if dist is None or dist.world_size == 1 else dist.sample_shard(sizes)
        sharded = dist is not None and dist.world_size > 1
        if devices is not None and len(devices) > 1 and sharded:
            raise ValueError("pass either `dist` (one process per GPU) or `devices` (one process, several GPUs)")
        liana_res = self._by_sample_blocks(adata, sample_key, samples, sample_col, sizes, list(mine), dist if sharded else None,
                                           gather, full_verbose, kwargs, devices)
        if inplace:
            adata.uns[key_added] = liana_res
        return None if inplace else liana_res

    def _by_sample_blocks(self, adata, sample_key, samples, sample_col, sizes, mine, dist, gather, verbose, kwargs, devices):
        """Every sample is scored into a frame of NUMBERS: the name columns are codes into tables that do not depend
        on the sample (the resource's genes / complexes, the clusters of the whole object, the sample list), so the
        per-sample results concatenate as arrays and the string columns are built ONCE at the end (in row chunks on a
        thread pool) instead of once per sample plus a pandas concat of string columns.
          ranks (`dist`)     contiguous blocks of the sample order per rank (`DistContext.sample_shard`); the numeric
                             columns are concatenated over ranks with one all-gather per column over NVLink - no
                             DataFrame is pickled
          `devices`          one process, several GPUs: contiguous blocks of samples per device, one thread each (the C
                             ABI calls - uploads, statistics, permutations - release the GIL; the pandas parts take turns)"""
        from ..resource import handle_resource
        if "groupby" not in kwargs:
            raise TypeError("by_sample: pass `groupby` as a keyword argument")
        lab = adata.obs[kwargs["groupby"]]
        cats = lab.cat.categories if isinstance(lab.dtype, pd.CategoricalDtype) else pd.Categorical(lab).categories
        res = handle_resource(interactions=kwargs.get("interactions"), resource=kwargs.get("resource"),
                              resource_name=kwargs.get("resource_name", V.resource_name), verbose=False)
        g_tables = PL.global_name_tables(res, cats)

        def score(i, **extra):
            temp = adata[sample_col == i]
            if getattr(temp, "is_view", True):      # anndata hands out views; a stand-in that already copied is not copied again
                temp = temp.copy()
            cols, state = self.__call__(temp, inplace=False, verbose=verbose, _names=False, **{**kwargs, **extra})
            cols = PL.to_global_codes(cols, PL.global_code_maps(state, g_tables))
            n_rows = len(next(iter(cols.values()))) if cols else 0
            return {sample_key: np.full(n_rows, i, dtype=PL.code_dtype(len(samples))), **cols}

        import os
        import time
        tlog = [] if os.environ.get("LRPERM_DEBUG_TIMING") else None

        def mark(what, t0):
            if tlog is not None:
                tlog.append(f"{what} {1e3 * (time.perf_counter() - t0):.1f} ms")

        t_all = time.perf_counter()
        if devices is not None and len(devices) > 1:
            from concurrent.futures import ThreadPoolExecutor
            blocks = [[i for i in DistContext(k, len(devices)).sample_shard(sizes) if i in set(mine)] for k in range(len(devices))]
            with ThreadPoolExecutor(len(devices)) as pool:
                parts = [p for part in pool.map(lambda k: [score(i, device=int(devices[k])) for i in blocks[k]], range(len(devices)))
                         for p in part]
        else:
            parts = [score(i) for i in mine]
        mark(f"score {len(parts)} samples", t_all)
        # column names / dtypes of a rank without samples come from the others (tiny object exchange)
        t0 = time.perf_counter()
        collect = dist is not None and gather is not False
        spec = [(c, parts[0][c].dtype.str) for c in parts[0]] if parts else None
        if collect:                                 # a collective: every rank takes part
            spec = next((s for s in dist.all_gather_objects(spec) if s is not None), None)
        if spec is None:                            # no sample on this rank (gather=False) or anywhere
            return pd.DataFrame({sample_key: pd.Series([], dtype=object)})
        local = {c: (np.concatenate([p[c] for p in parts]) if parts else np.zeros(0, dtype=np.dtype(dt))) for c, dt in spec}
        mark("concatenate", t0)
        if collect:
            t0 = time.perf_counter()
            local = dist.all_gather_rows(local, dst=None if gather is True else int(gather))
            mark("all_gather_rows", t0)
            if local is None:                       # gather=<another rank>
                return None
        g_tables = dict(g_tables)
        g_tables[sample_key] = pd.Index(samples)
        t0 = time.perf_counter()
        out = PL.frame_from_codes(local, g_tables)
        mark("frame_from_codes", t0)
        if tlog is not None:
            import sys
            print(f"[by_sample rank {dist.rank if dist is not None else 0}] " + " | ".join(tlog), file=sys.stderr, flush=True)
        return out


class Method(MethodMeta):
    def __init__(self, _method: MethodMeta):
        super().__init__(method_name=_method.method_name, complex_cols=_method.complex_cols,
                         add_cols=_method.add_cols, fun=_method.fun, magnitude=_method.magnitude,
                         magnitude_ascending=_method.magnitude_ascending, specificity=_method.specificity,
                         specificity_ascending=_method.specificity_ascending, permute=_method.permute,
                         reference=_method.reference, engine_score=_method.engine_score)
        self._method = _method

    def __call__(self, adata, groupby: str, resource_name: str = V.resource_name, expr_prop: float = V.expr_prop,
                 min_cells: int = V.min_cells, groupby_pairs: Optional[pd.DataFrame] = V.groupby_pairs,
                 base: float = V.logbase, supp_columns: Optional[list] = V.supp_columns,
                 return_all_lrs: bool = V.return_all_lrs, key_added: str = K.Keys.uns_key,
                 use_raw: Optional[bool] = V.use_raw, layer: Optional[str] = V.layer, de_method: str = V.de_method,
                 n_perms: Optional[int] = V.n_perms, seed: int = V.seed, n_jobs: int = 1,
                 resource: Optional[pd.DataFrame] = V.resource, interactions: Optional[list] = V.interactions,
                 mdata_kwargs: dict = dict(), inplace: bool = V.inplace, verbose: Optional[bool] = V.verbose,
                 *, perm_source: Optional[str] = None, order: Optional[str] = None, gmean_mode: str = "product",
                 device: int = 0, dist: Optional[DistContext] = None, devices: Optional[list] = None,
                 _names: bool = True):
        device, devices = PL.resolve_devices(device, devices)
        # supplementary statistic columns ride along with the method's own (`_liana_pipe.py:111-114`)
        add_cols = PL.method_columns(self.add_cols, supp_columns)
        state = PL.prepare(adata, groupby, resource_name, resource, interactions, groupby_pairs, min_cells,
                           use_raw, layer, verbose, device=device, base=base, want_max=K.MAT_MAX in add_cols,
                           want_cluster_stats=("ligand_cdf" in add_cols) or ("receptor_cdf" in add_cols),
                           mdata_kwargs=mdata_kwargs, dist=dist)
        orderby, ascending = (self.magnitude, self.magnitude_ascending) if self.magnitude is not None \
            else (self.specificity, self.specificity_ascending)             # `_liana_pipe.py:246-250`
        if (_names and self.permute and n_perms is not None and self.magnitude is not None and not return_all_lrs
                and (dist is None or dist.world_size == 1) and devices is None):
            liana_res = _call_overlapped(state, self._method, expr_prop, n_perms, seed, perm_source, order, gmean_mode, add_cols)
            if liana_res is not None:
                if inplace:
                    adata.uns[key_added] = liana_res
                return None if inplace else liana_res
        liana_res = run_method(state, self._method, expr_prop, n_perms, seed, return_all_lrs,
                               perm_source, order, gmean_mode, dist, add_cols=add_cols, devices=devices)
        if not _names:                          # by_sample: (sorted numeric columns, prepared state), see `_by_sample_blocks`
            return PL.sorted_code_columns(state.engine, liana_res, orderby, ascending), state
        liana_res = PL.finish_frame(state, liana_res, orderby, ascending)
        if inplace:
            adata.uns[key_added] = liana_res
        return None if inplace else liana_res


_NULL_POOL = None


def _call_overlapped(state, method, expr_prop, n_perms, seed, perm_source, order, gmean_mode, add_cols):
    """One permutation method, result frame assembled WHILE the permutations run.  The frame is sorted by the magnitude
    (`_liana_pipe.py:246-250`), which does not depend on the null: the row order is taken first (device argsort), the
    null is started on a worker thread (the C ABI call releases the GIL; nothing else touches the engine until it
    returns), the main thread gathers every column in that order and builds the name columns, and the p-values join
    at the end.  Same frame as `run_method` + `finish_frame`; None when the shortcut does not apply (few rows, NaN keys)."""
    global _NULL_POOL
    stat = "trimean" if K.LIGAND_TRIMEAN in method.complex_cols else "means"
    tri, lr_res = state.resolved(expr_prop, False, stat)
    if add_cols:
        lr_res = PL.add_method_columns(state, tri, lr_res, add_cols)
        lr_res = lr_res[sorted(lr_res.columns)]
    magnitude = method.fun(lr_res, None)[0]
    vals = np.asarray(magnitude, dtype=np.float64)
    if state.engine is None or len(vals) < 4096 or np.isnan(vals).any():
        return None
    row_order = state.engine.argsort(vals, descending=not method.magnitude_ascending)
    if _NULL_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _NULL_POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="lrperm-null")
    fut = _NULL_POOL.submit(_compute_null, state, method, tri, tri.keep, expr_prop, n_perms, seed, False, perm_source, order,
                            gmean_mode, None)
    try:
        cols = PL.gather_columns(state, lr_res, row_order)
        cols[method.magnitude] = np.asarray(magnitude).take(row_order)
    finally:
        null = fut.result()                      # the engine is free again (also when the gather failed)
    cols[method.specificity] = null.pvals().take(row_order)
    return pd.DataFrame(cols, index=pd.Index(row_order), copy=False)


def run_method(state, method: MethodMeta, expr_prop, n_perms, seed, return_all_lrs, perm_source, order,
               gmean_mode, dist, aggregate_flag=False, null_cache=None, permuting_methods=None,
               add_cols=None, devices=None, reference_quirks=False) -> pd.DataFrame:
    """`_run_method` (`_liana_pipe.py:363-452`) for one method on a prepared PipelineState.

    In a consensus (`null_cache` given) every permutation method named in `permuting_methods` is counted
    from ONE pass over the permutations, and the counts are cached for the methods that follow."""
    stat = "trimean" if K.LIGAND_TRIMEAN in method.complex_cols else "means"
    tri, lr_res = state.resolved(expr_prop, return_all_lrs, stat)       # shared by the methods of a consensus
    add_cols = method.add_cols if add_cols is None else add_cols
    if add_cols:
        lr_res = PL.add_method_columns(state, tri, lr_res, add_cols)
        lr_res = lr_res[sorted(lr_res.columns)]            # `np.union1d` of the column sets (`_liana_pipe.py:385`)
    kept = tri.keep
    x = lr_res[kept] if return_all_lrs else lr_res
    null = _compute_null(state, method, tri, kept, expr_prop, n_perms, seed, return_all_lrs, perm_source, order, gmean_mode,
                         dist, aggregate_flag, null_cache, permuting_methods, devices, reference_quirks)
    magnitude, specificity = method.fun(x, null) if method.permute else method.fun(x)
    spec_name = method.specificity if (not method.permute or n_perms is not None) else None
    # the cached base frame stays untouched: a shallow copy shares its (never written) columns, the score columns
    # added below are new arrays owned by the copy
    lr_res = (lr_res[K.PRIMARY] if aggregate_flag else lr_res).copy(deep=False)
    if return_all_lrs:
        lr_res[K.LRS_TO_KEEP] = kept
    if method.magnitude is not None:
        lr_res[method.magnitude] = _scatter(magnitude, kept, return_all_lrs)
    if spec_name is not None:
        lr_res[spec_name] = _scatter(specificity, kept, return_all_lrs)
    if return_all_lrs:
        # rows that failed expr_prop get the worst kept value (`_liana_pipe.py:433-443`)
        for col, asc in ((method.magnitude, method.magnitude_ascending), (spec_name, method.specificity_ascending)):
            if col is not None:
                vals = lr_res[col]
                lr_res.loc[~kept, col] = vals.max() if asc else vals.min()
    if aggregate_flag:
        cols = K.PRIMARY + [c for c in (method.magnitude, spec_name) if c is not None]
        lr_res = lr_res[cols]
    return lr_res


def _compute_null(state, method, tri, kept, expr_prop, n_perms, seed, return_all_lrs, perm_source, order, gmean_mode,
                  dist, aggregate_flag=False, null_cache=None, permuting_methods=None, devices=None, reference_quirks=False):
    """The permutation null of one method (`_liana_pipe.py:402-419`): exceedance counts from the engine, or None for a
    method without permutations / `n_perms=None`."""
    null = None
    if method.permute and n_perms is not None:
        sub = tri if not return_all_lrs else _subset(tri, kept)
        key = (method.engine_score, n_perms, seed, perm_source, order, gmean_mode, expr_prop, return_all_lrs)
        if null_cache is not None and key in null_cache:
            null = null_cache[key]
        elif method.engine_score == "cellchat":
            obs = PL.observed_cellchat(sub.ligand_means, sub.receptor_means, cellchat._kh)
            out = PL.permutation_counts_trimean(state, sub, obs, n_perms, seed, perm_source, cellchat._kh, dist, devices)
            null = NullCounts(out["cellchat"], n_perms)
            if null_cache is not None:
                null_cache[key] = null
            if reference_quirks and aggregate_flag:
                # the reference's in-place `adata.X /= norm_factor` stays in effect for the methods that follow
                state.scale_matrix_like_reference(state.prep.mat_max)
        else:
            wanted = {method.engine_score}
            # (one pass for all mean-based nulls of a consensus - unless the reference's matrix carry-over is being
            # reproduced: then every method must see the matrix as it is when ITS turn comes)
            if null_cache is not None and permuting_methods and not reference_quirks:
                wanted |= {m.engine_score for m in permuting_methods if m.permute and m.engine_score in ("cpdb", "gmean")}
            oc = PL.observed_cpdb(sub.ligand_means, sub.receptor_means) if "cpdb" in wanted else None
            og = PL.observed_gmean(sub.ligand_means, sub.receptor_means) if "gmean" in wanted else None
            out = PL.permutation_counts(state, sub, oc, og, n_perms, seed, perm_source, order, gmean_mode, dist, devices)
            if null_cache is not None:
                for name in ("cpdb", "gmean"):
                    if name in out:
                        null_cache[(name,) + key[1:]] = NullCounts(out[name], n_perms)
            null = NullCounts(out[method.engine_score], n_perms)
    return null


def _scatter(values, kept, return_all_lrs):
    if not return_all_lrs:
        return values
    out = np.full(kept.shape[0], np.nan, dtype=np.asarray(values).dtype if np.asarray(values).dtype.kind == 'f' else np.float64)
    out[kept] = values
    return out


def _subset(tri, mask):
    return _tables.ResolvedTriples(**{k: getattr(tri, k)[mask] for k in tri.__dataclass_fields__})


# ---- the two permutation score functions ------------------------------------------------------
def _cpdb_score(x: pd.DataFrame, perm_stats: Optional[NullCounts]):
    """CellPhoneDB-like lr_means and p-values (`liana/method/sc/_cellphonedb.py:9-30`)."""
    lr_means = PL.observed_cpdb(x[K.LIGAND_MEANS].values, x[K.RECEPTOR_MEANS].values)
    return lr_means, (perm_stats.pvals() if perm_stats is not None else None)


def _gmean_score(x: pd.DataFrame, perm_stats: Optional[NullCounts]):
    """Geometric-mean magnitude and p-values (`liana/method/sc/_geometric_mean.py:6-25`)."""
    lr_gmeans = PL.observed_gmean(x[K.LIGAND_MEANS].values, x[K.RECEPTOR_MEANS].values)
    return lr_gmeans, (perm_stats.pvals() if perm_stats is not None else None)


def _cellchat_score(x: pd.DataFrame, perm_stats: Optional[NullCounts]):
    """CellChat-like lr_probs and p-values (`liana/method/sc/_cellchat.py:13-32`)."""
    lr_prob = PL.observed_cellchat(x[K.LIGAND_TRIMEAN].values, x[K.RECEPTOR_TRIMEAN].values, cellchat._kh)
    return lr_prob, (perm_stats.pvals() if perm_stats is not None else None)


_cellphonedb = MethodMeta(method_name="CellPhoneDB", complex_cols=[K.LIGAND_MEANS, K.RECEPTOR_MEANS], add_cols=[],
                          fun=_cpdb_score, magnitude="lr_means", magnitude_ascending=False,
                          specificity="cellphone_pvals", specificity_ascending=True, permute=True,
                          reference="Efremova et al., 2020. Nature protocols 15(4).", engine_score="cpdb")
_geometric_mean = MethodMeta(method_name="Geometric Mean", complex_cols=[K.LIGAND_MEANS, K.RECEPTOR_MEANS], add_cols=[],
                             fun=_gmean_score, magnitude="lr_gmeans", magnitude_ascending=False,
                             specificity="gmean_pvals", specificity_ascending=True, permute=True,
                             reference="CellPhoneDBv2's permutation approach applied to the geometric means.",
                             engine_score="gmean")

_cellchat = MethodMeta(method_name="CellChat", complex_cols=[K.LIGAND_TRIMEAN, K.RECEPTOR_TRIMEAN], add_cols=[K.MAT_MAX],
                       fun=_cellchat_score, magnitude="lr_probs", magnitude_ascending=False,
                       specificity="cellchat_pvals", specificity_ascending=True, permute=True,
                       reference="Jin et al., 2021. Nature communications 12(1).", engine_score="cellchat")

cellphonedb = Method(_method=_cellphonedb)
geometric_mean = Method(_method=_geometric_mean)
cellchat = Method(_method=_cellchat)
cellchat._kh = 0.5               # `_cellchat.py:53`: a module-level knob in the reference too
