"""`li.mt.rank_aggregate` / `AggregateClass` backed by the CUDA engine.

Same call signature, defaults, result columns and sort as the reference
(`liana/method/sc/_rank_aggregate.py:14-164`, `liana/method/__init__.py:13-14`); engine-only
options are keyword-only (see `_methods.Method.__call__`).  One prepared PipelineState is shared by
all methods (the reference's `_get_lr` also runs once, `_liana_pipe.py:171-180`); every permutation
method of the consensus is counted from ONE pass over the permutations (the reference regenerates
the same cube per method, `_liana_pipe.py:199-217`).
"""
from __future__ import annotations

from typing import Optional

import pandas as pd

from .. import _constants as K
from . import _pipeline as PL
from ._aggregate import aggregate
from ._dist import DistContext
from ._methods import MethodMeta, cellphonedb, run_method
from ._scores import connectome, logfc, natmi, singlecellsignalr

V = K.DefaultValues


class AggregateClass(MethodMeta):
    """LIANA's method consensus class (`_rank_aggregate.py:14-54`)."""

    def __init__(self, _SCORE, methods):
        super().__init__(method_name=_SCORE.method_name, complex_cols=[], add_cols=[], fun=_SCORE.fun,
                         magnitude=_SCORE.magnitude, magnitude_ascending=True, specificity=_SCORE.specificity,
                         specificity_ascending=True, permute=_SCORE.permute, reference=_SCORE.reference)
        self._SCORE = _SCORE
        self.methods = methods
        self.specificity_specs = {m.method_name: (m.specificity, m.specificity_ascending)
                                  for m in methods if m.specificity is not None}
        self.magnitude_specs = {m.method_name: (m.magnitude, m.magnitude_ascending)
                                for m in methods if m.magnitude is not None}
        self.add_cols = list({x for li in [m.add_cols for m in methods] for x in li})
        self.complex_cols = list({x for li in [m.complex_cols for m in methods] for x in li})

    def describe(self):
        print(f"{self.method_name} returns `{self.magnitude}`, `{self.specificity}`. "
              f"{self.magnitude} and {self.specificity} respectively represent an aggregate of the "
              f"`magnitude`- and `specificity`-related scoring functions from the different methods.")

    def __call__(self, adata, groupby: str, resource_name: str = V.resource_name, expr_prop: float = V.expr_prop,
                 min_cells: int = V.min_cells, groupby_pairs: Optional[pd.DataFrame] = V.groupby_pairs,
                 base: float = V.logbase, aggregate_method: str = "rra", consensus_opts: Optional[list] = None,
                 return_all_lrs: bool = V.return_all_lrs, key_added: str = K.Keys.uns_key,
                 use_raw: Optional[bool] = V.use_raw, layer: Optional[str] = V.layer, de_method: str = V.de_method,
                 n_perms: Optional[int] = V.n_perms, seed: int = V.seed, n_jobs: int = 1,
                 resource: Optional[pd.DataFrame] = V.resource, interactions: Optional[list] = V.interactions,
                 mdata_kwargs: dict = dict(), inplace: bool = V.inplace, verbose: Optional[bool] = V.verbose,
                 *, perm_source: Optional[str] = None, order: Optional[str] = None, gmean_mode: str = "product",
                 device: int = 0, dist: Optional[DistContext] = None, devices: Optional[list] = None,
                 reference_quirks: bool = False, _names: bool = True):
        """`reference_quirks=True` (engine-only, validation): reproduce the one known side effect of the reference's
        consensus - CellChat's null scales the shared matrix in place (`_get_mean_perms.py:43-44`), so the permutation
        methods that run AFTER CellChat in `methods` shuffle `X / mat_max` against unscaled observed scores.  Default
        False: every method's null is computed on the matrix the caller passed (what the reference does for any
        consensus without CellChat, e.g. the default `rank_aggregate`)."""
        device, devices = PL.resolve_devices(device, devices)
        if n_perms is None:
            consensus_opts = "Magnitude"                       # `_liana_pipe.py:108-109`
        state = PL.prepare(adata, groupby, resource_name, resource, interactions, groupby_pairs, min_cells,
                           use_raw, layer, verbose, device=device, base=base, want_max=K.MAT_MAX in self.add_cols,
                           want_cluster_stats="ligand_cdf" in self.add_cols, mdata_kwargs=mdata_kwargs, dist=dist)
        null_cache = {}
        lrs = {}
        for method in self.methods:
            if verbose:
                print(f"Running {method.method_name}")
            meta = getattr(method, "_method", method)
            lrs[method.method_name] = run_method(state, meta, expr_prop, n_perms, seed, return_all_lrs, perm_source,
                                                 order, gmean_mode, dist, aggregate_flag=True, null_cache=null_cache,
                                                 permuting_methods=[getattr(m, "_method", m) for m in self.methods],
                                                 devices=devices, reference_quirks=reference_quirks)
        if consensus_opts is False:                             # per-method results as they are (`:220-221`)
            return {k: PL.materialise_names(state, v) for k, v in lrs.items()}
        # the reference outer-merges the per-method frames on the key columns (`_aggregate.py:39-49`);
        # they all hold the same rows in the same order here, so the join is a column concatenation
        frames = list(lrs.values())
        lr_res = frames[0].reset_index(drop=True)
        for f in frames[1:]:
            for col in f.columns:
                if col not in lr_res.columns:                   # duplicated score names keep the first (`:47-49`)
                    lr_res[col] = f[col].values
        # `_aggregate` sorts by its last consensus column and `liana_pipe` by the magnitude (`_aggregate.py:65`,
        # `_liana_pipe.py:246-250`); sorting a sorted frame again changes nothing, so the frame is ordered once
        lr_res = aggregate(lr_res, consensus=self, aggregate_method=aggregate_method, consensus_opts=consensus_opts,
                           engine=state.engine, sort=False)
        orderby = self.magnitude if self.magnitude in lr_res.columns else self.specificity
        if not _names:                          # by_sample: (sorted numeric columns, prepared state)
            if orderby not in lr_res.columns:
                return {c: lr_res[c].values for c in lr_res.columns}, state
            return PL.sorted_code_columns(state.engine, lr_res, orderby, True), state
        if orderby not in lr_res.columns:       # no consensus column was asked for: rows stay as they are
            lr_res = PL.materialise_names(state, lr_res)
        else:
            lr_res = PL.finish_frame(state, lr_res, orderby, True)
        if inplace:
            adata.uns[key_added] = lr_res
        return None if inplace else lr_res


_rank_aggregate_meta = MethodMeta(method_name="Rank_Aggregate", complex_cols=[], add_cols=[], fun=None,
                                  magnitude="magnitude_rank", magnitude_ascending=True,
                                  specificity="specificity_rank", specificity_ascending=True, permute=False,
                                  reference="Dimitrov et al., 2022. Nature Communications 13(1).")
aggregate_meta = _rank_aggregate_meta

# callable consensus instance (`liana/method/__init__.py:13-14`)
_methods = [cellphonedb, connectome, logfc, natmi, singlecellsignalr]
rank_aggregate = AggregateClass(_rank_aggregate_meta, methods=_methods)
