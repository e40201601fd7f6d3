"""Consensus of method scores: average ranks per score, then RobustRankAggregate or mean rank.

Restates `liana/method/_pipe_utils/_aggregate.py:7-181` for frames that already share their rows
(every method of an AggregateClass filters with the same `complex_cols` and `expr_prop`, so the
reference's outer merges on the four key columns are 1:1 joins of identical key sets).

Reference behaviour kept on purpose: `_rank_aggregate` converts scores to ranks **once per
(method, score) entry of the specs dict** (`_aggregate.py:91-98`).  `expr_prod` is named by both
Connectome and NATMI, so in the default `rank_aggregate` it is ranked, and then its *ranks* are
ranked again with the sign flipped - this is what the reference's golden frame
(`liana/tests/data/aggregate_rank_rest.csv`) contains, so it is reproduced literally.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.stats import beta, rankdata


def _rank_average(values: np.ndarray, negate: bool, engine) -> np.ndarray:
    """rankdata(method='average') of +-values: on the device (radix sort + tie groups, `lrperm_rank_average`) when an
    engine is given - the product path always gives one - else SciPy (CPU tests of the host logic / checker)."""
    if engine is None:
        return rankdata(values * -1 if negate else values, method="average")
    if np.isnan(values).any():                    # nan_policy='propagate': every rank is NaN
        return np.full(values.shape[0], np.nan)
    return engine.rank_average(values, negate)


def rank_aggregate_scores(lr_res: pd.DataFrame, specs: dict, aggregate_method: str = "rra", engine=None) -> np.ndarray:
    """`_rank_aggregate` (`_aggregate.py:70-107`): specs = {method_name: (score_name, ascending)}."""
    assert aggregate_method in ["rra", "mean"]
    ranks = {}
    # float64 ranks whatever the score dtype: the pinned scipy 1.10.1 (`poetry.lock:3574`) returns float64 from
    # rankdata; newer scipy returns float32 ranks for float32 scores, which would move the result by ~1e-7
    cols = {name: np.asarray(lr_res[name].values, dtype=np.float64) for name in {v[0] for v in specs.values()}}
    for spec in specs:
        score_name, ascending = specs[spec]
        cur = ranks.get(score_name, cols[score_name])
        ranks[score_name] = _rank_average(cur, not ascending, engine)
    # the reference iterates a *set* of names for the column order; RRA sorts every row and the mean
    # is symmetric, so the order of the columns does not change the result
    rmat = np.stack([ranks[name] for name in ranks], axis=1)
    if aggregate_method == "rra":
        if engine is None:
            return robust_rank_aggregate(rmat)
        # column maxima from the contiguous rank vectors (np.max propagates NaN like the axis-0 maximum of the matrix)
        col_max = [np.max(ranks[name]) if len(ranks[name]) else 1.0 for name in ranks]
        return engine.rra(rmat, col_max)
    return np.mean(rmat, axis=1) / rmat.shape[0]


def robust_rank_aggregate(rmat: np.ndarray) -> np.ndarray:
    """RRA of Kolde et al. 2012 as in `_aggregate.py:110-181`: ranks / column max, sort each row,
    beta.cdf(r_(k); k, n-k+1), row minimum, Bonferroni over the n methods, clipped to [0, 1]."""
    rmat = rmat / np.max(rmat, axis=0)
    n = rmat.shape[1]
    rmat = np.sort(rmat, axis=1)
    k = np.arange(n) + 1
    p = beta.cdf(rmat, k[None, :], (n - k + 1)[None, :])
    return np.clip(np.min(p, axis=1) * n, a_min=0, a_max=1)


def aggregate(lr_res: pd.DataFrame, consensus, aggregate_method: str = "rra", consensus_opts=None, engine=None,
              sort: bool = True) -> pd.DataFrame:
    """`_aggregate` (`_aggregate.py:7-67`) on the already-joined frame (key columns + every method's
    score columns).  Adds `consensus.specificity` / `consensus.magnitude` and sorts by the last one."""
    if consensus_opts is None:
        consensus_opts = ["Magnitude", "Specificity"]
    order_col = ""
    if "Specificity" in consensus_opts:
        lr_res[consensus.specificity] = rank_aggregate_scores(lr_res, consensus.specificity_specs, aggregate_method, engine)
        order_col = consensus.specificity
    if "Magnitude" in consensus_opts:
        lr_res[consensus.magnitude] = rank_aggregate_scores(lr_res, consensus.magnitude_specs, aggregate_method, engine)
        order_col = consensus.magnitude
    if not sort:                                 # the caller orders the frame (once) itself
        return lr_res
    from ._pipeline import sort_frame
    return sort_frame(engine, lr_res, order_col, True)
