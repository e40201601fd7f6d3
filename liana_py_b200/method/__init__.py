"""`li.mt`-style namespace: cellphonedb, geometric_mean, rank_aggregate (+ AggregateClass) and the
non-permutation methods the default consensus ranks (connectome, logfc, natmi, singlecellsignalr)."""
from ._methods import Method, MethodMeta, NullCounts, cellphonedb, geometric_mean, cellchat  # noqa: F401
from ._scores import connectome, logfc, natmi, singlecellsignalr  # noqa: F401
from ._rank_aggregate import AggregateClass, aggregate_meta, rank_aggregate  # noqa: F401
from ._dist import DistContext  # noqa: F401
