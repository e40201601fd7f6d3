"""`li.mt`-style namespace: cellphonedb, geometric_mean, rank_aggregate (+ AggregateClass)."""
from ._methods import Method, MethodMeta, NullCounts, cellphonedb, geometric_mean  # noqa: F401
from ._dist import DistContext  # noqa: F401
