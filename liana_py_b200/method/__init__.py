"""`li.mt`-style namespace: cellphonedb, geometric_mean, rank_aggregate (+ AggregateClass) and the
non-permutation methods the default consensus ranks (connectome, logfc, natmi, singlecellsignalr)."""
from ._methods import Method, MethodMeta, NullCounts, cellphonedb, geometric_mean, cellchat  # noqa: F401
from ._scores import connectome, logfc, natmi, singlecellsignalr, scseqcomm  # noqa: F401
from ._rank_aggregate import AggregateClass, aggregate_meta, rank_aggregate  # noqa: F401
from ._dist import DistContext, bind_to_gpu_numa  # noqa: F401


def show_methods():
    """Methods available here, like `liana.method.show_methods()` (`liana/method/__init__.py:17-19`)."""
    import pandas as pd
    return pd.concat([m.get_meta() for m in _rank_aggregate_methods + [rank_aggregate, geometric_mean, scseqcomm, cellchat]])


def get_method_scores():
    """{score column: ascending?} of every method score (`liana/method/__init__.py:21-32`)."""
    out = {}
    for m in _rank_aggregate_methods + [rank_aggregate, geometric_mean, scseqcomm, cellchat]:
        if m.specificity is not None:
            out[m.specificity] = m.specificity_ascending
    for m in _rank_aggregate_methods + [rank_aggregate, geometric_mean, scseqcomm, cellchat]:
        if m.magnitude is not None:
            out[m.magnitude] = m.magnitude_ascending
    return out


from ._rank_aggregate import _methods as _rank_aggregate_methods  # noqa: E402
