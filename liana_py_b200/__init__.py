"""B200-native permutation-null engine for ligand-receptor scoring (drop-in for the hot path of
saezlab/liana-py: li.mt.cellphonedb / li.mt.geometric_mean / rank_aggregate).

    import liana_py_b200 as li
    li.mt.cellphonedb(adata, groupby="cell_type", use_raw=False, n_perms=1000)
"""
from . import method as mt  # noqa: F401
from . import resource as rs  # noqa: F401

__version__ = "0.1.0"
