"""Defaults and column names kept identical to the reference (values from
`liana/_constants.py:3-79` of saezlab/liana-py) so that results drop in."""
from math import e as _E


class DefaultValues:
    logbase = _E
    min_cells = 5
    expr_prop = 0.1
    n_perms = 1000
    seed = 1337
    de_method = 't-test'
    resource_name = 'consensus'
    resource = None
    interactions = None
    layer = None
    use_raw = True
    verbose = False
    return_all_lrs = False
    supp_columns = None
    inplace = True
    groupby_pairs = None
    complex_sep = '_'


class Keys:
    uns_key = 'liana_res'


SOURCE, TARGET = 'source', 'target'
LIGAND, RECEPTOR = 'ligand', 'receptor'
LIGAND_COMPLEX, RECEPTOR_COMPLEX = 'ligand_complex', 'receptor_complex'
PRIMARY = [SOURCE, TARGET, LIGAND_COMPLEX, RECEPTOR_COMPLEX]
LIGAND_MEANS, RECEPTOR_MEANS = 'ligand_means', 'receptor_means'
LIGAND_PROPS, RECEPTOR_PROPS = 'ligand_props', 'receptor_props'
LRS_TO_KEEP = 'lrs_to_keep'
LABEL = '@label'
LIGAND_TRIMEAN, RECEPTOR_TRIMEAN = 'ligand_trimean', 'receptor_trimean'
MAT_MAX = 'mat_max'
