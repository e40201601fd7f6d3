"""The non-permutation methods the default `li.mt.rank_aggregate` ranks next to CellPhoneDB's
permutation p-values: Connectome, log2FC, NATMI, SingleCellSignalR; and scSeqComm's inter-cellular score
(`liana/method/sc/_scseqcomm.py:4-21`).

Same score formulas, column names and MethodMeta records as the reference
(`liana/method/sc/_connectome.py:5-22`, `_logfc.py:4-6`, `_natmi.py:4-17`,
`_singlecellsignalr.py:6-22`).  They are O(T) float32 arithmetic on columns of the result frame;
the per-(cluster, gene) statistics behind those columns (z-score means, 1-vs-rest log2FC, mean
sums, matrix mean) come from the engine's observed-moment kernels (`_pipeline.PipelineState`).
"""
from __future__ import annotations

import numpy as np

from .. import _constants as K
from ._methods import Method, MethodMeta

_LM, _RM = K.LIGAND_MEANS, K.RECEPTOR_MEANS


def _connectome_score(x):
    expr_prod = x[_LM].values * x[_RM].values
    scaled_weight = np.mean((x["ligand_zscores"].values, x["receptor_zscores"].values), axis=0)
    return expr_prod, scaled_weight


def _logfc_score(x):
    return None, np.mean((x["ligand_logfc"].values, x["receptor_logfc"].values), axis=0)


def _natmi_score(x):
    expr_prod = x[_LM].values * x[_RM].values
    spec_weight = (x[_LM].values / x["ligand_means_sums"].values) * (x[_RM].values / x["receptor_means_sums"].values)
    return expr_prod, spec_weight


def _sca_score(x):
    lr_sqrt = np.sqrt(x[_LM].values) * np.sqrt(x[_RM].values)
    return lr_sqrt / (lr_sqrt + x["mat_mean"].values), None


_connectome = MethodMeta(method_name="Connectome", complex_cols=[_LM, _RM],
                         add_cols=["ligand_zscores", "receptor_zscores"], fun=_connectome_score,
                         magnitude="expr_prod", magnitude_ascending=False, specificity="scaled_weight",
                         specificity_ascending=False, permute=False,
                         reference="Raredon et al., 2022. Scientific reports 12(1).")
_logfc = MethodMeta(method_name="log2FC", complex_cols=[_LM, _RM], add_cols=["ligand_logfc", "receptor_logfc"],
                    fun=_logfc_score, magnitude=None, magnitude_ascending=None, specificity="lr_logfc",
                    specificity_ascending=False, permute=False,
                    reference="Dimitrov et al., 2022. Nature Communications 13(1).")
_natmi = MethodMeta(method_name="NATMI", complex_cols=[_LM, _RM],
                    add_cols=["ligand_means_sums", "receptor_means_sums"], fun=_natmi_score,
                    magnitude="expr_prod", magnitude_ascending=False, specificity="spec_weight",
                    specificity_ascending=False, permute=False,
                    reference="Hou et al., 2020. Nature communications 11(1).")
_singlecellsignalr = MethodMeta(method_name="SingleCellSignalR", complex_cols=[_LM, _RM], add_cols=["mat_mean"],
                                fun=_sca_score, magnitude="lrscore", magnitude_ascending=False, specificity=None,
                                specificity_ascending=None, permute=False,
                                reference="Cabello-Aguilar et al., 2020. Nucleic Acids Research 48(10).")

def _inter_score(x):
    return np.minimum(x["ligand_cdf"].values, x["receptor_cdf"].values), None


_scseqcomm = MethodMeta(method_name="scSeqComm", complex_cols=[_LM, _RM], add_cols=["ligand_cdf", "receptor_cdf"],
                        fun=_inter_score, magnitude="inter_score", magnitude_ascending=False, specificity=None,
                        specificity_ascending=None, permute=False,
                        reference="Baruzzo et al., 2022. Bioinformatics 38(7).")

connectome = Method(_method=_connectome)
scseqcomm = Method(_method=_scseqcomm)
logfc = Method(_method=_logfc)
natmi = Method(_method=_natmi)
singlecellsignalr = Method(_method=_singlecellsignalr)
