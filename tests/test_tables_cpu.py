"""CPU: the host logic between the observed statistics and the score function - resource tables, expr_prop filter,
min-subunit resolution (`method/_tables.py`, restating `filter_reassemble_complexes` + `_join_stats`,
liana/resource/_reassemble_complexes.py:11-101, liana/method/_pipe_utils/_common.py:5-39) - against the frames the
reference's own code produced (tests/golden/*_cellphonedb.csv.gz).  The device work is stood in by the CPU oracle
(observed means / stored counts) and by the NumPy formulation of the row resolution (engine=None), so the whole frame
up to the p-values is rebuilt without a GPU and compared column by column."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from liana_py_b200.method import _pre, _tables
from liana_py_b200.resource import select_resource
from liana_py_b200.testing._anndata_lite import AnnDataLite
from oracle import c_oracle, numpy_port
from test_host_logic import SciPyStage

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = ["source", "target", "ligand_complex", "receptor_complex"]


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    X = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=tuple(z["in_shape"]))
    obs = pd.DataFrame({"cluster": pd.Categorical(z["in_labels"])}, index=[f"cell{i}" for i in range(X.shape[0])])
    return AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(z["in_var_names"]))), int(z["n_perms"]), int(z["seed"]), float(z["expr_prop"])


def _host_pipeline(adata, expr_prop, resource):
    prep = _pre.prepare_input(adata, "cluster", 5, SciPyStage())
    tables = _tables.build_resource_tables(resource, prep.var_names)
    cols = prep.var_cols[np.searchsorted(prep.var_names, tables.entities)]
    rows = adata.obs.index.get_indexer(prep.obs_index)
    Xs = adata.X[rows][:, cols].tocsr()
    Xs.sort_indices()
    L = len(prep.categories)
    means, stored, sizes = c_oracle.observed(Xs, prep.labels, L)
    props = stored / sizes[:, None].astype(np.float64)
    tri = _tables.resolve_triples(tables, means, props, expr_prop, None, False, engine=None)
    frame = pd.DataFrame({
        "ligand": tables.entities[tri.lig_gene], "ligand_complex": tables.ligand_complex[tri.inter],
        "ligand_means": tri.ligand_means, "ligand_props": tri.ligand_props,
        "receptor": tables.entities[tri.rec_gene], "receptor_complex": tables.receptor_complex[tri.inter],
        "receptor_means": tri.receptor_means, "receptor_props": tri.receptor_props,
        "source": np.asarray(prep.categories)[tri.src], "target": np.asarray(prep.categories)[tri.tgt]})
    return frame, tri, Xs, prep, L


@pytest.mark.parametrize("name", ["toy700", "small300", "ragged90"])
def test_host_tables_rebuild_the_reference_frame(name):
    adata, P, seed, ep = _load(name)
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_cellphonedb.csv.gz"), float_precision="round_trip")
    frame, tri, Xs, prep, L = _host_pipeline(adata, ep, select_resource("consensus"))
    a = frame.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:                 # the kept rows and the subunit chosen for every complex
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in ("ligand_means", "receptor_means"):
        assert a[c].dtype == np.float32 and np.array_equal(a[c].values, b[c].values.astype(np.float32)), c
    for c in ("ligand_props", "receptor_props"):
        assert np.array_equal(a[c].values, b[c].values), c
    # and the scores: observed magnitude + permutation p-values from the oracle's cube through the resolved indices
    obs = numpy_port.cpdb_observed(tri.ligand_means, tri.receptor_means)
    perm_idx = numpy_port.perm_indices(Xs.shape[0], P, seed)
    cube = c_oracle.cube(Xs, prep.labels, L, perm_idx)
    counts = numpy_port.counts_from_cube(cube, tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, "cellphonedb")
    frame["lr_means"], frame["cellphone_pvals"] = obs, counts / P
    a = frame.sort_values(KEY).reset_index(drop=True)
    assert np.array_equal(a["lr_means"].values, b["lr_means"].values.astype(np.float32))
    assert np.array_equal(a["cellphone_pvals"].values, b["cellphone_pvals"].values)


def test_resolve_rules_on_a_hand_made_case():
    """min-subunit choice (first minimal subunit in '_' order), the all-subunits expr_prop rule, pair masks and
    `return_all_lrs` (first subunits, `keep` flags) on a 2-cluster toy where every answer is known by hand."""
    res = pd.DataFrame({"ligand": ["A_B", "C", "A"], "receptor": ["D", "D_E", "Z"]})       # Z is absent: dropped
    names = np.array(["A", "B", "C", "D", "E"])
    t = _tables.build_resource_tables(res, names)
    assert list(t.ligand_complex) == ["A_B", "C"] and list(t.entities) == list(names)
    means = np.array([[1.0, 1.0, 3.0, 2.0, 0.5],        # cluster 0: A == B (tie -> A), E < D
                      [4.0, 2.0, 0.0, 1.0, 1.5]], dtype=np.float32)   # cluster 1: B < A, D < E
    props = np.array([[0.5, 0.05, 0.9, 0.4, 0.3],       # cluster 0: B fails expr_prop=0.1 -> A_B not expressed as source 0
                      [0.6, 0.7, 0.0, 0.2, 0.8]])
    tri = _tables.resolve_triples(t, means, props, 0.1, None, False, engine=None)
    got = {(int(s), int(g), int(i)): (int(lg), int(rg)) for s, g, i, lg, rg in zip(tri.src, tri.tgt, tri.inter, tri.lig_gene, tri.rec_gene)}
    # interaction 0 (A_B -> D): source must be cluster 1 (B fails in 0); ligand subunit there = B (mean 2 < 4)
    # interaction 1 (C -> D_E): source must be cluster 0 (C has prop 0 in 1); receptor subunit = E in 0, D in 1
    assert got == {(1, 0, 0): (1, 3), (1, 1, 0): (1, 3), (0, 0, 1): (2, 4), (0, 1, 1): (2, 3)}
    mask = np.array([[False, True], [False, False]])
    only = _tables.resolve_triples(t, means, props, 0.1, mask, False, engine=None)
    assert list(zip(only.src, only.tgt, only.inter)) == [(0, 1, 1)]
    full = _tables.resolve_triples(t, means, props, 0.1, None, True, engine=None)
    assert len(full.inter) == 2 * 2 * 2 and int(full.keep.sum()) == 4
    assert set(full.lig_gene[full.inter == 0]) == {0} and set(full.rec_gene[full.inter == 1]) == {3}   # first subunits


@pytest.mark.parametrize("name", ["small300", "ragged90"])
def test_host_tables_rebuild_the_reference_cellchat_frame(name):
    """The same for CellChat: complexes resolved by the cluster TRIMEANS of X / mat_max (`complex_cols` of
    liana/method/sc/_cellchat.py:34-52), `mat_max` from the prepared matrix, L-R probabilities and p-values from the
    oracle's trimean cube through the resolved indices - against the reference's frame."""
    z = np.load(os.path.join(GOLDEN, f"{name}_cellchat.npz"))
    P, seed, ep = int(z["n_perms"]), int(z["seed"]), float(z["expr_prop"])
    X = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=tuple(z["in_shape"]))
    obs_df = pd.DataFrame({"cluster": pd.Categorical(z["in_labels"])}, index=[f"cell{i}" for i in range(X.shape[0])])
    adata = AnnDataLite(X=X, obs=obs_df, var=pd.DataFrame(index=pd.Index(z["in_var_names"])))
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_cellchat.csv.gz"), float_precision="round_trip")
    prep = _pre.prepare_input(adata, "cluster", 5, SciPyStage(), want_max=True)
    assert prep.mat_max == np.float32(z["mat_max"]) == np.float32(gold["mat_max"].iloc[0])
    tables = _tables.build_resource_tables(select_resource("consensus"), prep.var_names)
    cols = prep.var_cols[np.searchsorted(prep.var_names, tables.entities)]
    rows = adata.obs.index.get_indexer(prep.obs_index)
    Xs = X[rows][:, cols].tocsr()
    Xs.sort_indices()
    L = len(prep.categories)
    _, stored, sizes = c_oracle.observed(Xs, prep.labels, L)
    props = stored / sizes[:, None].astype(np.float64)
    tm = c_oracle.trimean_cube(Xs, prep.labels, L, None, prep.mat_max)[0]              # observed trimeans [L, G]
    tri = _tables.resolve_triples(tables, tm, props, ep, None, False, engine=None)
    obs = numpy_port.lr_probability((tri.ligand_means, tri.receptor_means))
    cube = c_oracle.trimean_cube(Xs, prep.labels, L, numpy_port.perm_indices(Xs.shape[0], P, seed), prep.mat_max)
    counts = numpy_port.counts_cellchat(cube, tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs)
    cats = np.asarray(prep.categories)
    frame = pd.DataFrame({"source": cats[tri.src], "target": cats[tri.tgt], "ligand_complex": tables.ligand_complex[tri.inter],
                          "receptor_complex": tables.receptor_complex[tri.inter], "ligand": tables.entities[tri.lig_gene],
                          "receptor": tables.entities[tri.rec_gene], "ligand_trimean": tri.ligand_means,
                          "receptor_trimean": tri.receptor_means, "ligand_props": tri.ligand_props,
                          "receptor_props": tri.receptor_props, "lr_probs": obs, "cellchat_pvals": counts / P})
    a = frame.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in ("ligand_trimean", "receptor_trimean", "ligand_props", "receptor_props", "lr_probs", "cellchat_pvals"):
        assert np.array_equal(a[c].values, b[c].values), c
