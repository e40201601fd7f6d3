"""GPU: the drop-in li.mt.cellphonedb / geometric_mean frames against frames produced by the
reference's own code on the same input (tests/golden/*_cellphonedb.csv.gz, scripts/make_golden.py)."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = ["source", "target", "ligand_complex", "receptor_complex"]


def load_input(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    X = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=tuple(z["in_shape"]))
    obs = pd.DataFrame({"cluster": pd.Categorical(z["in_labels"])}, index=[f"cell{i}" for i in range(X.shape[0])])
    var = pd.DataFrame(index=pd.Index(z["in_var_names"]))
    return AnnDataLite(X=X, obs=obs, var=var), int(z["n_perms"]), int(z["seed"]), float(z["expr_prop"])


def compare(res, gold, mag, spec):
    assert list(res.columns) == list(gold.columns)
    a = res.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in ["ligand_means", "receptor_means", mag]:
        assert a[c].dtype == np.float32
        assert np.array_equal(a[c].values, b[c].values.astype(np.float32)), c     # bit-identical float32
    for c in ["ligand_props", "receptor_props", spec]:
        assert a[c].dtype == np.float64
        assert np.array_equal(a[c].values, b[c].values), c
    # same sort key as the reference: magnitude descending
    assert np.all(np.diff(res[mag].values) <= 0)


@pytest.mark.parametrize("name", ["toy700", "small300", "ragged90"])
def test_cellphonedb_frame_identical_to_reference(name):
    adata, P, seed, ep = load_input(name)
    res = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                            inplace=False, perm_source="reference")
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")
    # inplace contract
    li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                      perm_source="reference")
    assert adata.uns["liana_res"].shape == res.shape


@pytest.mark.parametrize("name", ["toy700", "small300", "ragged90"])
def test_geometric_mean_frame_identical_to_reference(name):
    adata, P, seed, ep = load_input(name)
    res = li.mt.geometric_mean(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                               inplace=False, perm_source="reference")
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_geometric_mean.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_gmeans", "gmean_pvals")


def test_n_perms_none_and_philox_default():
    adata, P, seed, ep = load_input("small300")
    res = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=None, inplace=False)
    assert "cellphone_pvals" not in res.columns and "lr_means" in res.columns
    ph = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=200, inplace=False)   # Philox, fast
    ex = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=200, inplace=False, order="exact")
    a = ph.sort_values(KEY)["cellphone_pvals"].values
    b = ex.sort_values(KEY)["cellphone_pvals"].values
    assert np.abs(a - b).max() <= 1.0 / 200 + 1e-12          # same permutations, summation order differs
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_cellphonedb.csv.gz"), float_precision="round_trip").sort_values(KEY)
    # a different RNG gives statistically equivalent p-values
    assert abs(a.mean() - gold["cellphone_pvals"].values.mean()) < 0.05


def test_error_conventions():
    adata, P, seed, ep = load_input("ragged90")
    with pytest.raises(ValueError, match="not found"):
        li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, resource_name="nope", n_perms=2)
    with pytest.raises(ValueError, match="raw"):
        li.mt.cellphonedb(adata, groupby="cluster", use_raw=True, n_perms=2)
    with pytest.raises(ValueError, match="layer"):
        li.mt.cellphonedb(adata, groupby="cluster", use_raw=True, layer="x", n_perms=2)
    with pytest.raises(AssertionError):
        li.mt.cellphonedb(adata, groupby="nope", use_raw=False, n_perms=2)
    bad = adata.copy()
    bad.var_names = [f"X{i}" for i in range(bad.shape[1])]
    with pytest.raises(ValueError, match="Too few features"):
        li.mt.cellphonedb(bad, groupby="cluster", use_raw=False, n_perms=2)


def test_by_sample_matches_per_sample_calls():
    adata, P, seed, ep = load_input("toy700")
    rng = np.random.default_rng(0)
    adata.obs["sample"] = rng.choice(["A", "B", "C"], size=adata.shape[0])     # not categorical yet
    kw = dict(groupby="cluster", use_raw=False, n_perms=20, seed=seed, expr_prop=ep, perm_source="reference")
    res = li.mt.cellphonedb.by_sample(adata, "sample", inplace=False, **kw)
    assert res.columns[0] == "sample" and list(res.columns[1:]) == [
        "ligand", "ligand_complex", "ligand_means", "ligand_props", "receptor", "receptor_complex",
        "receptor_means", "receptor_props", "source", "target", "lr_means", "cellphone_pvals"]
    assert list(dict.fromkeys(res["sample"])) == ["A", "B", "C"]
    for s in ["A", "B", "C"]:
        sub = adata[np.asarray(adata.obs["sample"] == s)].copy()
        one = li.mt.cellphonedb(sub, inplace=False, **kw)
        got = res[res["sample"] == s].drop(columns="sample")
        a = got.sort_values(KEY).reset_index(drop=True)
        b = one.sort_values(KEY).reset_index(drop=True)
        assert np.array_equal(a["lr_means"].values, b["lr_means"].values)
        assert np.array_equal(a["cellphone_pvals"].values, b["cellphone_pvals"].values)
    with pytest.raises(ValueError, match="was not found"):
        li.mt.cellphonedb.by_sample(adata, "nope", **kw)
    li.mt.cellphonedb.by_sample(adata, "sample", key_added="by_s", **kw)
    assert adata.uns["by_s"].shape == res.shape


# ---- rank_aggregate (default consensus) against frames produced by the reference's own code ----------------
SCORES = ["lr_means", "cellphone_pvals", "expr_prod", "scaled_weight", "lr_logfc", "spec_weight", "lrscore"]


def compare_aggregate(res, gold, rank_cols, min_frac=0.999, max_rel=2e-2, check_order=True):
    assert list(res.columns) == list(gold.columns)
    a = res.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in [c for c in SCORES if c in gold.columns]:
        if c in ("lr_means",):
            assert np.array_equal(a[c].values, b[c].values.astype(np.float32)), c           # bit-identical
        elif c == "cellphone_pvals":
            assert np.array_equal(a[c].values, b[c].values), c                              # bit-identical
        else:
            # float32 statistics computed by a different (float64-accumulating) route than the reference's
            # float32 NumPy/SciPy one: same tolerance as the reference's own frame test (rtol 1e-4,
            # liana/tests/test_rank_aggregate.py:41-47), observed ~1e-6
            assert a[c].dtype == np.float32, c
            np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-6, err_msg=c)
    for c in rank_cols:
        # The consensus is a function of the RANKS of the float32 scores.  A pair of scores the reference's
        # float32 arithmetic separates by ~1 ulp can come out in the other order here (the statistics are
        # accumulated in float64), which moves two ranks by one and the aggregate of those rows by ~1/n.
        # So: (nearly) all rows to the reference's own test tolerance, the rest within a one-rank shift.
        rel = np.abs(a[c].values - b[c].values) / np.maximum(np.abs(b[c].values), 1e-12)
        assert (rel <= 1e-4).mean() >= min_frac, (c, float((rel <= 1e-4).mean()))
        assert rel.max() <= max_rel, (c, float(rel.max()))
        # rank ORDER identical except among rows whose aggregate ranks are numerically tied
        oa, ob = np.argsort(a[c].values, kind="stable"), np.argsort(b[c].values, kind="stable")
        moved = oa != ob
        if moved.any() and check_order:
            assert np.allclose(b[c].values[oa][moved], b[c].values[ob][moved], rtol=max_rel, atol=1e-9), c
    assert np.all(np.diff(res[rank_cols[-1]].values) >= 0)            # sorted by magnitude_rank ascending


@pytest.mark.parametrize("name", ["toy700", "small300", "ragged90"])
def test_rank_aggregate_frame_matches_reference(name):
    adata, P, seed, ep = load_input(name)
    res = li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                               inplace=False, perm_source="reference")
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_rank_aggregate.csv.gz"), float_precision="round_trip")
    compare_aggregate(res, gold, ["specificity_rank", "magnitude_rank"])


def test_rank_aggregate_mean_and_noperms():
    adata, P, seed, ep = load_input("small300")
    res = li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                               inplace=False, perm_source="reference", aggregate_method="mean")
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_rank_aggregate_aggregate_method-mean.csv.gz"),
                       float_precision="round_trip")
    compare_aggregate(res, gold, ["specificity_rank", "magnitude_rank"])
    res = li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=None, seed=seed, expr_prop=ep,
                               inplace=False)
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_rank_aggregate_noperms.csv.gz"), float_precision="round_trip")
    compare_aggregate(res, gold, ["magnitude_rank"])
    # the method objects are not mutated by an n_perms=None call (the reference sets specificity=None for good)
    assert li.mt.cellphonedb.specificity == "cellphone_pvals"


def test_custom_consensus_counts_both_nulls_in_one_pass():
    """BASELINE.json configs[3]: cellphonedb + geometric_mean permutation nulls plus non-permutation scores in
    one AggregateClass; both nulls come from the same permutations and must equal the single-method runs."""
    adata, P, seed, ep = load_input("small300")
    methods = [li.mt.cellphonedb, li.mt.geometric_mean, li.mt.connectome, li.mt.logfc, li.mt.natmi,
               li.mt.singlecellsignalr]
    custom = li.mt.AggregateClass(li.mt.aggregate_meta, methods=methods)
    res = custom(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep, inplace=False,
                 perm_source="reference")
    assert {"lr_means", "cellphone_pvals", "lr_gmeans", "gmean_pvals", "specificity_rank", "magnitude_rank"} <= set(res.columns)
    cp = pd.read_csv(os.path.join(GOLDEN, "small300_cellphonedb.csv.gz"), float_precision="round_trip")
    gm = pd.read_csv(os.path.join(GOLDEN, "small300_geometric_mean.csv.gz"), float_precision="round_trip")
    m = res.merge(cp[KEY + ["cellphone_pvals"]], on=KEY, suffixes=("", "_ref")).merge(
        gm[KEY + ["gmean_pvals"]], on=KEY, suffixes=("", "_ref"))
    assert len(m) == len(res) == len(cp)
    assert np.array_equal(m["cellphone_pvals"].values, m["cellphone_pvals_ref"].values)
    assert np.array_equal(m["gmean_pvals"].values, m["gmean_pvals_ref"].values)


def test_single_non_permutation_methods_run():
    adata, P, seed, ep = load_input("small300")
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_rank_aggregate.csv.gz"), float_precision="round_trip")
    for method, cols in ((li.mt.connectome, ["expr_prod", "scaled_weight"]), (li.mt.logfc, ["lr_logfc"]),
                         (li.mt.natmi, ["expr_prod", "spec_weight"]), (li.mt.singlecellsignalr, ["lrscore"])):
        res = method(adata, groupby="cluster", use_raw=False, expr_prop=ep, inplace=False)
        m = res.merge(gold[KEY + cols], on=KEY, suffixes=("", "_ref"))
        assert len(m) == len(gold)
        for c in cols:
            np.testing.assert_allclose(m[c].values, m[c + "_ref"].values, rtol=1e-4, atol=2e-6)


# ---- the optional arguments of the call signature, against the reference's frames ---------------------------
@pytest.mark.parametrize("name", ["pairs", "interactions", "resource", "raw", "layer", "resname"])
def test_call_signature_variants_identical_to_reference(name):
    from variants import variant_inputs
    adata, variants = variant_inputs()
    res = li.mt.cellphonedb(adata, groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                            perm_source="reference", **variants[name])
    gold = pd.read_csv(os.path.join(GOLDEN, f"variant_{name}_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")


def test_supp_columns_identical_to_reference():
    """`supp_columns` (`_liana_pipe.py:111-114`): the statistic columns of the other methods ride along with
    CellPhoneDB's; same column order (np.union1d = sorted, then the two scores) and values as the reference's frame."""
    from variants import variant_inputs, SUPP_COLUMNS
    adata, variants = variant_inputs()
    res = li.mt.cellphonedb(adata, groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                            perm_source="reference", **variants["supp"])
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_supp_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")
    a = res.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    for c in ("mat_mean", "mat_max"):
        assert a[c].dtype == np.float32 and np.array_equal(a[c].values, b[c].values.astype(np.float32)), c
    for c in ("ligand_trimean", "receptor_trimean"):
        assert np.array_equal(a[c].values, b[c].values), c                          # exact selection: bit-identical
    for c in [c for c in SUPP_COLUMNS if c not in ("mat_mean", "mat_max", "ligand_trimean", "receptor_trimean")]:
        # float32 / float64 statistics accumulated by another route: the reference's own frame tolerance
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-6, err_msg=c)
    with pytest.raises(KeyError):                      # trimeans exist only when mat_max is requested too
        li.mt.cellphonedb(adata, groupby="cluster", n_perms=None, use_raw=False, inplace=False,
                          supp_columns=["ligand_trimean"])
    with pytest.raises(NotImplementedError):
        li.mt.cellphonedb(adata, groupby="cluster", n_perms=None, use_raw=False, inplace=False,
                          supp_columns=["ligand_pvals"])


def test_mudata_input_identical_to_reference():
    """A MuData as input (`_liana_pipe.py:126-129`, `liana/utils/mdata_to_anndata.py`): the two modalities named in
    `mdata_kwargs` side by side, `use_raw` / `layer` overridden, the result written into the MuData's `.uns`."""
    from variants import mdata_inputs
    mdata, mkw = mdata_inputs()
    kw = dict(groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, perm_source="reference", use_raw=True, mdata_kwargs=mkw)
    res = li.mt.cellphonedb(mdata, inplace=False, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_mdata_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")
    li.mt.cellphonedb(mdata, **kw)
    assert mdata.uns["liana_res"].shape == res.shape
    agg = li.mt.rank_aggregate(mdata, groupby="cluster", n_perms=None, use_raw=False, inplace=False, mdata_kwargs=mkw)
    assert len(agg) == len(res) and "magnitude_rank" in agg.columns
    # by_sample on a MuData: the samples are row subsets of every modality (`_Method.py:138-152`)
    rng = np.random.default_rng(1)
    mdata.obs["sample"] = pd.Categorical(rng.choice(["s1", "s2"], size=mdata.obs.shape[0]))
    per = li.mt.cellphonedb.by_sample(mdata, "sample", inplace=False, **kw)
    one = li.mt.cellphonedb(mdata[np.asarray(mdata.obs["sample"] == "s2")].copy(), inplace=False, **kw)
    got = per[per["sample"] == "s2"].drop(columns="sample").sort_values(KEY).reset_index(drop=True)
    assert got.equals(one.sort_values(KEY).reset_index(drop=True))
    with pytest.raises(ValueError, match="is not in the mdata"):
        li.mt.cellphonedb(mdata, groupby="cluster", n_perms=2, mdata_kwargs=dict(x_mod="rna", y_mod="nope"))
    with pytest.raises(TypeError):                    # x_mod / y_mod are required, as in the reference
        li.mt.cellphonedb(mdata, groupby="cluster", n_perms=2)


def test_dense_matrix_input_equals_sparse_input():
    """`adata.X` as a dense 2-D array (`_choose_mtx_rep` converts it with `csr_matrix(X)`, `_pre.py:259-262`): converted
    on the device, same frame as for the CSR input - and so as the reference's."""
    adata, P, seed, ep = load_input("small300")
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_cellphonedb.csv.gz"), float_precision="round_trip")
    for dtype in (np.float32, np.float64):
        dense = AnnDataLite(X=adata.X, obs=adata.obs.copy(), var=adata.var.copy())
        dense.X = adata.X.toarray().astype(dtype)
        res = li.mt.cellphonedb(dense, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                                inplace=False, perm_source="reference")
        compare(res, gold, "lr_means", "cellphone_pvals")


def test_non_canonical_matrix_gives_the_frame_of_its_canonical_form():
    """A CSR matrix whose rows are not sorted, or that stores a (cell, gene) entry in two pieces: SciPy - and so the
    reference - computes with the matrix they add up to.  The staging pass counts the out-of-order entries
    (`n_unordered`); only then does the host canonicalise (`sum_duplicates`) and stage again - with no pass of its own
    over every index.  Both ways in: a fresh matrix (no cached SciPy flag: the device finds out) and one whose flag says no."""
    adata, P, seed, ep = load_input("small300")
    gold = pd.read_csv(os.path.join(GOLDEN, "small300_cellphonedb.csv.gz"), float_precision="round_trip")
    X = adata.X.tocsr()
    rng = np.random.default_rng(5)

    def shuffled():
        ix, dt = X.indices.copy(), X.data.copy()
        for r in range(X.shape[0]):
            s, e = X.indptr[r], X.indptr[r + 1]
            p = rng.permutation(e - s)
            ix[s:e], dt[s:e] = ix[s:e][p], dt[s:e][p]
        return sparse.csr_matrix((dt, ix, X.indptr.copy()), shape=X.shape)

    def split():                                         # every entry as two halves (exact in float32)
        rep = np.repeat(np.arange(X.nnz), 2)
        return sparse.csr_matrix((X.data[rep] * np.float32(0.5), X.indices[rep], X.indptr * 2), shape=X.shape)

    for make in (shuffled, split):
        for ask_scipy_first in (False, True):
            Xn = make()
            assert "_has_canonical_format" not in Xn.__dict__
            if ask_scipy_first:
                assert not Xn.has_canonical_format
            other = AnnDataLite(X=Xn, obs=adata.obs.copy(), var=adata.var.copy())
            res = li.mt.cellphonedb(other, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                                    inplace=False, perm_source="reference")
            compare(res, gold, "lr_means", "cellphone_pvals")
            assert other.X is Xn and Xn.nnz == (X.nnz if make is shuffled else 2 * X.nnz)      # the caller's matrix is not touched


def test_cells_without_a_label_identical_to_reference():
    """NaN in `groupby`: the reference keeps such cells (`prep_check_adata` filters by category only) - they are
    shuffled with the others (`_get_mean_perms.py:46-52`: an all-False row of labels_mask), count in mat_mean / the
    z-score scaling, and are part of the "rest" of every cluster's log fold change (`_liana_pipe.py:341-360`).  Here
    they are one more, nameless cluster on the engine.  cellphonedb frame and the consensus against the reference's."""
    from variants import nanlabel_inputs
    adata = nanlabel_inputs()
    assert adata.obs["cluster"].isna().sum() > 10
    kw = dict(groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, perm_source="reference", use_raw=False)
    res = li.mt.cellphonedb(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_nanlabels_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")
    agg = li.mt.rank_aggregate(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_nanlabels_rank_aggregate.csv.gz"), float_precision="round_trip")
    compare_aggregate(agg, gold, ["specificity_rank", "magnitude_rank"])
    cc = li.mt.cellchat(adata, **kw)                       # the trimean path with the extra cluster: runs, names only real clusters
    assert set(cc["source"]) <= set(adata.obs["cluster"].cat.categories) and cc["cellchat_pvals"].between(0, 1).all()


def test_signed_values_and_integer_cluster_names_identical_to_reference():
    """Scaled / centred data (a third of the stored values negative: sums, the zero mask of lr_means and the
    quartile selection of CellChat all see both signs) and integer cluster names (`source` / `target` are int64
    columns in the reference's frames)."""
    from variants import signed_inputs
    adata = signed_inputs()
    kw = dict(groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, perm_source="reference", use_raw=False)
    res = li.mt.cellphonedb(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_signed_cellphonedb.csv.gz"), float_precision="round_trip")
    assert res["source"].dtype == np.int64 and res["target"].dtype == np.int64
    compare(res, gold, "lr_means", "cellphone_pvals")
    cc = li.mt.cellchat(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_signed_cellchat.csv.gz"), float_precision="round_trip")
    assert list(cc.columns) == list(gold.columns)
    a = cc.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b) and (a[KEY].values == b[KEY].values).all()
    for c in ("ligand_trimean", "receptor_trimean", "ligand_props", "receptor_props", "lr_probs", "cellchat_pvals"):
        assert np.array_equal(a[c].values, b[c].values), c


def test_geometric_mean_of_signed_values_follows_the_reference_nan_rules():
    """A geometric mean of negative cluster means is NaN in the reference (`scipy.stats.gmean`: log of a negative), a
    NaN never compares `>=`, so such rows get p-value 0 and null means with a negative factor never count
    (`_geometric_mean.py:22-23`, `_get_mean_perms.py:133-137`).  Both comparison modes of the engine follow that."""
    from variants import signed_inputs
    adata = signed_inputs()
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_signed_geometric_mean.csv.gz"), float_precision="round_trip")
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert b["lr_gmeans"].isna().sum() > 50
    for mode in ("product", "literal"):
        with np.errstate(invalid="ignore"):
            res = li.mt.geometric_mean(adata, groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                                       perm_source="reference", use_raw=False, gmean_mode=mode)
        assert list(res.columns) == list(gold.columns)
        a = res.sort_values(KEY).reset_index(drop=True)
        assert len(a) == len(b) and (a[KEY].values == b[KEY].values).all()
        assert np.array_equal(a["lr_gmeans"].values, b["lr_gmeans"].values.astype(np.float32), equal_nan=True), mode
        assert np.array_equal(a["gmean_pvals"].values, b["gmean_pvals"].values), mode
        assert res["lr_gmeans"].isna().values[-int(b["lr_gmeans"].isna().sum()):].all()       # NaN magnitudes sort last


def test_tied_subunits_resolved_like_the_reference():
    """Complexes whose subunits have EQUAL cluster means (copied expression): which subunit represents the complex
    is decided by the tie rule alone - the reference keeps every minimal row and then the first per key
    (`_reduce_complexes` + `drop_duplicates(keep='first')`, `_reassemble_complexes.py:60-101`), i.e. the first minimal
    subunit in the complex's own order; for CellChat the same on the trimeans."""
    from variants import ties_inputs
    adata, n_tied = ties_inputs()
    assert n_tied >= 20
    kw = dict(groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, perm_source="reference", use_raw=False)
    res = li.mt.cellphonedb(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_ties_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")                # includes the `ligand` / `receptor` columns
    cc = li.mt.cellchat(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_ties_cellchat.csv.gz"), float_precision="round_trip")
    a, b = cc.sort_values(KEY).reset_index(drop=True), gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in ("ligand_trimean", "receptor_trimean", "lr_probs", "cellchat_pvals"):
        assert np.array_equal(a[c].values, b[c].values), c


def test_consensus_options_identical_to_reference():
    """`base=2` (the log fold changes undo a log2 instead of a natural log, `_liana_pipe.py:266-273`) with
    `expr_prop=0` (every (pair, interaction) row is scored: T = T_max of this input) and another seed; and a
    magnitude-only consensus (`consensus_opts=['Magnitude']`: no `specificity_rank` column)."""
    from variants import variant_inputs
    adata, _ = variant_inputs()
    kw = dict(groupby="cluster", n_perms=16, seed=3, inplace=False, perm_source="reference", use_raw=False)
    res = li.mt.rank_aggregate(adata, base=2.0, expr_prop=0.0, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_base2_all_rank_aggregate.csv.gz"), float_precision="round_trip")
    assert len(res) == adata.obs["cluster"].nunique() ** 2 * 314
    # With every row scored, thousands of rows are mathematically TIED in the z-score / log2FC columns (a gene that is
    # not expressed in several clusters).  The reference's float32 cluster means of the scaled matrix differ in the
    # last bit with the cluster size, so it ranks such rows by rounding noise (6,832 distinct `scaled_weight` values
    # against 6,302 here, where they tie and share the average rank): 1 % of the rows move by up to 3 % in
    # `specificity_rank` (scripts/consensus_diag.py).  Scores themselves: p-values identical, the rest within 1e-4 / 2e-6.
    compare_aggregate(res, gold, ["magnitude_rank"])
    compare_aggregate(res, gold, ["specificity_rank", "magnitude_rank"], min_frac=0.98, max_rel=5e-2, check_order=False)
    res = li.mt.rank_aggregate(adata, consensus_opts=["Magnitude"], expr_prop=0.1, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_magnitude_only_rank_aggregate.csv.gz"), float_precision="round_trip")
    assert "specificity_rank" not in res.columns
    compare_aggregate(res, gold, ["magnitude_rank"])


def test_by_sample_frames_identical_to_reference():
    """`Method.by_sample` / `AggregateClass.by_sample` against the frames of the reference's own by_sample
    (`sc/_Method.py:92-159`): `sample` first, samples stacked in category order under a fresh RangeIndex, every
    sample's block equal to the reference's block."""
    from variants import sample_inputs
    kw = dict(sample_key="sample", groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
              perm_source="reference", use_raw=False)
    res = li.mt.cellphonedb.by_sample(sample_inputs(), **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_bysample_cellphonedb.csv.gz"), float_precision="round_trip")
    assert list(res.columns) == list(gold.columns) and len(res) == len(gold)
    assert isinstance(res.index, pd.RangeIndex) and list(dict.fromkeys(res["sample"])) == list(dict.fromkeys(gold["sample"]))
    for s_name in ("s1", "s2"):
        compare(res[res["sample"] == s_name].drop(columns="sample"), gold[gold["sample"] == s_name].drop(columns="sample"),
                "lr_means", "cellphone_pvals")
    agg = li.mt.rank_aggregate.by_sample(sample_inputs(), **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_bysample_rank_aggregate.csv.gz"), float_precision="round_trip")
    assert list(agg.columns) == list(gold.columns) and len(agg) == len(gold)
    for s_name in ("s1", "s2"):
        compare_aggregate(agg[agg["sample"] == s_name].drop(columns="sample"),
                          gold[gold["sample"] == s_name].drop(columns="sample"), ["specificity_rank", "magnitude_rank"])


def test_return_all_lrs_frames_identical_to_reference():
    """`return_all_lrs=True` against frames of the reference itself (run under the pinned-pandas emulation of
    oracle/ref_shim.py: its chained in-place `fillna` is a no-op under pandas 3): every (cluster pair, interaction)
    row, complexes represented by their first subunits (`drop_duplicates` before the min-subunit step,
    `_reassemble_complexes.py:52-57`), filtered rows filled with the worst kept score (`_liana_pipe.py:433-443`)."""
    from variants import variant_inputs
    adata, _ = variant_inputs()
    kw = dict(groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, perm_source="reference", use_raw=False,
              return_all_lrs=True)
    res = li.mt.cellphonedb(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_allrs_cellphonedb.csv.gz"), float_precision="round_trip")
    compare(res, gold, "lr_means", "cellphone_pvals")
    a, b = res.sort_values(KEY).reset_index(drop=True), gold.sort_values(KEY).reset_index(drop=True)
    assert a["lrs_to_keep"].dtype == bool and np.array_equal(a["lrs_to_keep"].values, b["lrs_to_keep"].values)
    res = li.mt.natmi(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_allrs_natmi.csv.gz"), float_precision="round_trip")
    assert list(res.columns) == list(gold.columns)
    a, b = res.sort_values(KEY).reset_index(drop=True), gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b) and (a[KEY + ["ligand", "receptor"]].values == b[KEY + ["ligand", "receptor"]].values).all()
    assert np.array_equal(a["lrs_to_keep"].values, b["lrs_to_keep"].values)
    for c in ("ligand_means", "receptor_means", "ligand_props", "receptor_props"):
        assert np.array_equal(a[c].values, b[c].values.astype(a[c].dtype)), c
    for c in ("ligand_means_sums", "receptor_means_sums", "expr_prod", "spec_weight"):
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-6, err_msg=c)
    res = li.mt.rank_aggregate(adata, **kw)
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_allrs_rank_aggregate.csv.gz"), float_precision="round_trip")
    compare_aggregate(res, gold, ["specificity_rank", "magnitude_rank"])


def test_return_all_lrs_in_the_other_entry_points():
    """The reference's own `return_all_lrs` tests (liana/tests/test_sc_methods.py:151-164, test_rank_aggregate.py:60-70)
    restated on the synthetic input: NATMI's fill rule (both of its scores are 'larger is better', so filtered rows get
    the minimum), `logfc.by_sample`, `rank_aggregate` without permutations (11 columns) and `rank_aggregate.by_sample`
    (14 columns) - every call returns every (cluster pair, interaction) row."""
    adata, P, seed, ep = load_input("small300")
    L = adata.obs["cluster"].nunique()
    kw = dict(groupby="cluster", use_raw=False, expr_prop=ep, inplace=False)
    kept = li.mt.natmi(adata, **kw)
    lr_all = li.mt.natmi(adata, return_all_lrs=True, **kw)
    n_all = len(lr_all)
    assert n_all % (L * L) == 0 and n_all > len(kept) and lr_all.shape[1] == kept.shape[1] + 1
    assert int(lr_all["lrs_to_keep"].sum()) == len(kept)
    rest = lr_all[~lr_all["lrs_to_keep"]]
    assert (rest[li.mt.natmi.magnitude] == lr_all[li.mt.natmi.magnitude].min()).all()
    assert (rest[li.mt.natmi.specificity] == lr_all[li.mt.natmi.specificity].min()).all()
    res = li.mt.rank_aggregate(adata, return_all_lrs=True, n_perms=None, **kw)
    assert res.shape == (n_all, 11) and "specificity_rank" not in res.columns
    res = li.mt.rank_aggregate(adata, return_all_lrs=True, n_perms=8, **kw)
    assert res.shape == (n_all, 13)
    assert res["magnitude_rank"].between(0, 1).all() and np.all(np.diff(res["magnitude_rank"].values) >= 0)
    _with_samples(adata)
    by = li.mt.logfc.by_sample(adata, "sample", return_all_lrs=True, **kw)
    assert by.columns[0] == "sample" and set(by["sample"]) == {"A", "B"}
    for s in ("A", "B"):                                  # per sample: all pairs x the interactions covered by ITS genes
        assert (by["sample"] == s).sum() % (L * L) == 0
    by = li.mt.rank_aggregate.by_sample(adata, "sample", return_all_lrs=True, n_perms=8, key_added="liana_by_sample",
                                        **{k: v for k, v in kw.items() if k != "inplace"})
    assert by is None and adata.uns["liana_by_sample"].shape[1] == 14 and "sample" in adata.uns["liana_by_sample"].columns


def test_nothing_expressed_gives_an_empty_frame():
    """`expr_prop` above 1: no interaction survives.  The reference returns an empty frame with the usual columns and
    dtypes for a single method (its consensus fails inside pandas: 'cannot set a frame with no defined index');
    here every method - and the consensus - returns the empty frame."""
    adata, P, seed, ep = load_input("small300")
    full = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=4, inplace=False)
    for method in (li.mt.cellphonedb, li.mt.geometric_mean, li.mt.cellchat, li.mt.natmi, li.mt.rank_aggregate):
        res = method(adata, groupby="cluster", use_raw=False, n_perms=4, expr_prop=1.01, inplace=False)
        assert len(res) == 0, method.method_name
        ok = method(adata, groupby="cluster", use_raw=False, n_perms=4, inplace=False)
        assert list(res.columns) == list(ok.columns), method.method_name
    res = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=4, expr_prop=1.01, inplace=False)
    for c in ("ligand_means", "receptor_means", "lr_means"):
        assert res[c].dtype == np.float32
    for c in ("ligand_props", "receptor_props", "cellphone_pvals"):
        assert res[c].dtype == np.float64
    assert res["source"].dtype == full["source"].dtype and res["ligand"].dtype == full["ligand"].dtype
    by = li.mt.cellphonedb.by_sample(_with_samples(adata), "sample", groupby="cluster", use_raw=False, n_perms=4,
                                     expr_prop=1.01, inplace=False)
    assert len(by) == 0 and by.columns[0] == "sample"


def _with_samples(adata):
    adata.obs["sample"] = pd.Categorical(np.where(np.arange(adata.shape[0]) % 2 == 0, "A", "B"))
    return adata


def test_return_all_lrs_contract():
    """`return_all_lrs=True` (liana/tests/test_sc_methods.py:166-190: every (pair, interaction) row comes back, the
    rows that fail expr_prop get the worst magnitude / specificity of the kept ones): the contract checked directly,
    next to the frame comparison of `test_return_all_lrs_frames_identical_to_reference`."""
    adata, P, seed, ep = load_input("small300")
    kept = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                             inplace=False, perm_source="reference")
    full = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep,
                             inplace=False, perm_source="reference", return_all_lrs=True)
    assert "lrs_to_keep" in full.columns and len(full) > len(kept)
    n_pairs = adata.obs["cluster"].nunique() ** 2
    assert len(full) % n_pairs == 0                                   # every interaction for every cluster pair
    assert int(full["lrs_to_keep"].sum()) == len(kept)
    rest = full[~full["lrs_to_keep"]]
    assert np.all(rest["lr_means"] == full.loc[full["lrs_to_keep"], "lr_means"].min())
    assert np.all(rest["cellphone_pvals"] == full.loc[full["lrs_to_keep"], "cellphone_pvals"].max())
    # the kept rows carry the same p-values as the filtered call (for simple, non-complex interactions the
    # resolved subunits are identical too)
    m = kept.merge(full[full["lrs_to_keep"]], on=KEY, suffixes=("", "_all"))
    assert len(m) == len(kept)
    simple = ~(m["ligand_complex"].str.contains("_") | m["receptor_complex"].str.contains("_"))
    assert np.array_equal(m.loc[simple, "cellphone_pvals"].values, m.loc[simple, "cellphone_pvals_all"].values)
    assert np.array_equal(m.loc[simple, "lr_means"].values, m.loc[simple, "lr_means_all"].values)


def test_devices_kwarg_equals_single_gpu():
    """Single-call multi-GPU: `devices=[0, 1, ...]` splits the permutations (or the samples of `by_sample`) over the GPUs
    of ONE process - the committed matrix is copied device to device - and must return exactly the one-GPU frame."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from liana_py_b200.testing._synth import make_synthetic_adata
    devices = list(range(min(torch.cuda.device_count(), 4)))
    adata = make_synthetic_adata(8000, 1200, 7, density=0.2, seed_data=15, seed_labels=16, n_samples=4)
    kw = dict(groupby="cluster", use_raw=False, n_perms=333, seed=5, inplace=False)
    for fn in (li.mt.cellphonedb, li.mt.cellchat, li.mt.rank_aggregate):
        pd.testing.assert_frame_equal(fn(adata, devices=devices, **kw), fn(adata, **kw))
    pd.testing.assert_frame_equal(li.mt.cellphonedb.by_sample(adata, "sample", devices=devices, **kw),
                                  li.mt.cellphonedb.by_sample(adata, "sample", **kw))


def test_consensus_containing_cellchat_reference_quirk():
    """A custom consensus [cellphonedb, cellchat, geometric_mean]: the reference's CellChat null scales the shared matrix
    in place (`_get_mean_perms.py:43-44`), so geometric_mean - which runs after it - is shuffled on X / mat_max against
    its unscaled observed scores.  `reference_quirks=True` reproduces the reference's frame (fixture from its own code);
    the default computes every null on the caller's matrix: geometric_mean's p-values are the stand-alone method's."""
    from variants import variant_inputs
    adata, _ = variant_inputs()
    kw = dict(groupby="cluster", n_perms=16, seed=3, inplace=False, perm_source="reference", use_raw=False)
    agg = li.mt.AggregateClass(li.mt.aggregate_meta, methods=[li.mt.cellphonedb, li.mt.cellchat, li.mt.geometric_mean])
    gold = pd.read_csv(os.path.join(GOLDEN, "variant_cellchat_consensus_rank_aggregate.csv.gz"), float_precision="round_trip")
    quirk = agg(adata, reference_quirks=True, **kw)
    assert list(quirk.columns) == list(gold.columns) and len(quirk) == len(gold)
    a, b = quirk.sort_values(KEY).reset_index(drop=True), gold.sort_values(KEY).reset_index(drop=True)
    for c in ("cellphone_pvals", "cellchat_pvals", "gmean_pvals"):
        assert np.array_equal(a[c].values, b[c].values), c
    for c in ("lr_means", "lr_probs", "lr_gmeans"):
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-6, atol=0)
    plain = agg(adata, **kw).sort_values(KEY).reset_index(drop=True)
    alone = li.mt.geometric_mean(adata, **kw).sort_values(KEY).reset_index(drop=True)
    assert np.array_equal(plain["cellphone_pvals"].values, b["cellphone_pvals"].values)
    assert np.array_equal(plain["gmean_pvals"].values, alone["gmean_pvals"].values)
    assert not np.array_equal(plain["gmean_pvals"].values, b["gmean_pvals"].values)


def test_overlapped_frame_assembly_equals_plain_path():
    """`Method.__call__` assembles the sorted frame while the permutations run (>= 4096 rows): same frame as
    `run_method` + `finish_frame`, column for column, index included."""
    from liana_py_b200.method import _methods as M, _pipeline as PL
    from liana_py_b200.testing._synth import make_synthetic_adata
    adata = make_synthetic_adata(400, 900, 6, density=0.3, seed_data=3, seed_labels=4)
    for fn in (li.mt.cellphonedb, li.mt.geometric_mean, li.mt.cellchat):
        kw = dict(groupby="cluster", use_raw=False, n_perms=5, seed=1, inplace=False)
        got = fn(adata, **kw)
        assert len(got) >= 4096
        add_cols = PL.method_columns(fn.add_cols, None)
        state = PL.prepare(adata, "cluster", "consensus", None, None, None, 5, False, None, False,
                           want_max="mat_max" in add_cols)
        plain = M.run_method(state, fn._method, 0.1, 5, 1, False, None, None, "product", None, add_cols=add_cols)
        plain = PL.finish_frame(state, plain, fn.magnitude, fn.magnitude_ascending)
        pd.testing.assert_frame_equal(got, plain)
