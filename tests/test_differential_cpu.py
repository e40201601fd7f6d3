"""CPU, differential: random small inputs and call options through the reference's own code (oracle/ref_shim.py, needs
/root/reference: skipped elsewhere) and through the drop-in layer with the oracle-backed engine stand-in
(tests/_oracle_engine.py).  Frames must agree exactly for the permutation methods (cellphonedb, geometric_mean,
cellchat) and to the reference's own tolerance for a non-permutation score - a fuzz test of the host logic (input
contract, cluster filtering, resource tables, complex resolution, score functions, frame assembly)."""
import numpy as np
import pandas as pd
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="needs the reference tree")

from liana_py_b200.method import _pipeline as PL          # noqa: E402
from liana_py_b200.testing._synth import make_synthetic_adata      # noqa: E402
from _oracle_engine import OracleEngine                   # noqa: E402
import liana_py_b200 as li                                # noqa: E402

KEY = ["source", "target", "ligand_complex", "receptor_complex"]


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setitem(PL._ENGINES, 0, OracleEngine())
    # the reference switches `specificity` of its shared method objects off for good after an `n_perms=None` call
    # (`_liana_pipe.py:420-422`); keep one test's calls from changing what the next one gets
    ref = ref_shim.load_reference()
    methods = [getattr(m, "_method", m) for m in (ref.cellphonedb, ref.geometric_mean, ref.cellchat)]   # the records that get mutated
    saved = [m.specificity for m in methods]
    yield
    for m, spec in zip(methods, saved):
        m.specificity = spec


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    n, g, L = int(rng.integers(60, 260)), int(rng.integers(150, 700)), int(rng.integers(2, 7))
    adata = make_synthetic_adata(n, g, L, density=float(rng.uniform(0.05, 0.3)), seed_data=seed, seed_labels=seed + 50)
    kw = dict(expr_prop=float(rng.choice([0.0, 0.05, 0.1, 0.3])), min_cells=int(rng.choice([1, 5, 12])),
              n_perms=int(rng.integers(3, 12)), seed=int(rng.integers(0, 10 ** 6)))
    if rng.random() < 0.4:                            # unbalanced clusters: some fall below min_cells
        codes = np.asarray(adata.obs["cluster"].cat.codes)
        drop = np.flatnonzero((codes == 0) & (rng.random(n) < 0.85))
        keep = np.setdiff1d(np.arange(n), drop)
        adata = adata[keep].copy()
    if rng.random() < 0.3:
        cats = list(adata.obs["cluster"].cat.categories)
        k = min(len(cats), 3)
        kw["groupby_pairs"] = pd.DataFrame({"source": rng.choice(cats, k), "target": rng.choice(cats, k)}).drop_duplicates()
    if rng.random() < 0.3:
        kw["resource_name"] = str(rng.choice(["cellphonedb", "cellchatdb", "connectomedb2020", "mouseconsensus"]))
    if rng.random() < 0.25:                           # unlabeled cells
        lab = adata.obs["cluster"].astype(object).copy()
        lab[rng.random(len(lab)) < 0.08] = np.nan
        adata.obs["cluster"] = pd.Categorical(lab)
    return adata, kw


def _same(a, b, exact, close=()):
    assert list(a.columns) == list(b.columns)
    a, b = a.sort_values(KEY).reset_index(drop=True), b.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in exact:
        assert np.array_equal(a[c].values, b[c].values.astype(a[c].dtype), equal_nan=True), c
    for c in close:
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-6, err_msg=c)


@pytest.mark.parametrize("seed", range(12))
def test_random_inputs_agree_with_the_reference(seed):
    ref = ref_shim.load_reference()
    adata, kw = _case(seed)
    common = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, **kw)
    try:
        want = ref.cellphonedb(adata.copy(), **common)
    except Exception as e:                             # e.g. too few resource genes in a tiny input: same refusal here
        with pytest.raises(type(e)):
            li.mt.cellphonedb(adata.copy(), perm_source="reference", **common)
        return
    got = li.mt.cellphonedb(adata.copy(), perm_source="reference", **common)
    stats = ["ligand_means", "receptor_means", "ligand_props", "receptor_props"]
    _same(got, want, stats + ["lr_means", "cellphone_pvals"])
    with np.errstate(divide="ignore", invalid="ignore"):
        want = ref.geometric_mean_mod.geometric_mean(adata.copy(), **common)
        got = li.mt.geometric_mean(adata.copy(), perm_source="reference", **common)
    _same(got, want, stats + ["lr_gmeans", "gmean_pvals"])
    want = ref.cellchat(adata.copy(), **common)
    got = li.mt.cellchat(adata.copy(), perm_source="reference", **common)
    _same(got, want, ["ligand_trimean", "receptor_trimean", "ligand_props", "receptor_props", "lr_probs", "cellchat_pvals"])
    want = ref.natmi_mod.natmi(adata.copy(), **common)
    got = li.mt.natmi(adata.copy(), **common)
    _same(got, want, stats, close=["expr_prod", "spec_weight"])


@pytest.mark.parametrize("seed", range(20, 26))
def test_random_inputs_non_permutation_scores(seed):
    """Connectome, log2FC, SingleCellSignalR, scSeqComm on random inputs (a random log base for the fold changes): key
    columns and resolved subunits identical, scores within the reference's own frame tolerance (`float32` statistics
    accumulated by another route), scSeqComm's normal cdf to 1e-5 (its cluster std is a float32 `np.std` there)."""
    ref = ref_shim.load_reference()
    adata, kw = _case(seed)
    kw.pop("n_perms"); kw.pop("seed")
    base = float(np.random.default_rng(seed).choice([np.e, 2.0, 10.0]))
    common = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, base=base, **kw)
    try:
        ref.natmi_mod.natmi(adata.copy(), **common)
    except ValueError:
        pytest.skip("input refused by the reference (covered by the other test)")
    for want_fn, got_fn, close, tol in (
            (ref.connectome_mod.connectome, li.mt.connectome, ["expr_prod", "scaled_weight"], None),
            (ref.logfc_mod.logfc, li.mt.logfc, ["lr_logfc"], None),
            (ref.sca_mod.singlecellsignalr, li.mt.singlecellsignalr, ["lrscore"], None),
            (ref.scseqcomm, li.mt.scseqcomm, [], 1e-5)):
        want, got = want_fn(adata.copy(), **common), got_fn(adata.copy(), **common)
        _same(got, want, [], close=close)
        if tol is not None:
            a, b = got.sort_values(KEY).reset_index(drop=True), want.sort_values(KEY).reset_index(drop=True)
            np.testing.assert_allclose(a["inter_score"].values, b["inter_score"].values, rtol=0, atol=tol)


MODES = ["return_all_lrs", "supp_columns", "dense", "interactions", "by_sample", "raw_layer"]


@pytest.mark.parametrize("seed", range(40, 52))
def test_random_inputs_call_options(seed):
    """One optional feature of the call signature per case, on a random input, cellphonedb on both sides: every
    (pair, interaction) row (`return_all_lrs`), supplementary columns, a dense matrix, a hand-picked interaction list,
    `by_sample`, `.raw` / a layer."""
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(seed)
    adata, kw = _case(seed)
    kw.pop("groupby_pairs", None)
    mode = MODES[seed % len(MODES)]
    common = dict(groupby="cluster", inplace=False, verbose=False, **kw)
    use = dict(use_raw=False)
    by_sample = False
    close = []
    if mode == "return_all_lrs":
        common["return_all_lrs"] = True
    elif mode == "supp_columns":
        pool = ["ligand_means_sums", "receptor_means_sums", "ligand_zscores", "receptor_zscores", "ligand_logfc", "receptor_logfc",
                "mat_mean", "ligand_cdf", "receptor_cdf"]
        common["supp_columns"] = close = list(rng.choice(pool, size=int(rng.integers(1, 5)), replace=False))
    elif mode == "dense":
        adata.X = adata.X.toarray().astype(rng.choice([np.float32, np.float64]))
    elif mode == "interactions":
        from liana_py_b200.resource import select_resource
        res = select_resource(common.pop("resource_name", "consensus"))
        present = set(adata.var_names)
        ok = res[[all(x in present for x in (l + "_" + r).split("_")) for l, r in zip(res["ligand"], res["receptor"])]]
        if len(ok) < 5:
            pytest.skip("too few covered interactions in this random input")
        pick = ok.iloc[rng.choice(len(ok), size=min(len(ok), 25), replace=False)]
        common["interactions"] = list(zip(pick["ligand"], pick["receptor"]))
    elif mode == "by_sample":
        by_sample = True
        adata.obs["sample"] = rng.choice(["a", "b", "c"], size=adata.shape[0])
    elif mode == "raw_layer":
        X = adata.X.tocsr()
        other = X.copy(); other.data = np.sqrt(other.data).astype(np.float32)
        if rng.random() < 0.5:
            adata.raw = type("Raw", (), {"X": other, "var_names": adata.var_names, "var": adata.var})()
            use = dict(use_raw=True)
        else:
            adata.layers["alt"] = other
            use = dict(use_raw=False, layer="alt")
    common.update(use)

    def call(method, a, **extra):
        return method.by_sample(a, "sample", **common, **extra) if by_sample else method(a, **common, **extra)

    try:
        want = call(ref.cellphonedb, adata.copy())
    except ValueError:
        with pytest.raises(ValueError):
            call(li.mt.cellphonedb, adata.copy(), perm_source="reference")
        return
    got = call(li.mt.cellphonedb, adata.copy(), perm_source="reference")
    key = (["sample"] if by_sample else []) + KEY
    assert list(got.columns) == list(want.columns)
    a, b = got.sort_values(key).reset_index(drop=True), want.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b)
    for c in key + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), (mode, c)
    for c in ["ligand_means", "receptor_means", "ligand_props", "receptor_props", "lr_means", "cellphone_pvals"]:
        assert np.array_equal(a[c].values, b[c].values.astype(a[c].dtype)), (mode, c)
    if "lrs_to_keep" in b:
        assert np.array_equal(a["lrs_to_keep"].values, b["lrs_to_keep"].values)
    for c in close:
        tol = dict(rtol=0, atol=1e-5) if c.endswith("_cdf") else dict(rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(a[c].values.astype(np.float64), b[c].values.astype(np.float64), err_msg=f"{mode} {c}", **tol)


@pytest.mark.parametrize("seed", [60, 61, 62, 64, 65, 72, 75, 78])
def test_random_inputs_consensus(seed):
    """`rank_aggregate` (RRA or mean rank, with and without `return_all_lrs`) on random inputs: every method's score
    column within the reference's frame tolerance, p-values identical; the two consensus columns agree except where
    the reference orders mathematically tied float32 scores by rounding noise (<= 5 % of the rows, |diff| <= 5e-3,
    see tests/test_dropin_gpu.py::test_consensus_options_identical_to_reference)."""
    ref = ref_shim.load_reference()
    adata, kw = _case(seed)
    rng = np.random.default_rng(seed)
    extra = {}
    if rng.random() < 0.3:
        extra["aggregate_method"] = "mean"
    if rng.random() < 0.3:
        extra["return_all_lrs"] = True
    common = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, **kw, **extra)
    want = ref.rank_aggregate(adata.copy(), **common)
    got = li.mt.rank_aggregate(adata.copy(), perm_source="reference", **common)
    assert list(got.columns) == list(want.columns)
    a, b = got.sort_values(KEY).reset_index(drop=True), want.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b) and (a[KEY].astype(str).values == b[KEY].astype(str).values).all()
    assert np.array_equal(a["cellphone_pvals"].values, b["cellphone_pvals"].values)
    for c in ("lr_means", "expr_prod", "scaled_weight", "lr_logfc", "spec_weight", "lrscore"):
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-6, err_msg=c)
    for c in ("specificity_rank", "magnitude_rank"):
        x, y = a[c].values, b[c].values
        assert np.array_equal(np.isnan(x), np.isnan(y)), c          # a NaN score makes every rank NaN (nan_policy='propagate')
        ok = ~np.isnan(y)
        if ok.any():
            d = np.abs(x[ok] - y[ok])
            assert (d <= 1e-4 * np.maximum(np.abs(y[ok]), 1e-12)).mean() >= 0.95 and d.max() <= 5e-3, (c, float(d.max()))


def test_degenerate_inputs_and_error_conventions_agree():
    """Degenerate inputs and wrong arguments: both sides return the same frame, or refuse with the same exception type
    (and, for the reference's own messages, the same text)."""
    ref = ref_shim.load_reference()

    def base():
        return make_synthetic_adata(150, 400, 4, density=0.15, seed_data=5, seed_labels=6)

    def both(adata, ref_fn=None, our_fn=None, **kw):
        ref_fn, our_fn = ref_fn or ref.cellphonedb, our_fn or li.mt.cellphonedb
        common = dict(groupby="cluster", inplace=False, verbose=False, **kw)
        extra = dict(perm_source="reference") if our_fn in (li.mt.cellphonedb, li.mt.geometric_mean, li.mt.cellchat) else {}
        try:
            want = ref_fn(adata.copy(), **common)
        except Exception as e:
            with pytest.raises(type(e)) as info:
                our_fn(adata.copy(), **common, **extra)
            return str(e), str(info.value)
        got = our_fn(adata.copy(), **common, **extra)
        assert list(got.columns) == list(want.columns) and len(got) == len(want)
        a, b = got.sort_values(KEY).reset_index(drop=True), want.sort_values(KEY).reset_index(drop=True)
        for c in a.columns:
            if a[c].dtype.kind == "f":
                np.testing.assert_allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=1e-4, atol=2e-6, err_msg=c)
            else:
                assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
        return None

    a = base(); a.obs["cluster"] = pd.Categorical(["only"] * 150)
    both(a, use_raw=False, n_perms=4)                                                        # a single cluster
    both(base(), use_raw=False, n_perms=4, groupby_pairs=pd.DataFrame({"source": ["c00", "zzz"], "target": ["c01", "c00"]}))
    a = base(); a.obs["cluster"] = np.asarray(a.obs["cluster"].cat.codes)
    both(a, use_raw=False, n_perms=4)                                                        # integer, non-categorical labels
    a = base(); names = np.array(a.var_names, dtype=object); names[5] = names[6]; a.var_names = list(names)
    both(a, use_raw=False, n_perms=4)                                                        # duplicated gene name
    both(base(), use_raw=False, n_perms=4, expr_prop=1.0)
    a = base(); a.X = a.X.tocsc()
    both(a, use_raw=False, n_perms=4)
    a = base(); a.X = a.X.astype(np.float64)
    both(a, use_raw=False, n_perms=4)
    a = base(); X = a.X.tolil(); X[3] = 0; X[:, 7] = 0; a.X = X.tocsr()
    both(a, use_raw=False, n_perms=4)                                                        # an empty cell, an empty gene
    both(base(), ref.connectome_mod.connectome, li.mt.connectome, use_raw=False, n_perms=4)
    both(base(), ref.logfc_mod.logfc, li.mt.logfc, use_raw=False, base=10)
    both(base(), ref.scseqcomm, li.mt.scseqcomm, use_raw=False, expr_prop=0)
    both(base(), ref.cellchat, li.mt.cellchat, use_raw=False, n_perms=None)
    both(base(), ref.geometric_mean, li.mt.geometric_mean, use_raw=False, n_perms=None)
    # refusals: same type, same words
    a = base(); X = a.X.copy(); X.data[5] = np.nan; a.X = X
    for msgs in (both(a, use_raw=False, n_perms=2), both(base(), n_perms=2), both(base(), n_perms=2, layer="x", use_raw=True),
                 both(base(), use_raw=False, n_perms=2, interactions=[("a", "b", "c")]),
                 both(base(), use_raw=False, n_perms=2, groupby_pairs=pd.DataFrame({"x": [1]}))):
        assert msgs is not None and msgs[0] == msgs[1], msgs
    assert both(base(), use_raw=False, n_perms=2, layer="x") is not None                     # KeyError: unknown layer


@pytest.mark.parametrize("combo", ["cpdb+gmean+natmi", "gmean+sca+scseqcomm", "connectome+logfc"])
def test_custom_consensus_agrees(combo):
    """`AggregateClass(aggregate_meta, methods=[...])` with method lists other than the default (two permutation nulls
    counted in one pass; a consensus without any permutation method): columns, scores and both rank columns."""
    ref = ref_shim.load_reference()
    RA = ref.rank_aggregate_mod
    rm, om = {
        "cpdb+gmean+natmi": ([ref.cellphonedb, ref.geometric_mean, ref.natmi_mod.natmi], [li.mt.cellphonedb, li.mt.geometric_mean, li.mt.natmi]),
        "gmean+sca+scseqcomm": ([ref.geometric_mean, ref.sca_mod.singlecellsignalr, ref.scseqcomm],
                                [li.mt.geometric_mean, li.mt.singlecellsignalr, li.mt.scseqcomm]),
        "connectome+logfc": ([ref.connectome_mod.connectome, ref.logfc_mod.logfc], [li.mt.connectome, li.mt.logfc]),
    }[combo]
    adata = make_synthetic_adata(150, 400, 4, density=0.15, seed_data=5, seed_labels=6)
    kw = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, n_perms=6, seed=2)
    want = RA.AggregateClass(RA._rank_aggregate_meta, methods=rm)(adata.copy(), **kw)
    got = li.mt.AggregateClass(li.mt.aggregate_meta, methods=om)(adata.copy(), perm_source="reference", **kw)
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    a, b = got.sort_values(KEY).reset_index(drop=True), want.sort_values(KEY).reset_index(drop=True)
    assert (a[KEY].astype(str).values == b[KEY].astype(str).values).all()
    for c in want.columns[4:]:
        tol = dict(rtol=0, atol=1e-5) if c == "inter_score" else dict(rtol=1e-4, atol=2e-6)
        if c.endswith("_pvals"):
            assert np.array_equal(a[c].values, b[c].values), c
        elif c.endswith("_rank"):
            d = np.abs(a[c].values - b[c].values)
            assert (d <= 1e-4 * np.maximum(np.abs(b[c].values), 1e-12)).mean() >= 0.95 and d.max() <= 5e-3, (c, float(d.max()))
        else:
            np.testing.assert_allclose(a[c].values.astype(float), b[c].values.astype(float), err_msg=c, **tol)


@pytest.mark.parametrize("seed", range(90, 94))
def test_random_mudata_inputs(seed):
    """A random input split into two modalities at a random column, with a random choice of `x_layer` / `y_use_raw` /
    transforms: the reference's `mdata_to_anndata` + pipeline against `mdata_kwargs` here."""
    from liana_py_b200.testing._anndata_lite import MuDataLite
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(seed)
    adata, kw = _case(seed)
    kw.pop("groupby_pairs", None)
    cut = int(rng.integers(adata.shape[1] // 4, 3 * adata.shape[1] // 4))
    x, y = adata[:, :cut].copy(), adata[:, cut:].copy()
    mkw = dict(x_mod="left", y_mod="right")
    if rng.random() < 0.5:
        lay = x.X.copy(); lay.data = np.log1p(lay.data).astype(np.float32)
        x.layers["log"] = lay
        mkw["x_layer"] = "log"
    if rng.random() < 0.5:
        mkw["y_transform"] = np.sqrt
    if rng.random() < 0.5:
        raw = y.X.copy(); raw.data = (raw.data * 2).astype(np.float32)
        y.raw = type("Raw", (), {"X": raw, "var_names": y.var_names, "var": y.var,
                                 "to_adata": lambda self=None, _y=y, _r=raw: type(_y)(X=_r, obs=_y.obs.copy(), var=_y.var.copy())})()
        mkw["y_use_raw"] = True
    mdata = MuDataLite({"left": x, "right": y}, obs=adata.obs.copy())
    common = dict(groupby="cluster", inplace=False, verbose=False, mdata_kwargs=mkw, **kw)
    try:
        want = ref.cellphonedb(mdata, **common)
    except ValueError:
        with pytest.raises(ValueError):
            li.mt.cellphonedb(mdata, perm_source="reference", **common)
        return
    got = li.mt.cellphonedb(mdata, perm_source="reference", **common)
    _same(got, want, ["ligand_means", "receptor_means", "ligand_props", "receptor_props", "lr_means", "cellphone_pvals"])


@pytest.mark.parametrize("name,supp", [
    ("cellchat", ["ligand_means", "receptor_means", "mat_mean"]),
    ("natmi", ["ligand_zscores", "receptor_logfc", "mat_max", "ligand_trimean"]),
    ("scseqcomm", ["ligand_means_sums", "mat_mean"]),
    ("logfc", ["ligand_cdf", "receptor_cdf", "ligand_props"]),
    ("cellphonedb", ["nonsense"]),                      # KeyError on both sides
    ("cellphonedb", ["ligand_trimean"]),                # trimeans exist only with mat_max: KeyError on both sides
])
def test_supp_columns_of_every_method_agree(name, supp):
    ref = ref_shim.load_reference()
    ref_fn = {"cellchat": ref.cellchat, "natmi": ref.natmi_mod.natmi, "scseqcomm": ref.scseqcomm, "logfc": ref.logfc_mod.logfc,
              "cellphonedb": ref.cellphonedb}[name]
    our_fn = getattr(li.mt, name)
    adata = make_synthetic_adata(150, 400, 4, density=0.15, seed_data=5, seed_labels=6)
    kw = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, supp_columns=supp)
    perm = dict(n_perms=4, seed=3) if name in ("cellchat", "cellphonedb") else {}
    extra = dict(perm_source="reference") if perm else {}
    try:
        want = ref_fn(adata.copy(), **kw, **perm)
    except KeyError:
        with pytest.raises(KeyError):
            our_fn(adata.copy(), **kw, **perm, **extra)
        return
    got = our_fn(adata.copy(), **kw, **perm, **extra)
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    a, b = got.sort_values(KEY).reset_index(drop=True), want.sort_values(KEY).reset_index(drop=True)
    for c in want.columns:
        if a[c].dtype.kind == "f":
            np.testing.assert_allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=1e-4, atol=1e-5, err_msg=c)
        else:
            assert (a[c].astype(str).values == b[c].astype(str).values).all(), c


@pytest.mark.parametrize("seed", [3, 5, 11])
def test_unsorted_row_order_is_the_references(seed):
    """Per-method frames of a consensus returned as they are (`consensus_opts=False`, `_liana_pipe.py:220-221`): the rows
    come out in the reference's own order - per (source, target) pair of `np.meshgrid(labels, labels).reshape(2, L*L).T`
    (`:310-312`, the SOURCE varies fastest), interactions inside - not merely the same set of rows."""
    ref = ref_shim.load_reference()
    adata, kw = _case(seed)
    kw.pop("groupby_pairs", None)
    agg_ref = ref.rank_aggregate_mod.AggregateClass(ref.rank_aggregate_mod._rank_aggregate_meta,
                                                    methods=[ref.cellphonedb, ref.geometric_mean])
    want = agg_ref(adata.copy(), groupby="cluster", use_raw=False, consensus_opts=False, inplace=False, **kw)
    agg = li.mt.AggregateClass(li.mt.aggregate_meta, methods=[li.mt.cellphonedb, li.mt.geometric_mean])
    got = agg(adata.copy(), groupby="cluster", use_raw=False, consensus_opts=False, inplace=False, perm_source="reference", **kw)
    assert list(got) == list(want)
    for name in got:
        a, b = got[name].reset_index(drop=True), want[name].reset_index(drop=True)
        assert len(a) == len(b) > 0
        assert (a[KEY].values == b[KEY].values).all()
        for col in b.columns:
            if col not in KEY:
                np.testing.assert_array_equal(a[col].values, b[col].values)


def test_cellchat_in_a_consensus_matrix_carry_over():
    """The reference's CellChat null divides the shared matrix in place (`adata.X /= norm_factor`,
    `_get_mean_perms.py:43-44`): in a consensus, the permutation methods that run AFTER CellChat shuffle X / mat_max
    against their unscaled observed scores.  The drop-in computes every null on the matrix the caller passed (default,
    pinned here: p-values equal the stand-alone method's) and reproduces the reference's behaviour with
    `reference_quirks=True` (p-values identical to the reference's consensus)."""
    ref = ref_shim.load_reference()
    RA = ref.rank_aggregate_mod
    adata = make_synthetic_adata(160, 420, 4, density=0.2, seed_data=15, seed_labels=16)
    kw = dict(groupby="cluster", use_raw=False, inplace=False, verbose=False, n_perms=8, seed=3)
    ref_methods = [ref.cellphonedb, ref.cellchat, ref.geometric_mean]
    own_methods = [li.mt.cellphonedb, li.mt.cellchat, li.mt.geometric_mean]
    want = RA.AggregateClass(RA._rank_aggregate_meta, methods=ref_methods)(adata.copy(), **kw).sort_values(KEY).reset_index(drop=True)
    agg = li.mt.AggregateClass(li.mt.aggregate_meta, methods=own_methods)
    quirk = agg(adata.copy(), perm_source="reference", reference_quirks=True, **kw).sort_values(KEY).reset_index(drop=True)
    plain = agg(adata.copy(), perm_source="reference", **kw).sort_values(KEY).reset_index(drop=True)
    alone = li.mt.geometric_mean(adata.copy(), perm_source="reference", **kw).sort_values(KEY).reset_index(drop=True)
    assert list(quirk.columns) == list(want.columns) and len(quirk) == len(want)
    for c in ("cellphone_pvals", "cellchat_pvals", "gmean_pvals"):
        assert np.array_equal(quirk[c].values, want[c].values), c            # the reference's consensus, reproduced
    # before CellChat nothing differs; after it the reference's null is computed on the scaled matrix
    assert np.array_equal(plain["cellphone_pvals"].values, want["cellphone_pvals"].values)
    assert np.array_equal(plain["cellchat_pvals"].values, want["cellchat_pvals"].values)
    assert np.array_equal(plain["gmean_pvals"].values, alone["gmean_pvals"].values)   # default: the stand-alone method's p-values
    assert not np.array_equal(plain["gmean_pvals"].values, want["gmean_pvals"].values)
