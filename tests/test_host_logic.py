"""CPU: the host logic of the drop-in's input contract (`method/_pre.py`, restating `prep_check_adata`,
liana/method/_pipe_utils/_pre.py:65-169) with the engine stood in by SciPy (the three matrix statistics the device
computes from the staged upload), checked against the reference's own `prep_check_adata` where /root/reference exists
and against first principles everywhere."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from liana_py_b200.method import _pre
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import make_synthetic_adata


def count_unordered(X) -> int:
    """Stored entries whose column is not greater than its predecessor's in the same row (what `stage_stats_kernel` counts)."""
    ix, ip = np.asarray(X.indices, dtype=np.int64), np.asarray(X.indptr, dtype=np.int64)
    if len(ix) < 2:
        return 0
    bad = np.diff(ix) <= 0                                   # bad[o - 1]: entry o against entry o - 1
    starts = ip[:-1][(ip[:-1] > 0) & (ip[:-1] < len(ix)) & (np.diff(ip) > 0)]
    bad[starts - 1] = False                                  # the first entry of a row has no predecessor
    return int(bad.sum())


class SciPyStage:
    """Stand-in of Engine.stage_matrix / staged_max / staged_cluster_stats (what the CUDA kernels compute)."""

    def stage_matrix(self, X, row_keep=None):
        X = X.tocsr()                 # a SciPy matrix, or the row-block view of multi-rank staging (`_pre.CSRBlock`)
        self.X, self.keep = X, row_keep
        rows = np.arange(X.shape[0]) if row_keep is None else np.flatnonzero(row_keep)
        return dict(col_sums=np.asarray(X.astype(np.float64).sum(axis=0)).ravel(),
                    kept_total=float(X[rows].astype(np.float64).sum()), n_nonfinite=int((~np.isfinite(X.data)).sum()),
                    n_empty_rows=int((np.asarray(X.sum(axis=1)).ravel() == 0).sum()), n_unordered=count_unordered(X))

    def stage_matrix_dense(self, X, row_keep=None):
        self.dense_calls = getattr(self, "dense_calls", 0) + 1
        return self.stage_matrix(sparse.csr_matrix(X).astype(np.float32), row_keep)

    def staged_max(self):
        rows = np.arange(self.X.shape[0]) if self.keep is None else np.flatnonzero(self.keep)
        sub = self.X[rows]
        return (np.float32(sub.data.max()) if sub.nnz else np.float32(-np.inf)), int(sub.nnz)

    def staged_cluster_stats(self, row_label, L):
        sx, sq = np.zeros(L), np.zeros(L)
        for c in range(L):
            d = self.X[np.flatnonzero(row_label == c)].data.astype(np.float64)
            sx[c], sq[c] = d.sum(), (d * d).sum()
        return sx, sq


def _ragged():
    rng = np.random.default_rng(3)
    n, g = 120, 40
    dense = (rng.random((n, g)) < 0.3) * rng.random((n, g))
    dense[:, 7] = 0                                   # empty gene -> removed
    dense[11] = 0                                     # empty cell -> kept
    labels = np.array(["b"] * 50 + ["a"] * 40 + ["tiny"] * 3 + ["c"] * 27)
    obs = pd.DataFrame({"cluster": pd.Categorical(labels, categories=["c", "a", "b", "tiny", "unused"])})
    var = pd.DataFrame(index=pd.Index([f"g{(i * 7) % g:02d}" for i in range(g)]))     # unsorted names
    return AnnDataLite(X=sparse.csr_matrix(dense.astype(np.float32)), obs=obs, var=var)


def test_prepare_input_contract():
    ad = _ragged()
    prep = _pre.prepare_input(ad, "cluster", min_cells=5, engine=SciPyStage(), want_max=True, want_cluster_stats=True)
    assert list(prep.categories) == ["c", "a", "b"]                        # category order, 'tiny' (< min_cells) and 'unused' dropped
    assert prep.n_cells == 117 and len(prep.labels) == 117 and prep.labels.max() == 2
    assert list(prep.var_names) == sorted(prep.var_names) and ad.var_names[7] not in prep.var_names and len(prep.var_names) == 39
    kept = np.flatnonzero(np.asarray(ad.obs["cluster"]) != "tiny")
    X = ad.X[kept]
    assert np.isclose(prep.mat_mean, X.sum() / (117 * 39))
    assert prep.mat_max == np.float32(X.max())
    codes = prep.labels
    for c in range(3):
        rows = X[codes == c].toarray()[:, np.asarray(ad.X.sum(axis=0)).ravel() != 0].astype(np.float64)
        assert np.isclose(prep.cluster_stats[0][c], rows.mean(), rtol=1e-12)
        assert np.isclose(prep.cluster_stats[1][c], rows.std(), rtol=1e-9)
    # groupby_subset keeps only the named clusters
    sub = _pre.prepare_input(ad, "cluster", min_cells=5, engine=SciPyStage(), groupby_subset=["a", "c"])
    assert list(sub.categories) == ["c", "a"] and sub.n_cells == 67


def test_prepare_input_errors_match_reference_conventions():
    ad = _ragged()
    with pytest.raises(AssertionError):
        _pre.prepare_input(ad, "nope", min_cells=5, engine=SciPyStage())
    with pytest.raises(ValueError, match="raw"):
        _pre.prepare_input(ad, "cluster", min_cells=5, engine=SciPyStage(), use_raw=True)
    with pytest.raises(ValueError, match="layer"):
        _pre.prepare_input(ad, "cluster", min_cells=5, engine=SciPyStage(), use_raw=True, layer="x")
    bad = ad.copy()
    bad.X = bad.X.copy()
    bad.X.data[3] = np.nan
    with pytest.raises(ValueError, match="non finite"):
        _pre.prepare_input(bad, "cluster", min_cells=5, engine=SciPyStage())
    with pytest.raises(ValueError, match="Too few features"):
        _pre.assert_covered(np.array(["A", "B", "C"]), np.array(["X", "Y"]))
    _pre.assert_covered(np.array(["A", "B"]), np.array(["A", "Y"]))       # half missing is allowed (0.98)


@pytest.mark.skipif(not os.path.isdir("/root/reference/liana"), reason="needs the reference tree (build container)")
def test_prepare_input_equals_reference_prep_check_adata():
    from oracle.ref_shim import load_reference
    ref = load_reference()
    for ad in (_ragged(), make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12)):
        out = ref.pre.prep_check_adata(adata=ad.copy(), groupby="cluster", min_cells=5, use_raw=False, layer=None, verbose=False)
        prep = _pre.prepare_input(ad.copy(), "cluster", min_cells=5, engine=SciPyStage())
        assert list(out.var_names) == list(prep.var_names)
        assert list(out.obs["@label"].cat.categories) == list(prep.categories)
        assert np.array_equal(np.asarray(out.obs["@label"].cat.codes), prep.labels)
        assert out.shape == (prep.n_cells, len(prep.var_names))
        assert np.isclose(float(np.mean(out.X, dtype="float32")), prep.mat_mean, rtol=1e-5)


# ---- MuData input ------------------------------------------------------------------------------------------
def test_mdata_to_anndata_matches_reference():
    """`_pre.mdata_to_anndata` against the reference's `liana/utils/mdata_to_anndata.py` run on the same MuData
    look-alike (when /root/reference is present), and its error conventions."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from variants import mdata_inputs
    from liana_py_b200.method import _pre
    mdata, mkw = mdata_inputs()
    got = _pre.mdata_to_anndata(mdata, **mkw, verbose=False)
    x, y = mdata.mod["rna"], mdata.mod["prot"]
    assert got.X.shape == (x.shape[0], x.shape[1] + y.shape[1]) and sparse.isspmatrix_csr(got.X)
    assert list(got.var_names) == list(x.var_names) + list(y.var_names)
    assert np.array_equal(got.X[:, :x.shape[1]].toarray(), x.layers["sqrt"].toarray())
    assert np.array_equal(got.X[:, x.shape[1]:].toarray(), np.sqrt(y.X.toarray()))
    assert got.obs.equals(mdata.obs) and got.obs is not mdata.obs
    from oracle import ref_shim
    if ref_shim.reference_available():
        ref_shim.load_reference()
        import importlib
        ref_fun = importlib.import_module("liana.utils.mdata_to_anndata").mdata_to_anndata
        want = ref_fun(mdata, **mkw, verbose=False)
        assert list(want.var_names) == list(got.var_names)
        assert (want.X != got.X).nnz == 0 and want.obs.equals(got.obs)
    # cells in a different order / a subset in one modality: inner join on the names, x's order
    y2 = y[np.arange(y.shape[0])[::-1][:250]].copy()
    from liana_py_b200.testing._anndata_lite import MuDataLite
    x2 = x[np.isin(x.obs_names, y2.obs_names)].copy()
    md2 = MuDataLite({"rna": x, "prot": y2}, obs=x2.obs.copy())
    got2 = _pre.mdata_to_anndata(md2, x_mod="rna", y_mod="prot", verbose=False)
    assert got2.X.shape[0] == 250
    assert np.array_equal(got2.X[:, x.shape[1]:].toarray(), y[np.isin(y.obs_names, y2.obs_names)].X.toarray())
    with pytest.raises(ValueError, match="must be provided"):
        _pre.mdata_to_anndata(mdata, x_mod=None, y_mod="prot")
    with pytest.raises(ValueError, match="is not in the mdata"):
        _pre.mdata_to_anndata(mdata, x_mod="rna", y_mod="nope")
    raw = _pre.mdata_to_anndata(mdata, x_mod="rna", y_mod="prot", y_use_raw=True)      # `.raw` keeps ALL its variables
    assert raw.X.shape[1] == x.shape[1] + y.raw.X.shape[1]
    y.raw = None
    with pytest.raises(ValueError, match="raw"):
        _pre.mdata_to_anndata(mdata, x_mod="rna", y_mod="prot", y_use_raw=True)
    with pytest.raises(ValueError):                              # obs of the MuData does not match the joined cells
        _pre.mdata_to_anndata(MuDataLite({"rna": x, "prot": y2}), x_mod="rna", y_mod="prot")


def test_prepare_input_dense_matrix_goes_to_the_device_conversion():
    """A dense `adata.X` (2-D array) is not converted on the host: `prepare_input` hands it to
    `Engine.stage_matrix_dense` and ends with the same prepared input as for its CSR form."""
    adata = _ragged()
    want = _pre.prepare_input(adata, "cluster", 5, SciPyStage(), want_max=True, want_cluster_stats=True)
    for dtype in (np.float32, np.float64, np.int64):
        d = AnnDataLite(X=sparse.csr_matrix(adata.X), obs=adata.obs.copy(), var=adata.var.copy())
        d.X = (np.rint(adata.X.toarray()) if dtype == np.int64 else adata.X.toarray()).astype(dtype)
        stage = SciPyStage()
        got = _pre.prepare_input(d, "cluster", 5, stage, want_max=True, want_cluster_stats=True, verbose=True)
        assert stage.dense_calls == 1
        if dtype != np.int64:
            assert list(got.var_names) == list(want.var_names) and np.array_equal(got.var_cols, want.var_cols)
            assert np.array_equal(got.labels, want.labels) and got.mat_max == want.mat_max
            assert got.mat_mean == want.mat_mean and np.allclose(got.cluster_stats[1], want.cluster_stats[1])


def test_resource_tables_vectorised_equal_plain_restatement():
    """`_tables.tables_for_genes` (array operations on the once-exploded resource, what `by_sample` runs per sample)
    against a plain loop over the interactions (`filter_resource`, liana/method/_pipe_utils/_pre.py:188-224, +
    explode_complexes): kept interactions, their order, subunit indices, padding widths, entities - for several
    bundled resources and gene sets from complete to empty."""
    from liana_py_b200.method import _tables
    from liana_py_b200.resource import select_resource

    def plain(resource, var_names, sep="_"):
        present = set(map(str, var_names))
        L, R, lc, rc = [], [], [], []
        for lig, rec in zip(resource["ligand"].astype(str), resource["receptor"].astype(str)):
            ls, rs = lig.split(sep), rec.split(sep)
            if all(x in present for x in ls) and all(x in present for x in rs):
                L.append(ls); R.append(rs); lc.append(lig); rc.append(rec)
        ents = sorted({g for l in L for g in l} | {g for l in R for g in l})
        pos = {g: i for i, g in enumerate(ents)}
        a = np.full((len(lc), max((len(x) for x in L), default=1)), -1, np.int32)
        b = np.full((len(lc), max((len(x) for x in R), default=1)), -1, np.int32)
        for i, (ls, rs) in enumerate(zip(L, R)):
            a[i, :len(ls)] = [pos[g] for g in ls]
            b[i, :len(rs)] = [pos[g] for g in rs]
        return np.array(lc, dtype=str), np.array(rc, dtype=str), a, b, np.array(ents, dtype=str)

    rng = np.random.default_rng(0)
    for name in ("consensus", "cellphonedb", "cellchatdb", "mouseconsensus", "kirouac2010"):
        res = select_resource(name)
        ex = _tables.explode_resource(res)
        assert np.array_equal(ex.entities, _tables.resource_entities(res)) and len(ex.ligand_complex) == len(res)
        for frac in (1.0, 0.7, 0.2, 0.0):
            sub = ex.entities[rng.random(len(ex.entities)) < frac] if frac < 1 else ex.entities
            names = np.concatenate([sub, np.array(["BG1", "BG2"])])
            rng.shuffle(names)
            t, w = _tables.tables_for_genes(ex, names), plain(res, names)
            assert np.array_equal(t.ligand_complex, w[0]) and np.array_equal(t.receptor_complex, w[1]), (name, frac)
            assert t.lig_sub.shape == w[2].shape and np.array_equal(t.lig_sub, w[2]), (name, frac)
            assert t.rec_sub.shape == w[3].shape and np.array_equal(t.rec_sub, w[3]), (name, frac)
            assert np.array_equal(t.entities, w[4]), (name, frac)
    # the parsed resource is cached: callers get independent copies
    a = select_resource("consensus"); a.loc[0, "ligand"] = "changed"
    assert select_resource("consensus").loc[0, "ligand"] != "changed"


def test_handle_resource_precedence_and_errors():
    """Mirrors liana/tests/test_select_resource.py: interactions > resource > resource_name; sizes of the bundled
    tables; ValueError when nothing is given, for a malformed interactions list and for a frame without the columns."""
    from liana_py_b200.resource import handle_resource, select_resource, show_resources
    r = handle_resource(interactions=[("a", "b"), ("c", "d"), ("a", "b")], resource=select_resource("consensus"),
                        resource_name="consensus")
    assert r.shape == (2, 2) and list(r.columns) == ["ligand", "receptor"]            # duplicates dropped, others ignored
    r = handle_resource(interactions=None, resource=select_resource("consensus"), resource_name="ignore me")
    assert r.shape[0] == 4624 and list(r.columns) == ["ligand", "receptor"]
    r = handle_resource(interactions=None, resource=None, resource_name="CellChatDB")   # names are case-insensitive
    assert r.shape[0] == 1912
    assert "consensus" in show_resources() and len(show_resources()) == 17
    with pytest.raises(ValueError):
        handle_resource(interactions=None, resource=None, resource_name=None)
    with pytest.raises(ValueError):
        handle_resource(interactions=[("a", "b", "c")])
    with pytest.raises(ValueError):
        handle_resource(resource=pd.DataFrame({"x": ["a"], "y": ["b"]}))
    with pytest.raises(ValueError, match="not found"):
        select_resource("no_such_resource")
    # a resource that names none of the data's genes is refused by the coverage check (`_pre.py:15-62`)
    with pytest.raises(ValueError, match="Too few features"):
        _pre.assert_covered(np.array(["A", "B", "C", "D"]), np.array(["X1", "X2"]))


def test_prepare_input_refuses_an_input_without_clusters():
    adata = _ragged()
    with pytest.raises(ValueError, match="No cells left"):
        _pre.prepare_input(adata, "cluster", 10 ** 6, SciPyStage())


def test_reference_perm_chunks_equal_the_whole_stream():
    """Chunked draws of the reference's permutation stream (bounded memory, any [lo, hi) range of a rank / device) are the
    same numbers as the whole P x N array drawn at once (`_get_mean_perms.py:69-72, 79`)."""
    from liana_py_b200.method import _pipeline as PL
    n, P, seed = 37, 23, 99
    whole = PL.reference_perm_indices(n, P, seed)
    rng = np.random.default_rng(seed)
    want = np.stack([rng.permutation(np.arange(n)) for _ in range(P)])
    assert np.array_equal(whole, want) and whole.dtype == np.int32
    for lo, hi in ((0, P), (5, 17), (22, 23), (7, 7)):
        parts = list(PL.reference_perm_chunks(n, P, seed, lo, hi, max_bytes=3 * n * 4))       # 3 draws per chunk
        assert [p0 for p0, _ in parts] == list(range(lo, hi, 3))
        got = np.concatenate([c for _, c in parts]) if parts else np.zeros((0, n), np.int32)
        assert np.array_equal(got, want[lo:hi])


def test_default_perm_source_switch(monkeypatch):
    from liana_py_b200.method import _pipeline as PL
    monkeypatch.delenv("LIANA_B200_PERM_SOURCE", raising=False)
    assert PL.default_perm_source() == "philox"
    monkeypatch.setenv("LIANA_B200_PERM_SOURCE", "Reference")
    assert PL.default_perm_source() == "reference"
    monkeypatch.setenv("LIANA_B200_PERM_SOURCE", "numpy")
    with pytest.raises(ValueError):
        PL.default_perm_source()


def test_row_blocks_partition_the_matrix_without_copies():
    """Multi-rank staging: contiguous row blocks balanced by stored entries, handed to the engine as VIEWS of the caller's
    arrays (a SciPy matrix built from such views would copy them: `_prune_array`)."""
    from liana_py_b200.method._dist import DistContext
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, 1000)
    counts[100:200] = 0
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    X = sparse.csr_matrix((rng.random(indptr[-1]).astype(np.float32), rng.integers(0, 30, indptr[-1]).astype(np.int32), indptr),
                          shape=(1000, 30))
    for world in (1, 2, 3, 8):
        cuts = [DistContext(r, world).row_block(1000, X.indptr) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        loads = [int(X.indptr[b] - X.indptr[a]) for a, b in cuts]
        assert max(loads) - min(loads) <= 2 * 50 + indptr[-1] // (world * 50)
        for a, b in cuts:
            blk = _pre.CSRBlock(X, a, b)
            assert blk.shape == (b - a, 30) and np.shares_memory(blk.data, X.data) and np.shares_memory(blk.indices, X.indices)
            assert (blk.tocsr() != X[a:b]).nnz == 0
    assert DistContext(1, 4).row_block(10) == (2, 5)            # no row pointers (dense input): by rows


def test_global_code_tables_cover_every_sample():
    """`by_sample` keeps its result numeric: a sample's gene / interaction / cluster codes are re-coded through integer
    maps into tables that do not depend on the sample."""
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    res = select_resource("consensus")
    _, ex = PL.exploded(res)
    g = PL.global_name_tables(res, ["a", "b", "c"])
    assert len(g["ligand_complex"]) == len(res) and len(g["ligand"]) == len(ex.entities)
    some = np.sort(np.random.default_rng(1).choice(ex.entities, 600, replace=False))
    tb = _tables.tables_for_genes(ex, some)
    assert np.array_equal(ex.entities[tb.entity_index], tb.entities)
    assert np.array_equal(ex.ligand_complex[tb.inter_index], tb.ligand_complex)
    assert np.array_equal(np.asarray(g["receptor_complex"])[tb.inter_index], tb.receptor_complex)


def test_global_codes_travel_in_the_narrowest_type_their_table_allows():
    """`by_sample` exchanges code columns between ranks: 8 bits for clusters, 16 for the genes / interactions of a resource."""
    from types import SimpleNamespace
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    assert PL.code_dtype(256) is np.uint8 and PL.code_dtype(257) is np.uint16
    assert PL.code_dtype(65536) is np.uint16 and PL.code_dtype(65537) is np.int32
    res = select_resource("consensus")
    _, ex = PL.exploded(res)
    cats = [f"c{i}" for i in range(30)]
    g = PL.global_name_tables(res, cats)
    some = np.sort(np.random.default_rng(2).choice(ex.entities, 500, replace=False))
    tb = _tables.tables_for_genes(ex, some)
    state = SimpleNamespace(tables=tb, prep=SimpleNamespace(categories=np.array(cats[3:20])))
    maps = PL.global_code_maps(state, g)
    assert maps["source"].dtype == np.uint8 and maps["ligand"].dtype == np.uint16 and maps["ligand_complex"].dtype == np.uint16
    assert np.array_equal(maps["source"], np.arange(3, 20)) and np.array_equal(maps["ligand"], tb.entity_index)
    cols = PL.to_global_codes({"source": np.array([0, 16, 5]), "lr_means": np.array([1.0, 2.0, 3.0])}, maps)
    assert cols["source"].dtype == np.uint8 and cols["source"].tolist() == [3, 19, 8]
    frame = PL.frame_from_codes(cols, g)
    assert frame["source"].tolist() == ["c3", "c19", "c8"] and frame["lr_means"].tolist() == [1.0, 2.0, 3.0]


def test_out_of_order_rows_are_canonicalised_only_after_the_staging_pass_says_so():
    """`prepare_input` does not ask SciPy whether the matrix is canonical (a pass over every index): the staging pass
    reports `n_unordered`; only a non-zero count - or a cached SciPy flag that already says no - costs a canonical copy."""
    class Counting(SciPyStage):
        def __init__(self):
            self.seen = []

        def stage_matrix(self, X, row_keep=None):
            out = super().stage_matrix(X, row_keep)
            self.seen.append((int(X.nnz), out["n_unordered"]))
            return out

    ad = _ragged()
    X = ad.X
    rep = np.repeat(np.arange(X.nnz), 2)
    split = sparse.csr_matrix((X.data[rep] * np.float32(0.5), X.indices[rep], X.indptr * 2), shape=X.shape)
    rev = sparse.csr_matrix((np.concatenate([X.data[s:e][::-1] for s, e in zip(X.indptr[:-1], X.indptr[1:])]),
                             np.concatenate([X.indices[s:e][::-1] for s, e in zip(X.indptr[:-1], X.indptr[1:])]),
                             X.indptr.copy()), shape=X.shape)
    want = _pre.prepare_input(ad, "cluster", min_cells=5, engine=SciPyStage(), want_max=True)
    eng = Counting()
    _pre.prepare_input(ad, "cluster", min_cells=5, engine=eng)
    assert eng.seen == [(X.nnz, 0)] and "_has_canonical_format" not in X.__dict__            # canonical: one staging, SciPy not asked
    for Xn in (split, rev):
        eng = Counting()
        other = AnnDataLite(X=Xn, obs=ad.obs.copy(), var=ad.var.copy())
        got = _pre.prepare_input(other, "cluster", min_cells=5, engine=eng, want_max=True)
        assert [n for n, _ in eng.seen] == [Xn.nnz, X.nnz] and eng.seen[0][1] > 0 and eng.seen[1][1] == 0
        assert np.array_equal(got.var_names, want.var_names) and got.mat_max == want.mat_max and got.mat_mean == pytest.approx(want.mat_mean)
        assert (eng.X != X).nnz == 0 and other.X is Xn and Xn.nnz == len(Xn.data)               # staged: the canonical form; input untouched
        assert not Xn.has_canonical_format                                                    # now SciPy knows, and says no:
        eng = Counting()
        _pre.prepare_input(other, "cluster", min_cells=5, engine=eng)
        assert eng.seen == [(X.nnz, 0)]                                                       # canonicalised before the first staging
