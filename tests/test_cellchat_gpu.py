"""GPU parity tests of the CellChat trimean null (lrperm_run_trimean / lrperm_observed_trimean, through the
C ABI) against the fixtures produced by the reference's own li.mt.cellchat and against the C oracle.
Order statistics are selected, not summed: every comparison here is BIT-EXACT."""
import numpy as np
import pytest
from scipy import sparse

from liana_py_b200 import _engine as E
from oracle import c_oracle, numpy_port

pytestmark = pytest.mark.gpu


def _idx(g):
    return g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"]


def _perm_idx(g):
    return numpy_port.perm_indices(g["X"].shape[0], int(g["n_perms"]), int(g["seed"]))


def test_trimean_cube_and_counts_match_reference_golden(engine, golden_cc):
    g = golden_cc
    P = int(g["n_perms"])
    engine.set_matrix(g["X"])
    engine.set_labels(g["labels"], g["L"])
    engine.set_triples(*_idx(g))
    out = engine.run_trimean(P, g["mat_max"], obs_prob=g["lr_probs"], perm_idx=_perm_idx(g), return_cube=True)
    assert out["cube"].dtype == np.float64
    assert np.array_equal(out["cube"], g["cube"])                       # bit-identical float64 cube
    assert np.array_equal(out["cellchat"] / P, g["pvals"])              # p-values bit-identical


def test_observed_trimean_matches_reference_golden(engine, golden_cc):
    g = golden_cc
    engine.set_matrix(g["X"])
    engine.set_labels(g["labels"], g["L"])
    obs = engine.observed_trimean(g["mat_max"])
    assert np.array_equal(obs, c_oracle.trimean_cube(g["X"], g["labels"], g["L"], None, g["mat_max"])[0])
    assert np.array_equal(obs[g["source_idx"], g["ligand_idx"]], g["ligand_trimean"])
    assert np.array_equal(obs[g["target_idx"], g["receptor_idx"]], g["receptor_trimean"])


def _random_case(seed, N, G, L, negatives=False, explicit_zeros=False, skew=False):
    rng = np.random.default_rng(seed)
    dens = rng.beta(0.8, 1.2, size=G)
    dense = (rng.random((N, G)) < dens[None, :]) * np.log1p(rng.gamma(1.5, 2.0, (N, G)))
    if negatives:
        dense *= np.where(rng.random((N, G)) < 0.3, -1.0, 1.0)
    X = sparse.csr_matrix(dense.astype(np.float32))
    if explicit_zeros:
        X.data[::13] = 0.0
    if skew:
        p = rng.dirichlet(np.full(L, 0.5))
        labels = rng.choice(L, size=N, p=p).astype(np.int32)
    else:
        labels = rng.integers(0, L, N).astype(np.int32)
    labels[:L] = np.arange(L)
    return X, labels


@pytest.mark.parametrize("case", [
    dict(seed=1, N=20000, G=24, L=7, chunk=1024),                                   # many chunks per gene: prefix + select
    dict(seed=2, N=9000, G=40, L=50, chunk=1024, skew=True),                        # 50 unequal clusters (tiny ones too)
    dict(seed=3, N=6000, G=30, L=5, chunk=2048, negatives=True, explicit_zeros=True),
    dict(seed=4, N=3000, G=16, L=130, chunk=1024, skew=True),                       # > 128 clusters: 2 permutations per lane
    dict(seed=5, N=50, G=12, L=3, chunk=0),                                          # tiny clusters, odd / even sizes
])
def test_trimean_cube_matches_oracle_multichunk(engine, case):
    X, labels = _random_case(case["seed"], case["N"], case["G"], case["L"], case.get("negatives", False),
                             case.get("explicit_zeros", False), case.get("skew", False))
    L = case["L"]
    mat_max = np.float32(X.max())
    P = 6
    perm_idx = numpy_port.perm_indices(X.shape[0], P, 11)
    engine.set_option("trimean_chunk_nnz", case["chunk"])
    try:
        engine.set_matrix(X)
        engine.set_labels(labels, L)
        out = engine.run_trimean(P, mat_max, perm_idx=perm_idx, return_cube=True)
        ref = c_oracle.trimean_cube(X, labels, L, perm_idx, mat_max)
        assert np.array_equal(out["cube"], ref)
        assert np.mean(ref != 0) > 0.1
        obs = engine.observed_trimean(mat_max)
        assert np.array_equal(obs, c_oracle.trimean_cube(X, labels, L, None, mat_max)[0])
    finally:
        engine.set_option("trimean_chunk_nnz", 0)


def test_trimean_philox_equals_index_mode_and_shards_add_up(engine):
    X, labels = _random_case(9, 12000, 20, 6)
    L = 6
    mat_max = np.float32(X.max())
    engine.set_option("trimean_chunk_nnz", 2048)
    try:
        engine.set_matrix(X)
        engine.set_labels(labels, L)
        rng = np.random.default_rng(0)
        T = 500
        lig, rec = rng.integers(0, X.shape[1], T), rng.integers(0, X.shape[1], T)
        src, tgt = rng.integers(0, L, T), rng.integers(0, L, T)
        engine.set_triples(lig, rec, src, tgt)
        obs_tm = engine.observed_trimean(mat_max)
        obs = numpy_port.lr_probability((obs_tm[src, lig], obs_tm[tgt, rec]))
        P = 200                                                      # two batches of 128
        engine.set_option("trimean_perm_batch", 128)
        ph = engine.run_trimean(P, mat_max, obs_prob=obs, seed=5, return_cube=True)
        perms = engine.philox_perms(X.shape[0], P, seed=5)
        ix = engine.run_trimean(P, mat_max, obs_prob=obs, perm_idx=perms, return_cube=True)
        assert np.array_equal(ph["cube"], ix["cube"]) and np.array_equal(ph["cellchat"], ix["cellchat"])
        # counts against the oracle's score on the engine's own cube (the cube itself is checked above)
        assert np.array_equal(ph["cellchat"].astype(np.int64), numpy_port.counts_cellchat(ph["cube"], lig, rec, src, tgt, obs))
        a = engine.run_trimean(90, mat_max, obs_prob=obs, seed=5, perm_offset=0)["cellchat"]
        b = engine.run_trimean(110, mat_max, obs_prob=obs, seed=5, perm_offset=90)["cellchat"]
        assert np.array_equal(a + b, ph["cellchat"])
        ref = c_oracle.trimean_cube(X, labels, L, perms[:8].astype(np.int64), mat_max)
        assert np.array_equal(ph["cube"][:8], ref)
    finally:
        engine.set_option("trimean_chunk_nnz", 0)
        engine.set_option("trimean_perm_batch", 0)


# ---- the drop-in li.mt.cellchat against frames produced by the reference's own code --------------------------
import os  # noqa: E402

import pandas as pd  # noqa: E402

import liana_py_b200 as li  # noqa: E402
from liana_py_b200.testing._anndata_lite import AnnDataLite  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = ["source", "target", "ligand_complex", "receptor_complex"]


def _load_input(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_cellchat.npz"))
    X = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=tuple(z["in_shape"]))
    obs = pd.DataFrame({"cluster": pd.Categorical(z["in_labels"])}, index=[f"cell{i}" for i in range(X.shape[0])])
    var = pd.DataFrame(index=pd.Index(z["in_var_names"]))
    return AnnDataLite(X=X, obs=obs, var=var), int(z["n_perms"]), int(z["seed"]), float(z["expr_prop"])


@pytest.mark.parametrize("name", ["small300", "dense700", "ragged90"])
def test_cellchat_frame_identical_to_reference(name):
    adata, P, seed, ep = _load_input(name)
    res = li.mt.cellchat(adata, groupby="cluster", use_raw=False, n_perms=P, seed=seed, expr_prop=ep, inplace=False,
                         perm_source="reference")
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_cellchat.csv.gz"), float_precision="round_trip")
    assert list(res.columns) == list(gold.columns)
    a = res.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    assert a["mat_max"].dtype == np.float32 and np.array_equal(a["mat_max"].values, b["mat_max"].values.astype(np.float32))
    for c in ["ligand_trimean", "receptor_trimean", "ligand_props", "receptor_props", "lr_probs", "cellchat_pvals"]:
        assert a[c].dtype == np.float64
        assert np.array_equal(a[c].values, b[c].values), c                  # bit-identical float64
    assert np.all(np.diff(res["lr_probs"].values) <= 0)                     # sorted by magnitude, descending


def test_cellchat_philox_default_and_no_perms():
    adata, P, seed, ep = _load_input("dense700")
    res = li.mt.cellchat(adata, groupby="cluster", use_raw=False, n_perms=None, inplace=False)
    assert "cellchat_pvals" not in res.columns and "lr_probs" in res.columns
    ph = li.mt.cellchat(adata, groupby="cluster", use_raw=False, n_perms=300, inplace=False)    # Philox permutations
    gold = pd.read_csv(os.path.join(GOLDEN, "dense700_cellchat.csv.gz"), float_precision="round_trip")
    a, b = ph.sort_values(KEY), gold.sort_values(KEY)
    assert np.array_equal(a["lr_probs"].values, b["lr_probs"].values)
    assert abs(a["cellchat_pvals"].values.mean() - b["cellchat_pvals"].values.mean()) < 0.03   # another RNG, same null
    # a custom consensus with CellChat in it
    agg = li.mt.AggregateClass(li.mt.aggregate_meta, methods=[li.mt.cellphonedb, li.mt.cellchat])
    out = agg(adata, groupby="cluster", use_raw=False, n_perms=50, inplace=False)
    assert {"lr_means", "cellphone_pvals", "lr_probs", "cellchat_pvals", "magnitude_rank", "specificity_rank"} <= set(out.columns)


def test_cellchat_mid_size_properties(engine):
    """50k cells x 600 genes x 20 clusters, skewed densities: several chunks per gene, two batches of 512, super-batch
    select.  Cube of the first permutations against the C oracle bit for bit; determinism; permutation shards add up;
    rows whose observed score is 0 count every permutation; counts equal the oracle's score on the engine's cube."""
    from liana_py_b200.testing._synth import synthetic_csr
    N, G, L, P = 50_000, 600, 20, 640
    X, labels = synthetic_csr(N, G, L, density=0.2, seed_data=5, seed_labels=6)
    mat_max = np.float32(X.max())
    engine.set_matrix(X)
    engine.set_labels(labels, L)
    rng = np.random.default_rng(2)
    T = 30_000
    lig, rec = rng.integers(0, G, T), rng.integers(0, G, T)
    src, tgt = rng.integers(0, L, T), rng.integers(0, L, T)
    engine.set_triples(lig, rec, src, tgt)
    obs_tm = engine.observed_trimean(mat_max)
    assert np.array_equal(obs_tm, c_oracle.trimean_cube(X, labels, L, None, mat_max)[0])
    obs = numpy_port.lr_probability((obs_tm[src, lig], obs_tm[tgt, rec]))
    assert 0.05 < np.mean(obs > 0) < 0.95
    a = engine.run_trimean(P, mat_max, obs_prob=obs, seed=9)["cellchat"]
    b = engine.run_trimean(P, mat_max, obs_prob=obs, seed=9)["cellchat"]
    assert np.array_equal(a, b)                                              # deterministic
    assert np.all(a[obs == 0] == P)                                          # null score >= 0 always
    h1 = engine.run_trimean(300, mat_max, obs_prob=obs, seed=9, perm_offset=0)["cellchat"]
    h2 = engine.run_trimean(P - 300, mat_max, obs_prob=obs, seed=9, perm_offset=300)["cellchat"]
    assert np.array_equal(h1 + h2, a)                                        # shards by global permutation id add up
    few = engine.run_trimean(3, mat_max, obs_prob=obs, seed=9, return_cube=True)
    perms = engine.philox_perms(N, 3, seed=9)
    ref = c_oracle.trimean_cube(X, labels, L, perms.astype(np.int64), mat_max)
    assert np.array_equal(few["cube"], ref) and np.mean(ref != 0) > 0.05
    assert np.array_equal(few["cellchat"].astype(np.int64), numpy_port.counts_cellchat(ref, lig, rec, src, tgt, obs))


@pytest.mark.parametrize("name", ["small300", "dense700", "ragged90"])
def test_scseqcomm_frame_matches_reference(name):
    """li.mt.scseqcomm (non-permutation; cluster statistics of the whole prepared matrix from the staged upload,
    `lrperm_staged_cluster_stats`) against the reference's own frame.  The reference computes the cluster mean / std
    in float32 (`np.std` of the densified float32 rows), here they are float64 sums: the normal cdf agrees to ~1e-5."""
    adata, _, _, ep = _load_input(name)
    res = li.mt.scseqcomm(adata, groupby="cluster", use_raw=False, expr_prop=ep, inplace=False)
    gold = pd.read_csv(os.path.join(GOLDEN, f"{name}_scseqcomm.csv.gz"), float_precision="round_trip")
    assert list(res.columns) == list(gold.columns)
    a = res.sort_values(KEY).reset_index(drop=True)
    b = gold.sort_values(KEY).reset_index(drop=True)
    assert len(a) == len(b)
    for c in KEY + ["ligand", "receptor"]:
        assert (a[c].astype(str).values == b[c].astype(str).values).all(), c
    for c in ["ligand_means", "receptor_means"]:
        assert np.array_equal(a[c].values, b[c].values.astype(np.float32)), c
    for c in ["ligand_cdf", "receptor_cdf", "inter_score"]:
        np.testing.assert_allclose(a[c].values, b[c].values, rtol=1e-4, atol=2e-5, err_msg=c)
    assert np.all(np.diff(res["inter_score"].values) <= 0)


@pytest.mark.parametrize("method", ["cellchat", "scseqcomm", "rank_aggregate"])
def test_by_sample_of_other_methods_matches_per_sample_calls(method):
    """`.by_sample` (`liana/method/sc/_Method.py:92-159`) is inherited by every method: the concatenated frame must be
    the per-sample calls in sample-category order (CellChat needs mat_max per sample, scSeqComm the per-sample
    cluster statistics of the staged matrix)."""
    adata, P, seed, ep = _load_input("dense700")
    rng = np.random.default_rng(1)
    adata.obs["sample"] = pd.Categorical(rng.choice(["s1", "s2", "s3"], size=adata.shape[0]))
    fn = getattr(li.mt, method)
    kw = dict(groupby="cluster", use_raw=False, expr_prop=ep, n_perms=40, seed=seed)
    res = fn.by_sample(adata, "sample", inplace=False, **kw)
    assert res.columns[0] == "sample" and list(dict.fromkeys(res["sample"])) == ["s1", "s2", "s3"]
    for s in ["s1", "s2", "s3"]:
        one = fn(adata[np.asarray(adata.obs["sample"] == s)].copy(), inplace=False, **kw)
        got = res[res["sample"] == s].drop(columns="sample")
        a = got.sort_values(KEY).reset_index(drop=True)
        b = one.sort_values(KEY).reset_index(drop=True)
        assert list(a.columns) == list(b.columns) and len(a) == len(b)
        for c in a.columns:
            if a[c].dtype.kind == "f":
                assert np.array_equal(a[c].values, b[c].values, equal_nan=True), (s, c)
