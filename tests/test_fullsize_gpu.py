"""GPU parity AT THE BENCHMARKED SHAPES (BASELINE.json configs[1] and configs[2]) against the CPU oracle - not against
another mode of the engine: a handful of Philox permutations is exported as index arrays (`lrperm_philox_perms`), the C
oracle (`oracle/lrperm_oracle.c`: the reference's sequential float32 arithmetic, `_get_mean_perms.py:61-87`) computes
their cube, the NumPy restatement of `_calculate_pvals` / `_cpdb_score` (`_get_mean_perms.py:117-142`,
`_cellphonedb.py:25-28`) the counts, and

  EXACT order  cube bit-identical, CellPhoneDB counts bit-identical, gmean counts within the 1-ulp tie rule
  FAST order   cube within 1e-5 relative of the ORACLE cube, counts inside the oracle's tolerance band
  FAST, production settings (no cube exported -> genes no triple reads are skipped): counts bit-identical to the
               FAST run that exported the cube
  trimean      (CellChat) cube and counts bit-identical

The triples are the rows the drop-in scores on these matrices: the bundled consensus resource, expr_prop = 0.1.
"""
import numpy as np
import pytest

from liana_py_b200 import _engine as E
from liana_py_b200.method import _pipeline as PL, _tables
from liana_py_b200.resource import select_resource
from liana_py_b200.testing._synth import consensus_genes
from oracle import c_oracle, numpy_port

pytestmark = pytest.mark.gpu

REL_MEANS = 1e-5          # north_star tolerance of the cluster means (fp32)
REL_BAND = 2e-5           # null scores this close to the observed score may compare either way in FAST order

SHAPES = {
    "configs1": (50_000, 20, 16),        # cells, clusters, permutations checked
    "configs2": (500_000, 50, 8),
}


def _workload(N, L):
    """The bench's generator (device) brought to the host as a SciPy CSR + the consensus triples at expr_prop 0.1."""
    import torch
    from scipy import sparse
    from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
    genes = consensus_genes()
    G = len(genes)
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, 1337, 1338, device=torch.device("cuda", 0))
    X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
    return X, labels.cpu().numpy(), (indptr, indices, data, labels), genes


def _triples(genes, means, stored, sizes, by=None):
    props = stored / sizes[:, None].astype(np.float64)
    tables = _tables.build_resource_tables(select_resource("consensus"), genes)
    return _tables.resolve_triples(tables, means if by is None else by, props, 0.1)


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_benchmarked_shape_against_oracle(engine, shape):
    N, L, P = SHAPES[shape]
    X, labels, dev, genes = _workload(N, L)
    G = X.shape[1]
    engine.set_matrix_csr(*dev[:3], N, G)
    engine.set_labels(dev[3], L)
    means, stored, sizes = engine.observed()
    o_means, o_stored, o_sizes = c_oracle.observed(X, labels, L)
    assert np.array_equal(means, o_means) and np.array_equal(stored, o_stored) and np.array_equal(sizes, o_sizes)
    tri = _triples(genes, means, stored, sizes)
    idx = (tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)
    assert len(tri.src) > 50_000
    obs_c = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    obs_g = PL.observed_gmean(tri.ligand_means, tri.receptor_means)
    assert np.array_equal(obs_c, numpy_port.cpdb_observed(tri.ligand_means.copy(), tri.receptor_means.copy()))
    engine.set_triples(*idx, obs_c, obs_g)

    perm_idx = engine.philox_perms(N, P, seed=1337)                       # what the production run shuffles with
    ref = c_oracle.cube(X, labels, L, perm_idx.astype(np.int64))          # the reference's arithmetic on the CPU
    ref_c = numpy_port.counts_from_cube(ref, *idx, obs_c, "cellphonedb", chunk=4)
    ref_g = numpy_port.counts_from_cube(ref, *idx, obs_g, "gmean", chunk=4)

    # ---- EXACT order: bit-identical ----
    ex = engine.run(P, seed=1337, order=E.ORDER_EXACT, scores=E.SCORE_CPDB | E.SCORE_GMEAN, return_cube=True)
    assert np.array_equal(ex["cube"], ref)
    assert np.array_equal(ex["cpdb"].astype(np.int64), ref_c)
    dg = np.abs(ex["gmean"].astype(np.int64) - ref_g)
    assert dg.max() <= 1 and (dg > 0).mean() < 1e-4                       # float64 exp/log ties within 1 ulp only
    ex_idx = engine.run(P, perm_idx=perm_idx, order=E.ORDER_EXACT, scores=E.SCORE_CPDB)
    assert np.array_equal(ex_idx["cpdb"], ex["cpdb"])                     # validation mode == Philox mode

    # ---- FAST order against the ORACLE cube ----
    fa = engine.run(P, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB | E.SCORE_GMEAN, return_cube=True)
    nz = ref != 0
    assert np.array_equal(fa["cube"] != 0, nz)
    rel = np.abs(fa["cube"][nz].astype(np.float64) - ref[nz]) / ref[nz]
    assert rel.max() <= REL_MEANS
    for name, obs, method, want in (("cpdb", obs_c, "cellphonedb", ref_c), ("gmean", obs_g, "gmean", ref_g)):
        strict, loose = numpy_port.count_band(ref, *idx, obs, method, REL_BAND, chunk=4)
        c = fa[name].astype(np.int64)
        assert np.all(c >= strict) and np.all(c <= loose)
        assert (c == want).mean() > 0.999

    # ---- production settings: no cube -> only the genes some triple reads are accumulated; same counts ----
    prod = engine.run(P, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB | E.SCORE_GMEAN)
    stats = engine.run_stats()
    needed = np.union1d(tri.lig_gene, tri.rec_gene)
    assert stats["genes_scored"] == len(needed) < G
    assert stats["nnz_scored"] == int(np.diff(X.tocsc().indptr)[needed].sum())
    assert np.array_equal(prod["cpdb"], fa["cpdb"]) and np.array_equal(prod["gmean"], fa["gmean"])


def test_configs2_trimean_against_oracle(engine):
    """CellChat null at the configs[2] shape, two permutations: float64 trimean cube and counts bit-identical."""
    N, L, P = 500_000, 50, 2
    X, labels, dev, genes = _workload(N, L)
    G = X.shape[1]
    engine.set_matrix_csr(*dev[:3], N, G)
    engine.set_labels(dev[3], L)
    _, stored, sizes = engine.observed()
    mat_max = np.float32(X.max())
    obs_tm = engine.observed_trimean(mat_max)
    assert np.array_equal(obs_tm, c_oracle.trimean_cube(X, labels, L, None, mat_max)[0])
    tri = _triples(genes, None, stored, sizes, by=obs_tm)
    idx = (tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)
    obs = PL.observed_cellchat(tri.ligand_means, tri.receptor_means)
    engine.set_triples(*idx)
    perm_idx = engine.philox_perms(N, P, seed=1337)
    ref = c_oracle.trimean_cube(X, labels, L, perm_idx.astype(np.int64), mat_max)
    out = engine.run_trimean(P, mat_max, obs_prob=obs, seed=1337, return_cube=True)
    assert np.array_equal(out["cube"], ref)
    want = numpy_port.counts_cellchat(ref, *idx, obs)
    assert np.array_equal(out["cellchat"].astype(np.int64), want)
    prod = engine.run_trimean(P, mat_max, obs_prob=obs, seed=1337)        # referenced genes only
    assert np.array_equal(prod["cellchat"], out["cellchat"])
