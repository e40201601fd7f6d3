"""CPU: pin the CellChat restatements (oracle/lrperm_oracle.c `oracle_trimean_cube`, oracle/numpy_port.py
`trimean` / `lr_probability`) against fixtures produced by the reference's own `li.mt.cellchat`
(scripts/make_golden.py --cellchat-only; liana/method/sc/_cellchat.py, _liana_pipe.py:307-308, 394-397, 462-465,
_get_mean_perms.py:43-44 executed unmodified)."""
import numpy as np

from oracle import c_oracle, numpy_port


def _idx(g):
    return g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"]


def test_c_trimean_cube_bit_exact_vs_reference(golden_cc):
    g = golden_cc
    P, N = int(g["n_perms"]), g["X"].shape[0]
    perm_idx = numpy_port.perm_indices(N, P, int(g["seed"]))
    cube = c_oracle.trimean_cube(g["X"], g["labels"], g["L"], perm_idx, g["mat_max"])
    assert cube.dtype == np.float64 and np.array_equal(cube, g["cube"])          # bit-for-bit
    assert np.mean(cube != 0) > 0.2                                               # not a trivial all-zero cube


def test_numpy_port_trimean_cube_bit_exact(golden_cc):
    g = golden_cc
    cube = numpy_port.get_trimean_perms(g["X"], g["labels"], g["L"], 3, int(g["seed"]), g["mat_max"])
    assert np.array_equal(cube, g["cube"][:3])


def test_observed_trimean_matches_reference(golden_cc):
    g = golden_cc
    obs = c_oracle.trimean_cube(g["X"], g["labels"], g["L"], None, g["mat_max"])[0]
    assert np.array_equal(obs, numpy_port.observed_trimean(g["X"], g["labels"], g["L"], g["mat_max"]))
    assert np.array_equal(obs[g["source_idx"], g["ligand_idx"]], g["ligand_trimean"])
    assert np.array_equal(obs[g["target_idx"], g["receptor_idx"]], g["receptor_trimean"])


def test_cellchat_scores_and_pvals_match_reference(golden_cc):
    g = golden_cc
    P = int(g["n_perms"])
    lr_prob = numpy_port.lr_probability((g["ligand_trimean"], g["receptor_trimean"]))
    assert np.array_equal(lr_prob, g["lr_probs"])
    cnt = numpy_port.counts_cellchat(g["cube"], *_idx(g), lr_prob)
    assert np.array_equal(cnt / P, g["pvals"])
    assert np.array_equal(c_oracle.counts_cellchat(g["cube"], *_idx(g), lr_prob).astype(np.int64), cnt)
    # the literal score function on the materialised perm_stats (what `_cellchat_score` does)
    ps = numpy_port.gather_perm_stats(g["cube"], *_idx(g))
    _, pv = numpy_port.cellchat_score(g["ligand_trimean"], g["receptor_trimean"], ps)
    assert np.array_equal(pv, g["pvals"])
