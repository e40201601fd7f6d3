"""CPU: pin the oracle (oracle/lrperm_oracle.c + oracle/numpy_port.py) against fixtures produced
by the reference's own code (scripts/make_golden.py; liana/method/_pipe_utils/_get_mean_perms.py,
liana/method/sc/_cellphonedb.py, _geometric_mean.py executed unmodified)."""
import numpy as np
import pytest

from oracle import c_oracle, numpy_port


def test_perm_indices_match_reference_draws(golden):
    # the reference draws rng.permutation(idx) P times from default_rng(seed) (_get_mean_perms.py:69,79)
    P, N = int(golden["n_perms"]), golden["X"].shape[0]
    assert np.array_equal(numpy_port.perm_indices(N, P, int(golden["seed"])), golden["perm_idx"])


def test_c_cube_bit_exact_vs_reference(golden):
    kc = golden["cube"].shape[0]
    cube = c_oracle.cube(golden["X"], golden["labels"], golden["L"], golden["perm_idx"][:kc])
    assert cube.dtype == np.float32
    assert np.array_equal(cube, golden["cube"])      # bit-for-bit


def test_c_cube_cluster_sums(golden):
    # the quantity liana/tests/test_get_mean_perms.py:27-35 pins (sum over perms and genes per cluster)
    cube = c_oracle.cube(golden["X"], golden["labels"], golden["L"], golden["perm_idx"])
    np.testing.assert_allclose(cube.astype(np.float64).sum(axis=(0, 2)), golden["cube_cluster_sums"], rtol=0, atol=1e-9)


def test_numpy_port_cube_bit_exact(golden):
    kc = min(4, golden["cube"].shape[0])
    cube = numpy_port.get_means_perms(golden["X"], golden["labels"], golden["L"], kc, int(golden["seed"]))
    assert np.array_equal(cube.astype(np.float32), golden["cube"][:kc])


def test_observed_means_match_reference(golden):
    means, stored, sizes = c_oracle.observed(golden["X"], golden["labels"], golden["L"])
    lm = means[golden["source_idx"], golden["ligand_idx"]]
    rm = means[golden["target_idx"], golden["receptor_idx"]]
    assert np.array_equal(lm, golden["ligand_means"])
    assert np.array_equal(rm, golden["receptor_means"])
    p_means, p_props = numpy_port.observed_stats(golden["X"], golden["labels"], golden["L"])
    assert np.array_equal(p_means, means)
    assert np.array_equal(p_props, stored / sizes[:, None])


def test_scores_and_pvals_match_reference(golden):
    g = golden
    cube = c_oracle.cube(g["X"], g["labels"], g["L"], g["perm_idx"])
    P = cube.shape[0]
    idx = (g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"])
    obs_c = numpy_port.cpdb_observed(g["ligand_means"].copy(), g["receptor_means"].copy())
    obs_g = numpy_port.gmean_observed(g["ligand_means"], g["receptor_means"])
    assert np.array_equal(obs_c, g["cpdb_magnitude"])
    assert np.array_equal(obs_g.astype(np.float32), g["gmean_magnitude"])
    cnt_c = numpy_port.counts_from_cube(cube, *idx, obs_c, "cellphonedb")
    cnt_g = numpy_port.counts_from_cube(cube, *idx, obs_g, "gmean")
    assert np.array_equal(cnt_c / P, g["cpdb_pvals"])
    assert np.array_equal(cnt_g / P, g["gmean_pvals"])
    # C restatement of the CellPhoneDB count
    assert np.array_equal(c_oracle.counts_cpdb(cube, *idx, obs_c).astype(np.int64), cnt_c)
    # zero-masked rows: every null mean >= 0 -> p == 1 exactly
    assert np.all(cnt_c[obs_c == 0] == P)


def test_run_method_block_port(golden):
    g = golden
    P = min(6, int(g["n_perms"]))
    lr_means, pv = numpy_port.run_method_block(
        g["X"], g["labels"], g["L"], P, int(g["seed"]), g["ligand_idx"], g["receptor_idx"], g["source_idx"],
        g["target_idx"], g["ligand_means"].copy(), g["receptor_means"].copy(), "cellphonedb")
    cube = c_oracle.cube(g["X"], g["labels"], g["L"], g["perm_idx"][:P])
    cnt = c_oracle.counts_cpdb(cube, g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"], lr_means)
    assert np.array_equal(pv, cnt / P)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 700, 1000, 4097])
def test_philox_bijection_is_a_permutation(n):
    perms = c_oracle.philox_perms(n, 5, 3, 1337)
    for p in range(5):
        assert np.array_equal(np.sort(perms[p]), np.arange(n))
        inv = c_oracle.philox_inverse(n, 3 + p, 1337)
        assert np.array_equal(inv[perms[p]], np.arange(n))
    if n > 3:
        assert not np.array_equal(perms[0], perms[1])
    assert np.array_equal(perms[2], c_oracle.philox_perms(n, 1, 5, 1337)[0])   # keyed by global perm id


def test_philox_positions_are_uniform():
    # each cell should land on each position ~uniformly over permutations (chi-square on a coarse grid)
    n, P, bins = 200, 4000, 10
    perms = c_oracle.philox_perms(n, P, 0, 99)
    # position of cell 0..4 across permutations
    for cell in range(5):
        pos = np.argmax(perms == cell, axis=1)
        hist = np.bincount(pos * bins // n, minlength=bins)
        chi2 = ((hist - P / bins) ** 2 / (P / bins)).sum()
        assert chi2 < 40.0, (cell, chi2)       # 9 dof, p ~ 1e-5
    # adjacent cells stay adjacent no more often than chance (~2/n)
    inv = np.argsort(perms, axis=1)
    adj = (np.abs(inv[:, 0] - inv[:, 1]) == 1).mean()
    assert adj < 4.0 / n + 0.02


def test_philox_small_domain_pairs_are_uniform():
    # N = 6: the joint position of any two cells is uniform over the 30 ordered pairs of distinct positions.
    # (A Feistel network only reaches even permutations of its power-of-two domain, so the distribution over
    # ALL n! permutations of a tiny domain is not uniform; the permutation null only needs the label-level,
    # i.e. low-order, statistics tested here and in test_philox_label_mixing_matches_hypergeometric.)
    n, P = 6, 30 * 1000
    perms = c_oracle.philox_perms(n, P, 0, 2024)
    inv = np.argsort(perms, axis=1)                         # inv[p][cell] = position
    for a_, b_ in ((0, 1), (2, 5), (4, 3)):
        table = np.zeros((n, n))
        np.add.at(table, (inv[:, a_], inv[:, b_]), 1)
        off = table[~np.eye(n, dtype=bool)]
        assert np.all(np.diag(table) == 0)
        chi2 = ((off - P / 30) ** 2 / (P / 30)).sum()
        assert chi2 < 75.0, (a_, b_, chi2)                  # 29 dof, mean 29, sd 7.6 -> 6 sd


def test_philox_consecutive_ids_are_independent():
    # where cell 0 sits under permutation id p vs id p+1 (and under seed s vs s+1): 7 x 7 contingency tables
    n, P = 7, 49 * 400
    a = c_oracle.philox_perms(n, P + 1, 0, 7)
    pos = np.argmax(a == 0, axis=1)
    table = np.zeros((n, n))
    np.add.at(table, (pos[:-1], pos[1:]), 1)
    chi2 = ((table - P / 49) ** 2 / (P / 49)).sum()
    assert chi2 < 110.0, chi2                   # 48 dof (36 for independence + margins), mean ~48, sd ~10
    b = c_oracle.philox_perms(n, P, 0, 8)
    posb = np.argmax(b == 0, axis=1)
    table = np.zeros((n, n))
    np.add.at(table, (pos[:P], posb), 1)
    chi2 = ((table - P / 49) ** 2 / (P / 49)).sum()
    assert chi2 < 110.0, chi2


def test_philox_label_mixing_matches_hypergeometric():
    # N cells in 3 clusters of unequal size: the number of cluster-a cells that receive label b under a random
    # permutation is hypergeometric; check mean and variance over many permutations at a non-power-of-two N
    for n, sizes, P in ((3001, [1700, 1000, 301], 1500), (64, [30, 20, 14], 4000), (17, [6, 6, 5], 4000)):
        sizes = np.array(sizes)
        lab = np.repeat(np.arange(3), sizes)
        perms = c_oracle.philox_perms(n, P, 5, 31337)            # perms[p][j] = cell at position j
        for a in range(3):
            for b in range(3):
                # cells of original cluster a sitting on positions labelled b
                x = (lab[perms][:, lab == b] == a).sum(axis=1)
                K, m = sizes[a], sizes[b]
                mean = m * K / n
                var = m * (K / n) * (1 - K / n) * (n - m) / (n - 1)
                assert abs(x.mean() - mean) < 5 * np.sqrt(var / P), (n, a, b, x.mean(), mean)
                assert 0.85 < x.var() / var < 1.18, (n, a, b, x.var(), var)


@pytest.mark.parametrize("seed", range(6))
def test_c_restatement_equals_literal_numpy_port_on_random_inputs(seed):
    """The two oracles against each other beyond the fixtures: the C restatement (sequential float32 sums; exact
    quartile selection) and the literal NumPy/SciPy port (the reference's own calls) on random shapes - ragged
    clusters, explicit zeros, negative values, dense and very sparse matrices - bit for bit."""
    from scipy import sparse
    rng = np.random.default_rng(seed)
    N, G, L = int(rng.integers(20, 160)), int(rng.integers(3, 40)), int(rng.integers(1, 7))
    dens = float(rng.choice([0.02, 0.2, 0.7, 1.0]))
    X = sparse.random(N, G, density=dens, format="csr", dtype=np.float32, random_state=seed,
                      data_rvs=lambda k: (rng.gamma(1.5, 2.0, k) - (1.0 if seed % 2 else 0.0)).astype(np.float32))
    if X.nnz:
        X.data[::7] = 0.0                                   # explicit stored zeros
    labels = rng.integers(0, L, N).astype(np.int32)
    labels[:L] = np.arange(L)                               # no empty cluster: SciPy's mean divides by the row count
    P = 5
    perm_idx = numpy_port.perm_indices(N, P, 100 + seed)
    want = numpy_port.get_means_perms(X, labels, L, P, 100 + seed)
    assert np.array_equal(c_oracle.cube(X, labels, L, perm_idx), want.astype(np.float32))
    if X.nnz:
        mat_max = np.float32(max(X.max(), 0.5))
        want_t = numpy_port.get_trimean_perms(X, labels, L, P, 100 + seed, mat_max)
        assert np.array_equal(c_oracle.trimean_cube(X, labels, L, perm_idx, mat_max), want_t)
        assert np.array_equal(c_oracle.trimean_cube(X, labels, L, None, mat_max)[0], numpy_port.observed_trimean(X, labels, L, mat_max))
