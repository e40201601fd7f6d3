"""Run by tests/test_debug_bounds_gpu.py in a subprocess with LRPERM_LIB = the bounds-checking build of the library
(-DLRPERM_BOUNDS_CHECK): the hot kernels on the golden fixtures and on ragged / tiled / filtered inputs; every kernel-side
index check must stay silent.  Prints `BOUNDS_OK <n_runs>` on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from liana_py_b200 import _engine as E                     # noqa: E402
from liana_py_b200.testing._synth import synthetic_csr     # noqa: E402
from conftest import load_golden, load_golden_cellchat     # noqa: E402

assert os.environ.get("LRPERM_LIB", "").endswith("liblrperm_debug.so"), "run me with LRPERM_LIB=<debug build>"
eng = E.Engine(0)
assert eng.debug_bounds()[0] == 0                           # the debug build answers (the production library raises)
runs = 0


def check(what):
    global runs
    flags, first = eng.debug_bounds()
    assert flags == 0, f"{what}: bounds violation bits {flags:#x}, first (bit, value) = {first}"
    runs += 1


for name in ("toy700", "small300", "ragged90"):
    g = load_golden(name)
    P = int(g["n_perms"])
    eng.set_matrix(g["X"]); eng.set_labels(g["labels"], g["L"])
    eng.observed()
    obs = ((g["ligand_means"] + g["receptor_means"]) / 2).astype(np.float32)
    eng.set_triples(g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"], obs, obs)
    for order in (E.ORDER_EXACT, E.ORDER_FAST):
        eng.run(P, perm_idx=g["perm_idx"], order=order, scores=E.SCORE_CPDB | E.SCORE_GMEAN, return_cube=True)
        eng.run(P, seed=3, order=order, scores=E.SCORE_CPDB)
        check(f"{name} order {order}")
for name in ("small300", "dense700", "ragged90"):
    g = load_golden_cellchat(name)
    eng.set_matrix(g["X"]); eng.set_labels(g["labels"], g["L"])
    eng.observed_trimean(g["mat_max"])
    eng.set_triples(g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"])
    eng.run_trimean(int(g["n_perms"]), g["mat_max"], obs_prob=g["lr_probs"], seed=int(g["seed"]), return_cube=True)
    eng.run_trimean(40, g["mat_max"], obs_prob=g["lr_probs"], seed=1)
    check(f"{name} trimean")

# ragged sizes, many clusters (2 permutations per lane), small tiles / chunks, the bulk-copy variant, the gene filter, a pinned upload
rng = np.random.default_rng(0)
for N, G, L, dens in ((257, 33, 4, 0.3), (8000, 120, 200, 0.15), (40000, 257, 9, 0.1)):
    X, labels = synthetic_csr(N, G, L, density=dens, seed_data=N, seed_labels=N + 1)
    labels = (np.arange(N) % L).astype(np.int32)
    rng.shuffle(labels)
    T = 5000
    genes = rng.choice(G, max(3, G // 3), replace=False)
    lig, rec = rng.choice(genes, T), rng.choice(genes, T)
    src, tgt = rng.integers(0, L, T), rng.integers(0, L, T)
    for tile_mib, chunk, variant in ((0, 0, 2), (1, 1024, 2), (1, 1024, 3), (0, 0, 3)):
        eng.set_tuning(128 if N > 1000 else 0, chunk, tile_mib)
        eng.set_option("fast_variant", variant)
        eng.set_matrix(X); eng.set_labels(labels, L)
        means, _, _ = eng.observed()
        obs = ((means[src, lig] + means[tgt, rec]) / 2).astype(np.float32)
        eng.set_triples(lig, rec, src, tgt, obs, obs)
        for P in (1, 33, 300):
            eng.run(P, seed=5, order=E.ORDER_FAST, scores=E.SCORE_CPDB | E.SCORE_GMEAN)                    # filtered
            eng.run(min(P, 64), seed=5, order=E.ORDER_FAST, scores=E.SCORE_CPDB, return_cube=True)         # every gene
            eng.run(min(P, 64), seed=5, order=E.ORDER_EXACT, scores=E.SCORE_CPDB)
        check(f"synthetic N={N} L={L} tile={tile_mib} chunk={chunk} variant={variant}")
    eng.set_option("fast_variant", 2)
    eng.set_tuning(0, 0, 0)
import torch                                               # noqa: E402
X, labels = synthetic_csr(40000, 257, 9, density=0.1, seed_data=21, seed_labels=22)
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in
          (X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32))]
eng.set_tuning(0, 0, 1)
eng.set_matrix_csr(pinned[0], pinned[1], pinned[2], 40000, 257)
eng.set_labels(labels, 9)
lig, rec, src, tgt = rng.integers(0, 100, 4000), rng.integers(0, 100, 4000), rng.integers(0, 9, 4000), rng.integers(0, 9, 4000)
eng.set_triples(lig, rec, src, tgt, np.ones(4000, np.float32), None)
eng.run(700, seed=5, order=E.ORDER_FAST, scores=E.SCORE_CPDB)                                               # tile-gated behind the upload
check("pinned upload, tile-gated run")
print("BOUNDS_OK", runs)
