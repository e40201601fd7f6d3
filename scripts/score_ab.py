"""A/B of the permutation blocking of the score kernels (option score_groups_per_block; 1024 = no blocking)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
cfg = sys.argv[1]; T = int(sys.argv[2])
N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40)}[cfg]
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, device="cuda")
eng = E.Engine(0)
eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
means, _, _ = eng.observed()
rng = np.random.default_rng(0)
lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
obs = ((means[src, lig] + means[tgt, rec]) / 2).astype(np.float32)
eng.set_triples(lig, rec, src, tgt, obs, obs)
ref = None
for gpb in (1024, 0, 2, 4, 8):
    eng.set_option("score_groups_per_block", gpb)
    for it in range(3):
        out = eng.run(1000, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB | E.SCORE_GMEAN)
    tm = eng.timings()
    if ref is None: ref = (out["cpdb"].copy(), out["gmean"].copy())
    assert np.array_equal(ref[0], out["cpdb"]) and np.array_equal(ref[1], out["gmean"])
    print(f"{cfg} T={T} gpb={gpb}: total {tm['total_ms']:.2f} ms score {tm['score_ms']:.2f} ms")
