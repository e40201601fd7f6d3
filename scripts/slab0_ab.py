"""A/B of the tile-wise first label slab (option slab0_by_tile) on the resident FAST path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
for cfg in sys.argv[1:]:
    N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40), "c5": (10000, 1877, 30)}[cfg]
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, device="cuda")
    eng = E.Engine(0)
    eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
    means, _, _ = eng.observed()
    rng = np.random.default_rng(0); T = 200000
    lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
    eng.set_triples(lig, rec, src, tgt, ((means[src, lig] + means[tgt, rec]) / 2).astype(np.float32), None)
    ref = None
    for P in (1000, 125):
        for opt in (0, 1):
            eng.set_option("slab0_by_tile", opt)
            ts = []
            for it in range(5):
                out = eng.run(P, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB)
                ts.append(eng.timings()["total_ms"])
            if opt == 0: ref = out["cpdb"].copy()
            assert np.array_equal(ref, out["cpdb"])
            print(f"{cfg} P={P} slab0_by_tile={opt}: dev total {np.median(ts):.2f} ms (min {min(ts):.2f})")
    eng.close()
