"""Short program for ncu captures: one FAST batch and one EXACT batch on a BASELINE.json shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40)}[cfg]
Pf = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Pe = int(sys.argv[3]) if len(sys.argv) > 3 else 128
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 2
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
eng = E.Engine(0); eng.set_option("fast_variant", variant); eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
rng = np.random.default_rng(0); T = 1_000_000
lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
eng.set_triples(lig, rec, src, tgt, np.ones(T, np.float32), None)
for _ in range(2 if Pf else 0):
    eng.run(Pf, seed=1, order=E.ORDER_FAST, scores=E.SCORE_CPDB); print("fast", eng.timings())
for _ in range(2 if Pe else 0):
    eng.run(Pe, seed=1, order=E.ORDER_EXACT, scores=E.SCORE_CPDB); print("exact", eng.timings())
