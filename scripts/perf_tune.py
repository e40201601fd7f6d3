"""Sweep FAST-mode tuning knobs on a synthetic shape (device time)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
cfg = sys.argv[1]
N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40), "c5": (10000, 1877, 30)}[cfg]
P = int(sys.argv[2]); T = 200000
pbs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [512]
chs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [2048, 4096, 8192, 16384]
tiles = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [8, 16, 24, 32, 48]
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
eng = E.Engine(0); eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
rng = np.random.default_rng(0)
lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
eng.set_triples(lig, rec, src, tgt, np.ones(T, np.float32), None)
nnz = indices.numel()
for pb, ch, tile in itertools.product(pbs, chs, tiles):
    eng.set_tuning(pb, ch, tile)
    best = None
    for _ in range(3):
        eng.run(P, seed=1, order=E.ORDER_FAST, scores=E.SCORE_CPDB); tm = eng.timings()
        if best is None or tm["total_ms"] < best["total_ms"]:
            best = tm
    print(f"{cfg} PB={pb} chunk={ch} tile={tile}MiB: total {best['total_ms']:.2f} ms (setup {best['setup_ms']:.2f} means {best['means_ms']:.2f} "
          f"main {best['main_kernel_ms']:.2f} score {best['score_ms']:.2f}) upd/s {P*nnz/best['main_kernel_ms']*1e3:.3e}", flush=True)
