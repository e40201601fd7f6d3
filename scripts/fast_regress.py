"""Regression harness of the FAST path: dumps / compares counts and cube checksums for fixed seeds, several tile and
chunk settings (the FAST summation order is defined by the tile / chunk tables, so any refactor of how those are built
must reproduce these bit for bit).   usage: fast_regress.py dump|check file.npz"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

mode, path = sys.argv[1], sys.argv[2]
out = {}
for name, (N, G, L, P) in {"a": (50000, 1877, 20, 300), "b": (9000, 300, 50, 200), "c": (300000, 500, 7, 130)}.items():
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, device="cuda")
    eng = E.Engine(0)
    eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
    means, _, _ = eng.observed()
    rng = np.random.default_rng(0); T = 50000
    lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
    obs = ((means[src, lig] + means[tgt, rec]) / 2).astype(np.float32)
    eng.set_triples(lig, rec, src, tgt, obs, obs)
    for tile_mib, chunk in ((0, 0), (1, 1024), (4, 4096), (0, 2048)):
        eng.set_tuning(0, chunk, tile_mib)
        for overlap in (1, 0):
            eng.set_option("overlap", overlap)
            print("run", name, tile_mib, chunk, overlap, flush=True)
            r = eng.run(P, seed=11, order=E.ORDER_FAST, scores=E.SCORE_CPDB | E.SCORE_GMEAN, return_cube=(N <= 50000 and P <= 200))
            key = f"{name}_t{tile_mib}_c{chunk}_o{overlap}"
            out[key + "_cpdb"] = r["cpdb"]; out[key + "_gmean"] = r["gmean"]
            if "cube" in r:
                out[key + "_cube"] = r["cube"].view(np.uint32).astype(np.uint64).sum(axis=(1, 2))   # bit checksum per permutation
    eng.close()
if mode == "dump":
    np.savez_compressed(path, **out)
    print("dumped", len(out), "arrays")
else:
    ref = np.load(path)
    bad = [k for k in out if not np.array_equal(out[k], ref[k])]
    print("compared", len(out), "arrays:", "ALL IDENTICAL" if not bad else f"DIFFERENT: {bad[:8]}")
    sys.exit(1 if bad else 0)
