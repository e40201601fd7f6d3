"""Pageable-upload throughput of lrperm_stage_matrix_csr on the box (host threads fill pinned staging pieces that the copy
engine drains): sweep of LRPERM_COPY_THREADS / LRPERM_PIN_PIECES / LRPERM_PIN_PIECE_MB, one fresh process per setting.
   python scripts/upload_probe.py            # sweep
   python scripts/upload_probe.py one        # one measurement with the current environment"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np
    from scipy import sparse
    from liana_py_b200 import _engine as E
    n, g, per_row = 200_000, 30_000, 2500
    rng = np.random.default_rng(0)
    nnz = n * per_row
    indptr = np.arange(0, nnz + 1, per_row, dtype=np.int32)
    indices = (rng.integers(0, g // per_row, nnz, dtype=np.int32) + np.tile(np.arange(per_row, dtype=np.int32) * (g // per_row), n))
    data = rng.random(nnz, dtype=np.float32) + 0.5
    X = sparse.csr_matrix((data, indices, indptr), shape=(n, g))
    eng = E.Engine(0)
    best = None
    for _ in range(4):
        t0 = time.perf_counter(); eng.stage_matrix(X); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    gb = (indices.nbytes + data.nbytes + indptr.nbytes) / 1e9
    print(f"threads={os.environ.get('LRPERM_COPY_THREADS', 'default')} pieces={os.environ.get('LRPERM_PIN_PIECES', '2')} "
          f"piece_mb={os.environ.get('LRPERM_PIN_PIECE_MB', '16')}: {gb:.2f} GB staged in {best * 1e3:.1f} ms = {gb / best:.1f} GB/s", flush=True)
else:
    for threads, pieces, mb in ((4, 2, 16), (6, 2, 16), (8, 2, 16), (12, 2, 16), (16, 2, 16), (8, 3, 16), (8, 4, 16), (8, 2, 32), (8, 4, 8),
                                (12, 4, 8), (12, 3, 32), (6, 4, 4)):
        env = dict(os.environ, LRPERM_COPY_THREADS=str(threads), LRPERM_PIN_PIECES=str(pieces), LRPERM_PIN_PIECE_MB=str(mb))
        subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, check=False)
