"""Throughput of the dense -> staged CSR conversion (`lrperm_stage_matrix_dense`) against its HBM floor
(2 reads of the dense rows + 8 B per stored entry written, then the statistics pass over the CSR), and against
SciPy's `csr_matrix(X)` for a host array."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import sparse
from liana_py_b200 import _engine as E

eng = E.Engine(0)
N, G, dens = 200_000, 10_000, 0.08
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((N, G), device="cuda", generator=g)
X = torch.where(X < dens, X * 10 + 0.1, torch.zeros((), device="cuda")).contiguous()
nnz = int((X != 0).sum())
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = eng.stage_matrix_dense(X)
    dt = time.perf_counter() - t0
    algo = 2 * N * G * 4 + 8 * nnz
    print(f"device float32 {N}x{G} ({N*G*4/1e9:.1f} GB, nnz {nnz/1e6:.0f} M): {dt*1e3:.1f} ms incl. statistics pass and D2H -> "
          f"{algo/dt/1e9:.0f} GB/s of dense reads + CSR writes")
# the statistics pass alone: the same matrix staged from device CSR arrays (D2D copy + statistics)
sp = X.to_sparse_csr()
ip, ix, dv = sp.crow_indices().to(torch.int64), sp.col_indices().to(torch.int32), sp.values().contiguous()
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st_c = eng.stage_matrix_device(ip, ix, dv, (N, G))
    print(f"   same matrix from device CSR arrays (copy + statistics pass only): {(time.perf_counter() - t0) * 1e3:.1f} ms")
assert np.allclose(st["col_sums"], st_c["col_sums"])
del X, sp, ip, ix, dv
torch.cuda.empty_cache()
Nh = 50_000
Xh = np.zeros((Nh, G), dtype=np.float32)
rng = np.random.default_rng(0)
m = rng.random((Nh, G), dtype=np.float32) < dens
Xh[m] = rng.random(int(m.sum()), dtype=np.float32) + 0.1
for it in range(2):
    t0 = time.perf_counter(); st = eng.stage_matrix_dense(Xh); dt = time.perf_counter() - t0
    print(f"host float32 {Nh}x{G} ({Xh.nbytes/1e9:.1f} GB pageable): upload + conversion + statistics {dt*1e3:.0f} ms")
t0 = time.perf_counter(); S = sparse.csr_matrix(Xh); dt = time.perf_counter() - t0
print(f"scipy csr_matrix(X) on the host: {dt*1e3:.0f} ms (+ upload of the CSR)")
t0 = time.perf_counter(); st2 = eng.stage_matrix(S); print(f"   stage_matrix of that CSR: {(time.perf_counter()-t0)*1e3:.0f} ms")
assert np.array_equal(st["col_sums"] == 0, st2["col_sums"] == 0) and np.allclose(st["col_sums"], st2["col_sums"])
