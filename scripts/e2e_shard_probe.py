"""What one rank of an 8-GPU e2e step does AFTER the sharded upload + all-gather (emulated on one GPU): device CSR ->
set_matrix_csr / set_labels / set_triples / run(125 permutations) / D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
N, G, L, P = 500000, 1877, 50, 125
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
hlab = labels.cpu().pin_memory()
rng = np.random.default_rng(0); T = 784548
tri = [torch.from_numpy(rng.integers(0, m, T).astype(np.int32)).pin_memory() for m in (G, G, L, L)]
obs = torch.ones(T, dtype=torch.float32).pin_memory()
counts = torch.zeros(T, dtype=torch.int32, device="cuda"); hc = torch.zeros(T, dtype=torch.int32).pin_memory()
eng = E.Engine(0)
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); ts = [time.perf_counter()]
    eng.set_matrix_csr(indptr, indices, data, N, G); sync(); ts.append(time.perf_counter())
    eng.set_labels(hlab, L); sync(); ts.append(time.perf_counter())
    eng.set_triples(tri[0], tri[1], tri[2], tri[3], obs, None); sync(); ts.append(time.perf_counter())
    eng.run(P, seed=1, order=E.ORDER_FAST, scores=E.SCORE_CPDB, counts_out=(counts, None)); sync(); ts.append(time.perf_counter())
    hc.copy_(counts); sync(); ts.append(time.perf_counter())
    d = np.diff(ts) * 1e3
    print(f"it{it}: set_matrix(device) {d[0]:.2f}  set_labels {d[1]:.2f}  set_triples {d[2]:.2f}  run {d[3]:.2f} (dev {eng.timings()['total_ms']:.2f})  d2h {d[4]:.2f}  total {d.sum():.2f} ms")
