"""Wall-clock breakdown of the host-buffer (e2e) path through the C ABI."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40)}[cfg]
P = 1000
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
host = [t.cpu().pin_memory() for t in (indptr, indices, data, labels)]
rng = np.random.default_rng(0); T = 800_000 if cfg == "c3" else 120_000
tri = [torch.from_numpy(rng.integers(0, m, T).astype(np.int32)).pin_memory() for m in (G, G, L, L)]
obs = torch.ones(T, dtype=torch.float32).pin_memory()
counts = torch.zeros(T, dtype=torch.int32, device="cuda")
hc = torch.zeros(T, dtype=torch.int32).pin_memory()
eng = E.Engine(0)
if len(sys.argv) > 2:
    eng.set_tuning(0, 0, int(sys.argv[2]))        # MiB of label slab per cell tile (default 64)
def sync(): torch.cuda.synchronize()
for it in range(4):
    ts = [time.perf_counter()]
    eng.set_matrix_csr(host[0], host[1], host[2], N, G); sync(); ts.append(time.perf_counter())
    eng.set_labels(host[3], L); sync(); ts.append(time.perf_counter())
    eng.set_triples(tri[0], tri[1], tri[2], tri[3], obs, None); sync(); ts.append(time.perf_counter())
    eng.run(P, seed=1, order=E.ORDER_FAST, scores=E.SCORE_CPDB, counts_out=(counts, None)); sync(); ts.append(time.perf_counter())
    hc.copy_(counts); sync(); ts.append(time.perf_counter())
    d = np.diff(ts) * 1e3
    print(f"{cfg} it{it}: set_matrix {d[0]:.1f}  set_labels {d[1]:.1f}  set_triples {d[2]:.1f}  run {d[3]:.1f} (dev {eng.timings()['total_ms']:.1f})  d2h {d[4]:.2f}  total {d.sum():.1f} ms")
# the same sequence without synchronising between the calls (what bench.py's e2e leg times)
for it in range(4):
    sync(); t0 = time.perf_counter()
    eng.set_matrix_csr(host[0], host[1], host[2], N, G)
    eng.set_labels(host[3], L)
    eng.set_triples(tri[0], tri[1], tri[2], tri[3], obs, None)
    eng.run(P, seed=1, order=E.ORDER_FAST, scores=E.SCORE_CPDB, counts_out=(counts, None))
    hc.copy_(counts); sync()
    print(f"{cfg} it{it}: unsynchronised sequence total {(time.perf_counter()-t0)*1e3:.1f} ms (run dev {eng.timings()['total_ms']:.1f})")
