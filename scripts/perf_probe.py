"""Quick device-time probe of the engine on synthetic shapes (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40), "c5": (10000, 1877, 30)}[cfg]
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["fast", "exact"]
    t0 = time.time()
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, device="cuda")
    torch.cuda.synchronize()
    print(f"gen {time.time()-t0:.2f}s nnz={indices.numel()}")
    eng = E.Engine(0)
    eng.set_matrix_csr(indptr, indices, data, N, G)
    eng.set_labels(labels, L)
    means, stored, sizes = eng.observed()
    rng = np.random.default_rng(0)
    lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
    obs = (means[src, lig] + means[tgt, rec]) / 2
    eng.set_triples(lig, rec, src, tgt, obs.astype(np.float32), obs.astype(np.float32))
    nnz = indices.numel()
    model_bytes = P * (8 * nnz + 4 * (N + 1) + 4 * N + 8 * L * G) + 24 * T
    for mode in modes:
        order = E.ORDER_FAST if mode == "fast" else E.ORDER_EXACT
        for it in range(3):
            t0 = time.time()
            out = eng.run(P, seed=1337, order=order, scores=E.SCORE_CPDB)
            wall = time.time() - t0
            tm = eng.timings()
            print(f"{cfg} {mode} P={P} T={T}: wall {wall*1e3:.1f} ms  dev total {tm['total_ms']:.2f} (setup {tm['setup_ms']:.2f} means {tm['means_ms']:.2f} score {tm['score_ms']:.2f}) "
                  f"launches {tm['launches']}  model {model_bytes/tm['total_ms']*1e3/1e12:.2f} TB/s  upd/s {P*nnz/tm['means_ms']*1e3:.3e}")
main()
