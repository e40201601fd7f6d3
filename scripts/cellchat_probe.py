"""Device-time probe of the CellChat trimean null on synthetic shapes (not the bench contract).
usage: cellchat_probe.py c2|c3|c4|c5 [P] [T] [chunk_nnz,...] [perm_batch,...]"""
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
    chunks = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    batches = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, device="cuda")
    torch.cuda.synchronize()
    nnz = indices.numel()
    eng = E.Engine(0)
    eng.set_matrix_csr(indptr, indices, data, N, G)
    eng.set_labels(labels, L)
    mat_max = np.float32(data.max().item())
    t0 = time.time()
    obs_tm = eng.observed_trimean(mat_max)
    print(f"{cfg}: nnz={nnz} observed trimean (incl. value-sorted CSC build) {1e3*(time.time()-t0):.1f} ms, nonzero {np.mean(obs_tm != 0):.3f}")
    t0 = time.time()
    obs_tm = eng.observed_trimean(mat_max)
    print(f"{cfg}: observed trimean again {1e3*(time.time()-t0):.1f} ms {eng.timings()['total_ms']:.2f} dev")
    rng = np.random.default_rng(0)
    lig, rec, src, tgt = rng.integers(0, G, T), rng.integers(0, G, T), rng.integers(0, L, T), rng.integers(0, L, T)
    prod = obs_tm[src, lig] * obs_tm[tgt, rec]
    obs = prod / (0.5 + prod)
    eng.set_triples(lig, rec, src, tgt)
    for ch in chunks:
        for pb in batches:
            eng.set_option("trimean_chunk_nnz", ch)
            eng.set_option("trimean_perm_batch", pb)
            for it in range(2):
                t0 = time.time()
                out = eng.run_trimean(P, mat_max, obs_prob=obs, seed=1337)
                wall = time.time() - t0
                tm, tt = eng.timings(), eng.trimean_timings()
            print(f"{cfg} chunk={ch} PB={tm['perm_batch']} P={P} T={T}: wall {wall*1e3:.1f} ms dev total {tm['total_ms']:.2f} "
                  f"(slab {tm['setup_ms']:.2f} count {tt['count_ms']:.2f} prefix {tt['prefix_ms']:.2f} select {tt['select_ms']:.2f} "
                  f"walked {tt['walked_warps']}/{tt['select_warps']} final {tt['finalize_ms']:.2f} score {tm['score_ms']:.2f}) upd/s(count) {P*nnz/tt['count_ms']*1e3:.3e} mean p {out['cellchat'].mean()/P:.4f}")
    # the cellphonedb FAST path on the same input for scale
    means, _, _ = eng.observed()
    o = ((means[src, lig] + means[tgt, rec]) / 2).astype(np.float32)
    eng.set_triples(lig, rec, src, tgt, o, None)
    for it in range(2):
        eng.run(P, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB)
    print(f"{cfg} cellphonedb FAST for scale: dev total {eng.timings()['total_ms']:.2f} ms")


main()
