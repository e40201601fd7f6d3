"""A/B of the staging pass: time of `stage_matrix_device` (device -> device copy + `stage_stats_kernel`) on a resident
200k x 30k matrix, for the library named by LRPERM_LIB (default: the in-tree build).
   python scripts/stage_stats_ab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from liana_py_b200 import _engine as E
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

N, G = 200_000, 30_000
ip, ix, dt, _ = synthetic_csr_torch(N, G, 40, 0.08)
eng = E.Engine(0)
ts = []
for _ in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = eng.stage_matrix_device(ip, ix, dt, (N, G))
    ts.append(time.perf_counter() - t0)
print(f"{os.environ.get('LRPERM_LIB', 'in-tree')}: nnz={int(ix.numel())} n_unordered={st['n_unordered']} "
      f"stage_matrix_device best {min(ts) * 1e3:.2f} ms median {sorted(ts)[len(ts) // 2] * 1e3:.2f} ms", flush=True)
