"""Single-call multi-GPU (`devices=[...]`, one process, several GPUs) against one GPU:
   python scripts/multidev_check.py [n_devices]
Every method must return exactly the frame one GPU returns (permutations keyed by global id, integer counts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import liana_py_b200 as li
from liana_py_b200.testing._synth import make_synthetic_adata

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
devices = list(range(n))
assert n >= 2, "needs at least 2 GPUs"
adata = make_synthetic_adata(20000, 2500, 10, density=0.15, seed_data=5, seed_labels=6, n_samples=6)
kw = dict(groupby="cluster", use_raw=False, n_perms=1000, seed=11, inplace=False)
for name in ("cellphonedb", "geometric_mean", "cellchat", "rank_aggregate"):
    fn = getattr(li.mt, name)
    one = fn(adata, **kw)
    t0 = time.perf_counter(); one = fn(adata, **kw); t1 = time.perf_counter() - t0
    many = fn(adata, devices=devices, **kw)
    t0 = time.perf_counter(); many = fn(adata, devices=devices, **kw); tn = time.perf_counter() - t0
    pd.testing.assert_frame_equal(many, one)
    print(f"{name}: devices={devices} == 1 GPU ({len(one)} rows; {t1 * 1e3:.1f} ms -> {tn * 1e3:.1f} ms)", flush=True)
ref = li.mt.cellphonedb(adata, perm_source="reference", n_perms=64, **{k: v for k, v in kw.items() if k != "n_perms"})
got = li.mt.cellphonedb(adata, perm_source="reference", n_perms=64, devices=devices, **{k: v for k, v in kw.items() if k != "n_perms"})
pd.testing.assert_frame_equal(got, ref)
one = li.mt.cellphonedb.by_sample(adata, "sample", **kw)
t0 = time.perf_counter(); one = li.mt.cellphonedb.by_sample(adata, "sample", **kw); t1 = time.perf_counter() - t0
t0 = time.perf_counter(); many = li.mt.cellphonedb.by_sample(adata, "sample", devices=devices, **kw); tn = time.perf_counter() - t0
pd.testing.assert_frame_equal(many, one)
print(f"by_sample: devices={devices} == 1 GPU ({len(one)} rows; {t1 * 1e3:.1f} ms -> {tn * 1e3:.1f} ms)", flush=True)
