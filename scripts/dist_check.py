"""Multi-GPU check of the drop-in layer (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node 2 scripts/dist_check.py
Every method called with dist=DistContext.from_torch() must return, on every rank, exactly the frame a single GPU
returns (permutations are keyed by global id, counts are summed as integers; by_sample shards whole samples)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch, torch.distributed as dist
import liana_py_b200 as li
from liana_py_b200.testing._synth import make_synthetic_adata

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = li.mt.DistContext.from_torch()
KEY = ["source", "target", "ligand_complex", "receptor_complex"]


def same(a, b):
    a = a.sort_values(KEY + [c for c in a.columns if c == "sample"]).reset_index(drop=True)
    b = b.sort_values(KEY + [c for c in b.columns if c == "sample"]).reset_index(drop=True)
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    for c in a.columns:
        if a[c].dtype.kind == "f":
            assert np.array_equal(a[c].values, b[c].values, equal_nan=True), c
        else:
            assert (a[c].astype(str).values == b[c].astype(str).values).all(), c


adata = make_synthetic_adata(6000, 900, 8, density=0.2, seed_data=5, seed_labels=6, n_samples=5)
kw = dict(groupby="cluster", use_raw=False, n_perms=300, seed=11, inplace=False, device=local)
for name in ("cellphonedb", "geometric_mean", "cellchat", "rank_aggregate"):
    fn = getattr(li.mt, name)
    t0 = time.time()
    one = fn(adata, **kw)
    many = fn(adata, dist=ctx, **kw)
    same(one, many)
    if ctx.rank == 0:
        print(f"{name}: {ctx.world_size} ranks == 1 rank ({len(one)} rows, {time.time() - t0:.2f} s)", flush=True)
sk = [c for c in adata.obs.columns if c != "cluster"][0]
for name in ("cellphonedb", "rank_aggregate"):
    fn = getattr(li.mt, name)
    one = fn.by_sample(adata, sk, **kw)
    many = fn.by_sample(adata, sk, dist=ctx, **kw)          # numeric columns all-gathered, names built on the consumer
    pd.testing.assert_frame_equal(many, one)                 # same rows, same ORDER, same dtypes
    own = fn.by_sample(adata, sk, dist=ctx, gather=False, **kw)
    assert sum(ctx.all_gather_objects(len(own))) == len(one)
    if ctx.rank == 0:
        print(f"{name}.by_sample over {adata.obs[sk].nunique()} samples: {ctx.world_size} ranks == 1 rank ({len(one)} rows)", flush=True)
# the same gather with every column going through the chunked pinned device -> host copy (gigabyte-sized frames take it by default)
from liana_py_b200.method import _dist as _D
_D.D2H_CHUNKED_MIN_BYTES = 1
for g in (True, 0):
    many = li.mt.cellphonedb.by_sample(adata, sk, dist=ctx, gather=g, **kw)
    if many is not None:
        pd.testing.assert_frame_equal(many, li.mt.cellphonedb.by_sample(adata, sk, **kw))
    else:
        li.mt.cellphonedb.by_sample(adata, sk, **kw)
if ctx.rank == 0:
    print("by_sample gather through the chunked device -> host copy: ok", flush=True)
_D.D2H_CHUNKED_MIN_BYTES = 8 << 20
# dense input and a min_cells filter through the sharded staging
dense = make_synthetic_adata(3000, 400, 6, density=0.3, seed_data=7, seed_labels=8)
dense.X = np.asarray(dense.X.todense(), dtype=np.float32)
same(li.mt.cellphonedb(dense, min_cells=400, **kw), li.mt.cellphonedb(dense, min_cells=400, dist=ctx, **kw))
if ctx.rank == 0:
    print("dense input + min_cells through the sharded staging: ok", flush=True)
dist.barrier()
dist.destroy_process_group()
