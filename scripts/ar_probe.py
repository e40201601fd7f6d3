import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
import bench
import liana_py_b200 as li
from liana_py_b200.method import _pipeline as PL
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = li.mt.DistContext.from_torch()
adata = bench.host_adata(bench.gen_full("c3", dev))
torch.cuda.empty_cache()
eng = PL.get_engine(local)
X = adata.X
from liana_py_b200.method._pre import _row_block
r0, r1 = ctx.row_block(X.shape[0], X.indptr)
def T(msg, t0):
    torch.cuda.synchronize()
    print(f"rank {ctx.rank} {msg}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); Xs = _row_block(X, False, r0, r1); T("row_block", t0)
    t0 = time.perf_counter(); st = eng.stage_matrix(Xs); T("stage", t0)
    a = np.concatenate([st["col_sums"], [1.0, 2.0, 3.0]])
    t0 = time.perf_counter(); t = torch.from_numpy(a.copy()).to(dev); T("to dev", t0)
    t0 = time.perf_counter(); dist.all_reduce(t); T("all_reduce", t0)
    t0 = time.perf_counter(); b = t.cpu().numpy(); T("cpu", t0)
    t0 = time.perf_counter(); dist.barrier(); T("barrier", t0)
dist.destroy_process_group()
