"""by_sample under torchrun: stage timeline per rank (LRPERM_DEBUG_TIMING=1).
   LRPERM_DEBUG_TIMING=1 torchrun --nproc-per-node N scripts/bysample_dist_probe.py [n_samples]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
import liana_py_b200 as li
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = li.mt.DistContext.from_torch()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
adata, _ = bench.by_sample_adata(dev, n_samples=S)
kw = dict(groupby="cluster", use_raw=False, n_perms=1000, inplace=False, device=local, dist=ctx)
for i in range(2):
    dist.barrier(); t0 = time.perf_counter()
    res = li.mt.cellphonedb.by_sample(adata, "sample", **kw)
    print(f"rank {ctx.rank} call {i}: {(time.perf_counter() - t0) * 1e3:.0f} ms ({len(res)} rows)", file=sys.stderr, flush=True)
dist.barrier()
dist.destroy_process_group()
