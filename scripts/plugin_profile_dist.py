"""Where the time of the plugin call goes under torchrun (dist=DistContext): rank 0 prints a cProfile.
   LRPERM_DEBUG_TIMING=1 torchrun --nproc-per-node N scripts/plugin_profile_dist.py [c3|c2|c4]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
import liana_py_b200 as li

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
print("numa bound:", li.mt.bind_to_gpu_numa(local) if os.environ.get("BIND_NUMA", "1") == "1" else "off", file=sys.stderr)
dist.init_process_group("nccl", device_id=dev)
ctx = li.mt.DistContext.from_torch()
adata = bench.host_adata(bench.gen_full(name, dev))
torch.cuda.empty_cache()
kw = dict(groupby="cluster", use_raw=False, n_perms=1000, inplace=False, device=local, dist=ctx)
for i in range(3):
    dist.barrier()
    t0 = time.perf_counter(); res = li.mt.cellphonedb(adata, **kw)
    print(f"rank {ctx.rank} call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms ({len(res)} rows)", file=sys.stderr, flush=True)
dist.barrier()
pr = cProfile.Profile(); pr.enable()
res = li.mt.cellphonedb(adata, **kw)
pr.disable()
if ctx.rank == 0:
    out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45); print(out.getvalue())
dist.barrier()
dist.destroy_process_group()
