"""Where the time of the plugin call goes at a BASELINE.json shape (full matrix on the host):
   LRPERM_DEBUG_TIMING=1 python scripts/plugin_profile.py [c3|c2|c4]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import liana_py_b200 as li

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda", 0)
adata = bench.host_adata(bench.gen_full(name, dev))
torch.cuda.empty_cache()
kw = dict(groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
for i in range(3):
    t0 = time.perf_counter(); res = li.mt.cellphonedb(adata, **kw); print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms ({len(res)} rows)", file=sys.stderr, flush=True)
pr = cProfile.Profile(); pr.enable()
res = li.mt.cellphonedb(adata, **kw)
pr.disable()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(35); print(out.getvalue())
