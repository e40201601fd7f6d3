"""Where the wall time of the user-facing call goes (configs[2] by default): cProfile of one warm
`li.mt.cellphonedb(adata, ...)` plus the engine's own stage timers (LRPERM_STAGE_TIMERS=1 prints them)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
method = getattr(li.mt, sys.argv[2] if len(sys.argv) > 2 else "cellphonedb")
N, G, L = {"c2": (50000, 1877, 20), "c3": (500000, 1877, 50), "c4": (200000, 1877, 40)}[cfg]
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels.cpu().numpy()])})
adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(consensus_genes())))
kw = dict(groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
for _ in range(2):
    t0 = time.perf_counter(); res = method(adata, **kw); print(f"warm-up call {1e3 * (time.perf_counter() - t0):.1f} ms, {len(res)} rows")
pr = cProfile.Profile()
pr.enable(); res = method(adata, **kw); pr.disable()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(int(os.environ.get('TOP', '45')))
print(out.getvalue()[:9000])
