"""Wall-clock breakdown of the user-facing call li.mt.cellphonedb(adata, ...) on a synthetic AnnData-like
input of a BASELINE.json shape (genes: the resource's 1,877 symbols + `extra` background genes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.method import _pipeline as PL, _tables, _methods
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0
method = sys.argv[3] if len(sys.argv) > 3 else "cellphonedb"
N, L = {"c2": (50000, 20), "c3": (500000, 50), "c4": (200000, 40), "c5": (10000, 30)}[cfg]
genes = consensus_genes()
G = len(genes) + extra
names = np.concatenate([genes, np.array([f"BG{i:05d}" for i in range(extra)])]) if extra else genes
t0 = time.perf_counter()
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels.cpu().numpy()])})
adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(names)))
print(f"input built in {time.perf_counter()-t0:.1f}s: {N} x {G}, nnz {X.nnz}")
fn = getattr(li.mt, method)
for it in range(3):
    t0 = time.perf_counter()
    res = fn(adata, groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
    dt = time.perf_counter() - t0
    print(f"{cfg} {method} call {it}: {dt*1e3:.0f} ms, rows {len(res)}")
# breakdown of the last configuration
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
res = fn(adata, groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
