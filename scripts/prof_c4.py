import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, pandas as pd, torch, cProfile, pstats
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
genes = consensus_genes(); G=len(genes)
N,L=200_000,40
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
adata = AnnDataLite(X=X, obs=pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels.cpu().numpy()])}), var=pd.DataFrame(index=pd.Index(genes)))
li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
pr=cProfile.Profile(); pr.enable()
li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
