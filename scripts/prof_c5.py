import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch, cProfile, pstats
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
genes = consensus_genes(); G=len(genes)
S,N,L=8,10_000,30
mats, labs, samp = [], [], []
for s in range(S):
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, seed_data=1337+2*s, seed_labels=1338+2*s)
    mats.append(sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))); labs.append(labels.cpu().numpy()); samp += [f"s{s:02d}"]*N
X = sparse.vstack(mats).tocsr()
obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in np.concatenate(labs)]), "sample": pd.Categorical(samp)})
adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(genes)))
li.mt.cellphonedb.by_sample(adata, sample_key="sample", groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
pr=cProfile.Profile(); pr.enable()
li.mt.cellphonedb.by_sample(adata, sample_key="sample", groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(32)
