"""BASELINE.json configs[3] and [4] through the drop-in API (wall clock):
  c4: custom AggregateClass (cellphonedb + geometric_mean nulls + connectome/log2FC/NATMI/SCA) on 200k cells x 40 clusters
  c5: li.mt.cellphonedb.by_sample over S samples x 10k cells x 30 clusters (samples sharded over ranks under torchrun)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

which = sys.argv[1]
genes = consensus_genes()
G = len(genes)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as td
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist = li.mt.DistContext.from_torch()

def make(N, L, seed):
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, seed_data=seed, seed_labels=seed + 1, device=f"cuda:{local}")
    X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
    return X, labels.cpu().numpy()

if which == "c4":
    N, L = 200_000, 40
    X, labels = make(N, L, 1337)
    adata = AnnDataLite(X=X, obs=pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels])}),
                        var=pd.DataFrame(index=pd.Index(genes)))
    custom = li.mt.AggregateClass(li.mt.aggregate_meta, methods=[li.mt.cellphonedb, li.mt.geometric_mean, li.mt.connectome,
                                                                  li.mt.logfc, li.mt.natmi, li.mt.singlecellsignalr])
    for name, fn in (("rank_aggregate(default)", li.mt.rank_aggregate), ("AggregateClass(cpdb+gmean+4)", custom)):
        for it in range(2):
            t0 = time.perf_counter()
            res = fn(adata, groupby="cluster", use_raw=False, n_perms=1000, inplace=False, device=local, dist=dist)
            dt = time.perf_counter() - t0
            if rank == 0:
                print(f"c4 {name} call {it}: {dt*1e3:.0f} ms, rows {len(res)}, cols {list(res.columns)[4:]}", flush=True)
else:
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    N, L = 10_000, 30
    mats, labs, samp = [], [], []
    for s in range(S):
        X, labels = make(N, L, 1337 + 2 * s)
        mats.append(X); labs.append(labels); samp += [f"s{s:02d}"] * N
    X = sparse.vstack(mats).tocsr()
    obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in np.concatenate(labs)]), "sample": pd.Categorical(samp)})
    adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(genes)))
    for gather in ((True, False) if world > 1 else (True,)):
        for it in range(2):
            if world > 1:
                td.barrier()
            t0 = time.perf_counter()
            res = li.mt.cellphonedb.by_sample(adata, sample_key="sample", groupby="cluster", use_raw=False, n_perms=1000,
                                              inplace=False, device=local, dist=dist, **({} if world == 1 else {"gather": gather}))
            if world > 1:
                td.barrier()
            dt = time.perf_counter() - t0
            if rank == 0:
                print(f"c5 by_sample {S} samples x {N} cells, world {world}, gather={gather}, call {it}: {dt:.2f} s, rows on rank 0 {len(res)}", flush=True)
if world > 1:
    td.destroy_process_group()
