"""cProfile of `li.mt.cellphonedb.by_sample` over S samples x 10k cells x 30 clusters (configs[4] shape, one GPU)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from scipy import sparse
import liana_py_b200 as li
from liana_py_b200.testing._anndata_lite import AnnDataLite
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

import torch
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 12
adata, _ = bench.by_sample_adata(torch.device("cuda", 0), n_samples=S)       # the bench's configs[4] input: 10k cells x 5k genes x 30 clusters per sample
kw = dict(sample_key="sample", groupby="cluster", use_raw=False, n_perms=1000, inplace=False)
for _ in range(2):
    t0 = time.perf_counter(); res = li.mt.cellphonedb.by_sample(adata, **kw)
    print(f"warm-up: {1e3 * (time.perf_counter() - t0):.0f} ms for {S} samples, {len(res)} rows")
pr = cProfile.Profile(); pr.enable(); res = li.mt.cellphonedb.by_sample(adata, **kw); pr.disable()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(int(os.environ.get("TOP", "40")))
print(out.getvalue()[:9000])
