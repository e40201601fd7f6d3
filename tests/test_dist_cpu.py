"""CPU (gloo, world_size 2): the host-side multi-rank logic - permutation sharding by global id,
the integer all-reduce of exceedance counts, sample sharding and frame gathering for by_sample.
The per-rank compute is stood in by the CPU oracle (the engine itself needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liana_py_b200.method._dist import DistContext


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden
        from oracle import c_oracle, numpy_port
        ctx = DistContext.from_torch()
        assert (ctx.rank, ctx.world_size) == (rank, world)
        g = load_golden("small300")
        P = int(g["n_perms"])
        lo, hi = ctx.perm_range(P)
        cube = c_oracle.cube(g["X"], g["labels"], g["L"], g["perm_idx"][lo:hi])
        idx = (g["ligand_idx"], g["receptor_idx"], g["source_idx"], g["target_idx"])
        obs = numpy_port.cpdb_observed(g["ligand_means"].copy(), g["receptor_means"].copy())
        local = numpy_port.counts_from_cube(cube, *idx, obs, "cellphonedb").astype(np.uint32)
        total = ctx.all_reduce_sum_u32(local)
        assert np.array_equal(total / P, g["cpdb_pvals"])          # identical to the single-process reference
        # CellChat null: same sharding, the oracle's trimean cube per rank, integer all-reduce
        from conftest import load_golden_cellchat
        gc = load_golden_cellchat("ragged90")
        Pc = int(gc["n_perms"])
        lo, hi = ctx.perm_range(Pc)
        perms = numpy_port.perm_indices(gc["X"].shape[0], Pc, int(gc["seed"]))[lo:hi]
        tm_cube = c_oracle.trimean_cube(gc["X"], gc["labels"], gc["L"], perms, gc["mat_max"])
        idx_c = (gc["ligand_idx"], gc["receptor_idx"], gc["source_idx"], gc["target_idx"])
        local_c = numpy_port.counts_cellchat(tm_cube, *idx_c, gc["lr_probs"]).astype(np.uint32)
        assert np.array_equal(ctx.all_reduce_sum_u32(local_c) / Pc, gc["pvals"])
        # sharded upload: every rank contributes 1/world of the flat CSR arrays, the all-gather rebuilds them everywhere
        cpu_ctx = DistContext(ctx.rank, ctx.world_size, torch.device("cpu"), None)
        X = g["X"]
        ip, ix, dt = cpu_ctx.upload_csr_sharded(X.indptr.astype(np.int32), X.indices.astype(np.int32), X.data.astype(np.float32))
        assert np.array_equal(ip.numpy(), X.indptr) and np.array_equal(ix.numpy(), X.indices) and np.array_equal(dt.numpy(), X.data)
        # by_sample: whole samples per rank, gathered as objects
        sizes = [50, 400, 30, 220, 90, 10, 300]
        mine = ctx.sample_shard(sizes)
        parts = ctx.all_gather_objects({i: sizes[i] * 2 for i in mine})
        merged = {}
        for p in parts:
            assert not (set(p) & set(merged))
            merged.update(p)
        assert merged == {i: s * 2 for i, s in enumerate(sizes)}
        loads = [sum(sizes[i] for i in q) for q in (ctx.all_gather_objects(mine))]
        assert max(loads) - min(loads) <= max(sizes)
        # the drop-in methods themselves under the 2-rank context (the engine stood in by the oracle, tests/_oracle_engine.py):
        # permutations sharded by global id + integer all-reduce -> the reference's frame on every rank; by_sample with
        # whole samples per rank, gathered (every rank holds all samples) or not (every rank keeps its own block)
        import pandas as pd
        import liana_py_b200 as li
        from liana_py_b200.method import _pipeline as PL
        from _oracle_engine import OracleEngine
        import test_dropin_gpu as G
        PL._ENGINES[0] = OracleEngine()
        adata, Pn, seed, ep = G.load_input("small300")
        kw = dict(groupby="cluster", use_raw=False, n_perms=Pn, seed=seed, expr_prop=ep, inplace=False, perm_source="reference")
        res = li.mt.cellphonedb(adata, dist=ctx, **kw)
        gold = pd.read_csv(os.path.join(G.GOLDEN, "small300_cellphonedb.csv.gz"), float_precision="round_trip")
        G.compare(res, gold, "lr_means", "cellphone_pvals")
        cc = li.mt.cellchat(adata, dist=ctx, **kw)
        goldc = pd.read_csv(os.path.join(G.GOLDEN, "small300_cellchat.csv.gz"), float_precision="round_trip")
        a, b = cc.sort_values(G.KEY).reset_index(drop=True), goldc.sort_values(G.KEY).reset_index(drop=True)
        assert np.array_equal(a["cellchat_pvals"].values, b["cellchat_pvals"].values)
        from variants import sample_inputs
        kws = dict(sample_key="sample", groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                   perm_source="reference", use_raw=False)
        golds = pd.read_csv(os.path.join(G.GOLDEN, "variant_bysample_cellphonedb.csv.gz"), float_precision="round_trip")
        full = li.mt.cellphonedb.by_sample(sample_inputs(), dist=ctx, **kws)
        assert len(full) == len(golds) and list(dict.fromkeys(full["sample"])) == ["s1", "s2"]
        for s_name in ("s1", "s2"):
            G.compare(full[full["sample"] == s_name].drop(columns="sample"),
                      golds[golds["sample"] == s_name].drop(columns="sample"), "lr_means", "cellphone_pvals")
        own = li.mt.cellphonedb.by_sample(sample_inputs(), dist=ctx, gather=False, **kws)
        mine_s = set(own["sample"])
        assert len(mine_s) == 1                                     # two samples, two ranks: one whole sample each
        assert set().union(*ctx.all_gather_objects(mine_s)) == {"s1", "s2"}
        G.compare(own.drop(columns="sample"), golds[golds["sample"] == next(iter(mine_s))].drop(columns="sample"),
                  "lr_means", "cellphone_pvals")
        t = torch.full((4,), rank + 1, dtype=torch.int32)
        ctx.all_reduce_sum_(t)
        assert t.tolist() == [sum(range(1, world + 1))] * 4
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world = 2
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ["PYTHONPATH"] = here + os.pathsep + os.path.dirname(here) + os.pathsep + os.environ.get("PYTHONPATH", "")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _worker_by_sample(rank, world, port, out_dir):
    """3 ranks, 2 samples: one rank owns no sample.  The frame gathered from numeric columns must equal the frame one
    process builds (same rows, order, column dtypes and values)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pandas as pd
        import liana_py_b200 as li
        from liana_py_b200.method import _pipeline as PL
        from _oracle_engine import OracleEngine
        from variants import sample_inputs
        PL._ENGINES[0] = OracleEngine()
        ctx = DistContext.from_torch()
        for method, extra in ((li.mt.cellphonedb, {}), (li.mt.rank_aggregate, {}), (li.mt.cellphonedb, {"return_all_lrs": True})):
            kws = dict(sample_key="sample", groupby="cluster", n_perms=8, seed=7, expr_prop=0.1, inplace=False,
                       perm_source="reference", use_raw=False, **extra)
            single = method.by_sample(sample_inputs(), **kws)
            full = method.by_sample(sample_inputs(), dist=ctx, **kws)
            pd.testing.assert_frame_equal(full, single)
            at0 = method.by_sample(sample_inputs(), dist=ctx, gather=0, **kws)       # only rank 0 assembles the frame
            assert (at0 is None) == (rank != 0)
            if rank == 0:
                pd.testing.assert_frame_equal(at0, single)
            own = method.by_sample(sample_inputs(), dist=ctx, gather=False, **kws)
            n_own = ctx.all_gather_objects(len(own))
            assert sum(n_own) == len(single) and min(n_own) == 0          # a rank without samples returns an empty frame
            if len(own):
                ref = single[single["sample"].isin(set(own["sample"]))].reset_index(drop=True)
                pd.testing.assert_frame_equal(own, ref)
        # all_gather_rows: ragged, empty and bool columns
        cols = {"a": np.arange(rank * 3, dtype=np.int32), "b": np.linspace(0, 1, rank * 3), "c": np.arange(rank * 3) % 2 == 0}
        got = ctx.all_gather_rows(cols)
        want_a = np.concatenate([np.arange(r * 3, dtype=np.int32) for r in range(world)])
        assert np.array_equal(got["a"], want_a) and got["a"].dtype == np.int32
        assert got["b"].dtype == np.float64 and got["c"].dtype == np.bool_ and len(got["c"]) == len(want_a)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_three_rank_by_sample_numeric_gather(tmp_path):
    world = 3
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ["PYTHONPATH"] = here + os.pathsep + os.path.dirname(here) + os.pathsep + os.environ.get("PYTHONPATH", "")
    mp.spawn(_worker_by_sample, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_perm_ranges_partition():
    for P in (1, 7, 1000, 1001):
        for W in (1, 2, 3, 8):
            seen = []
            for r in range(W):
                lo, hi = DistContext(r, W).perm_range(P)
                seen.extend(range(lo, hi))
            assert seen == list(range(P))


def test_single_rank_context_is_identity():
    ctx = DistContext.from_torch()
    assert ctx.world_size == 1
    a = np.arange(5, dtype=np.uint32)
    assert ctx.all_reduce_sum_u32(a) is a
    assert ctx.sample_shard([3, 1, 2]) == [0, 1, 2]
