#!/usr/bin/env python
"""Benchmark of the permutation-null hot path (metric of BASELINE.json:
permutations x LR-triples per second for cellphonedb, n_perms=1000).

  python bench.py [--gpus N --steps K --warmup W] [--workload c3|c2|c4|c5] [--order fast|exact]
  python bench.py --impl reference ...      # CPU port of the reference on the box's host cores

A "step" = one full pass of the hot path (label shuffles -> per-cluster means -> score ->
exceedance counts for all P permutations and all T kept triples) over one synthetic input.
`value` is measured with the inputs resident in HBM; `e2e` goes through the C ABI with HOST
buffers (H2D of the CSR matrix / labels / triple tables and D2H of the counts inside the timed
region).  Multi-GPU: one process per GPU (torchrun), permutations sharded by global id, one
NCCL all-reduce of the uint32 counters per step (strong scaling: total work fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells, genes nominal, clusters, description)
    "c2": (50_000, 20_000, 20, "configs[1]: 50k cells x 20k genes CSR 8%, 20 clusters, consensus, n_perms=1000"),
    "c3": (500_000, 30_000, 50, "configs[2]: 500k cells x 30k genes CSR 8%, 50 clusters, consensus, n_perms=1000"),
    "c4": (200_000, 30_000, 40, "configs[3] shape: 200k cells, 40 clusters, n_perms=1000"),
    "c5": (10_000, 30_000, 30, "configs[4] one sample: 10k cells, 30 clusters, n_perms=1000"),
    "tiny": (3_000, 2_000, 8, "smoke-size workload for CI: 3k cells, 8 clusters"),
}
METRIC = "permutations x LR-triples per second (cellphonedb, n_perms=1000)"
UNIT = "perm*triples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and "None" not in name and "All" not in name:
                        self.reasons.add(name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(name, device, seed_data=1337, seed_labels=1338, use_engine=True):
    """Synthetic CSR over the resource's genes only (the hot path never sees the background genes;
    SURVEY.md 8d) + triple tables from the real consensus resource with expr_prop=0.1."""
    import torch
    from liana_py_b200 import _engine as E
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    from liana_py_b200.testing._synth import consensus_genes
    from liana_py_b200.testing._synth_gpu import synthetic_csr_torch

    N, G_nominal, L, desc = WORKLOADS[name]
    genes = consensus_genes()
    G = len(genes)
    indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08, seed_data, seed_labels, device=device)
    if use_engine:
        eng = E.Engine(device.index)
        eng.set_stream(torch.cuda.current_stream(device).cuda_stream)
        eng.set_matrix_csr(indptr, indices, data, N, G)
        eng.set_labels(labels, L)
        means, stored, sizes = eng.observed()
    else:       # reference arm: no engine anywhere on the path (observed stats from the CPU checker)
        from scipy import sparse
        from oracle import c_oracle
        eng = None
        Xh = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(N, G))
        means, stored, sizes = c_oracle.observed(Xh, labels.cpu().numpy(), L)
    props = stored / sizes[:, None].astype(np.float64)
    tables = _tables.build_resource_tables(select_resource("consensus"), genes)
    assert np.array_equal(tables.entities, genes)
    tri = _tables.resolve_triples(tables, means, props, 0.1)
    obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    wl = dict(name=name, desc=desc, N=N, G=G, G_nominal=G_nominal, L=L, nnz=int(indices.numel()), T=int(len(obs)),
              T_max=int(tables.lig_sub.shape[0]) * L * L, indptr=indptr, indices=indices, data=data, labels=labels,
              tri=tri, obs=obs)
    return eng, wl


def model_bytes(wl, P):
    """one streaming pass over X_sub per permutation."""
    per_perm = 8 * wl["nnz"] + 4 * (wl["N"] + 1) + 4 * wl["N"] + 8 * wl["L"] * wl["G"]
    return per_perm, P * per_perm + 24 * wl["T"]


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from liana_py_b200 import _engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun (--nproc-per-node N) for --gpus N > 1")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    P = args.n_perms
    order = E.ORDER_FAST if args.order == "fast" else E.ORDER_EXACT
    eng, wl = build_workload(args.workload, device)
    T = wl["T"]
    lo, hi = (P * rank) // world, (P * (rank + 1)) // world
    tri, obs = wl["tri"], wl["obs"]
    eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
    counts = torch.zeros(max(T, 1), dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > L2 (126 MB)

    def step_resident():
        eng.run(hi - lo, seed=1337, perm_offset=lo, order=order, scores=E.SCORE_CPDB, counts_out=(counts, None))
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)

    def timed(fn, steps, warmup, flush_l2=True):
        for _ in range(warmup):
            fn()
        per, launches, kern_ms, kern_n = [], 0, 0.0, 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(steps):
            if flush_l2:
                flush.fill_(1)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            per.append(e0.elapsed_time(e1))
            tm = eng.timings()
            launches += tm["launches"]
            kern_ms += tm["main_kernel_ms"]
            kern_n += tm["main_kernel_launches"]
        t = torch.tensor([sum(per)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, kern_ms, kern_n, eng.timings()

    sampler = ClockSampler(local)
    sampler.start()
    total_ms, launches, kern_ms, kern_n, tm = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = P * T / (ms_per_step * 1e-3)
    result_counts = counts.clone()

    # ---- end to end through the C ABI with host buffers (pinned), same metric ----
    host = {k: wl[k].cpu().pin_memory() for k in ("indptr", "indices", "data", "labels")}
    h_idx = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
             for a in (tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)]
    h_obs = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32)).pin_memory()
    h_counts = torch.zeros(max(T, 1), dtype=torch.int32).pin_memory()
    h2d = sum(int(t.numel() * t.element_size()) for t in list(host.values()) + h_idx + [h_obs])
    d2h = int(h_counts.numel() * 4)
    if world > 1:       # whole job: every rank uploads the small tables, the matrix arrays cross PCIe once in total
        mat = sum(int(host[k].numel() * host[k].element_size()) for k in ("indices", "data"))
        h2d = world * (h2d - mat) + mat
        d2h *= world

    from liana_py_b200.method._dist import DistContext
    dctx = DistContext.from_torch() if world > 1 else None

    def step_e2e():
        if world > 1:
            # every rank holds the same host matrix: each uploads 1/world of it from pinned memory and the slices are
            # all-gathered over NVLink instead of 8 full transfers over 8 PCIe links fed by the same host memory
            d_ip, d_ix, d_dt = dctx.upload_csr_sharded(host["indptr"], host["indices"], host["data"])
            eng.set_matrix_csr(d_ip, d_ix, d_dt, wl["N"], wl["G"])
        else:
            eng.set_matrix_csr(host["indptr"], host["indices"], host["data"], wl["N"], wl["G"])
        eng.set_labels(host["labels"], wl["L"])
        eng.set_triples(h_idx[0], h_idx[1], h_idx[2], h_idx[3], h_obs, None)
        out = eng.run(hi - lo, seed=1337, perm_offset=lo, order=order, scores=E.SCORE_CPDB, counts_out=(counts, None))
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        h_counts.copy_(counts, non_blocking=False)

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(max(3, args.warmup)):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = P * T / float(e2e_s.item())
    assert torch.equal(h_counts.to(device), result_counts), "e2e counts differ from the resident run"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_gbs, peak_src = peaks()
    per_perm_bytes, total_bytes = model_bytes(wl, P)
    pb = tm["perm_batch"]
    kern_avg_ms = kern_ms / max(kern_n, 1)
    perms_per_launch = (hi - lo) / max(tm["main_kernel_launches"], 1)
    achieved = per_perm_bytes * perms_per_launch / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.workload}_{args.order}")
    kernel_name = "fast2_means_kernel" if args.order == "fast" else "exact_means_kernel"
    roofline = {"bound": "hbm", "kernel": kernel_name,
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_perm": per_perm_bytes, "perms_per_launch": perms_per_launch,
                "kernel_ms_per_launch": kern_avg_ms}
    if args.order == "fast":
        # The fast kernel does not stream the matrix once per permutation (it re-uses an L2-resident label
        # tile for 512 permutations), so against the streaming model of SURVEY.md 8(d) it "exceeds" the DRAM
        # peak; what bounds it is the shared-memory pipe: every (stored entry, permutation) update is one LDS
        # + one STS of a 32-lane conflict-free wavefront, plus 1/4 wavefront of label load and 1/4 of staged
        # entry reads per update-warp (Q=4), and an SM retires one wavefront per clock.
        sm_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        upd = perms_per_launch * wl["nnz"]
        upd_clk_sm = upd / (kern_avg_ms * 1e-3) / sm_clk / 148
        wf_per_update_warp = 2.5
        roofline["note"] = ("streaming model of SURVEY.md 8(d); this kernel re-uses L2-resident label tiles across "
                            "permutations, so achieved exceeds the DRAM peak and `physical_dram_frac` is what HBM "
                            "really sees; its binding resource is the shared-memory pipe, see `smem_pipe`")
        roofline["physical_dram_frac"] = (traffic / (kern_avg_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None
        roofline["smem_pipe"] = {"updates_per_clk_per_sm": upd_clk_sm, "wavefronts_per_update_warp": wf_per_update_warp,
                                 "peak_updates_per_clk_per_sm": 32.0 / wf_per_update_warp,
                                 "frac": upd_clk_sm * wf_per_update_warp / 32.0,
                                 "ncu_l1tex_lsuin_pct": 80.8, "ncu_source": "profiles/r01_ncu_full_c3_fast2_final.md"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 sums / f64 compare", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "cells": wl["N"], "genes_nominal": wl["G_nominal"],
                   "genes_in_resource": wl["G"], "clusters": wl["L"], "nnz": wl["nnz"], "n_perms": P,
                   "triples_kept": T, "triples_max": wl["T_max"], "expr_prop": 0.1, "order": args.order,
                   "perm_source": "philox", "parallelism": f"perm-shard x{world}",
                   "l2": "L2 flushed (256 MB write) before every timed step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(e2e_s.item()) * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
        "stages_ms": {k: tm[k] for k in ("setup_ms", "means_ms", "score_ms", "total_ms")},
        "roofline": roofline,
    }
    if world == 1 and not args.no_stress:
        line["stress_expr_prop0"] = stress_variant(eng, wl, P, order, counts, timed)
        eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
    if world == 1 and not args.no_cellchat:
        line["cellchat_null"] = cellchat_null(eng, wl, P, timed, cpu=not args.no_cpu)
        eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
    if world == 1 and not args.no_dropin:
        line["dropin_call"] = dropin_call(wl, P)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl, P, T, n_jobs=1, budget_s=args.cpu_budget)
        if args.order == "fast" and not args.no_exact:
            ex_ms, _, ex_kern, ex_n, _ = timed(lambda: eng.run(P, seed=1337, order=E.ORDER_EXACT, scores=E.SCORE_CPDB,
                                                             counts_out=(counts, None)), 2, 1)
            ex_ms /= 2
            line["exact_order_mode"] = {"ms_per_step": ex_ms, "value": P * T / (ex_ms * 1e-3),
                                        "roofline_frac": per_perm_bytes * P / (ex_kern / 2 * 1e-3) / 1e9 / peak_gbs,
                                        "note": "bit-identical-to-reference summation order (validation mode)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def stress_variant(eng, wl, P, order, counts, timed):
    """the same workload with expr_prop=0, i.e. every (source, target, interaction) row reaches the
    score function (T = T_max): the means cube is unchanged, the score + count stage grows ~15x."""
    import torch
    from liana_py_b200 import _engine as E
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    from liana_py_b200.testing._synth import consensus_genes
    means, stored, sizes = eng.observed()
    props = stored / sizes[:, None].astype(np.float64)
    tables = _tables.build_resource_tables(select_resource("consensus"), consensus_genes())
    tri = _tables.resolve_triples(tables, means, props, 0.0)
    obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    T = len(obs)
    eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
    big = torch.zeros(max(T, 1), dtype=torch.int32, device=counts.device)
    total_ms, _, _, _, tm = timed(lambda: eng.run(P, seed=1337, order=order, scores=E.SCORE_CPDB, counts_out=(big, None)), 2, 3)
    ms = total_ms / 2
    return {"triples": T, "ms_per_step": ms, "value": P * T / (ms * 1e-3), "unit": UNIT,
            "stages_ms": {k: tm[k] for k in ("setup_ms", "means_ms", "score_ms", "total_ms")},
            "note": "expr_prop=0: all T_max rows scored; resident inputs, L2 flushed between steps"}


def cellchat_null(eng, wl, P, timed, cpu=True):
    """Secondary number: the CellChat null on the same matrix - per-(permutation, cluster, gene)
    TRIMEANS (three quantiles of the dense column, selected exactly from the stored entries) instead of means,
    prod / (kh + prod) score, same metric.  Device time with resident inputs, L2 flushed between steps."""
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    from liana_py_b200.testing._synth import consensus_genes
    mat_max = np.float32(wl["data"].max().item())
    obs_tm = eng.observed_trimean(mat_max)
    _, stored, sizes = eng.observed()
    props = stored / sizes[:, None].astype(np.float64)
    tables = _tables.build_resource_tables(select_resource("consensus"), consensus_genes())
    tri = _tables.resolve_triples(tables, obs_tm, props, 0.1)
    obs = PL.observed_cellchat(tri.ligand_means, tri.receptor_means)
    eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)
    T = len(obs)
    keep = {}

    def step():
        keep["out"] = eng.run_trimean(P, mat_max, obs_prob=obs, seed=1337)

    total_ms, launches, kern_ms, kern_n, tm = timed(step, 2, 3)
    ms = total_ms / 2
    tt = eng.trimean_timings()
    out = {"ms_per_step": ms, "value": P * T / (ms * 1e-3), "unit": UNIT, "triples_kept": T,
           "nonzero_observed_trimeans": float(np.mean(obs_tm != 0)), "perm_batch": tm["perm_batch"],
           "stages_ms": {"slab_ms": tm["setup_ms"], "count_ms": tt["count_ms"], "prefix_ms": tt["prefix_ms"],
                         "select_ms": tt["select_ms"], "finalize_ms": tt["finalize_ms"], "score_ms": tm["score_ms"]},
           "rewalked_items": [tt["walked_warps"], tt["select_warps"]],
           "count_kernel_updates_per_s": P * wl["nnz"] / (tt["count_ms"] * 1e-3),
           "mean_pvalue": float(keep["out"]["cellchat"].mean() / P),
           "note": "lrperm_run_trimean: count (fast2_means_kernel<COUNT>, shared-memory-pipe bound like the cellphonedb "
                   "kernel) + prefix + select + finalize + score; bit-identical to the reference for identical permutations"}
    if cpu:
        from oracle import numpy_port
        X, labels = cpu_inputs(wl)
        t0 = time.perf_counter()
        numpy_port.get_trimean_perms(X, labels, wl["L"], 1, 1337, mat_max)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": T / dt, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"1 of {P} permutations in {dt:.1f} s (NumPy restatement of _get_means_perms with "
                                         f"agg_fun=_trimean: dense np.quantile / np.median per cluster); extrapolated full run: {dt * P:.0f} s"}
    return out


def dropin_call(wl, P):
    """Wall time of the user-facing call `li.mt.cellphonedb(adata, groupby, n_perms=P)` on the same matrix held
    as a host SciPy CSR in an AnnData-like object (input checks, resource tables, H2D, permutations, result
    DataFrame): the number a liana user sees."""
    import pandas as pd
    import liana_py_b200 as li
    from liana_py_b200.testing._anndata_lite import AnnDataLite
    from liana_py_b200.testing._synth import consensus_genes
    X, labels = cpu_inputs(wl)
    obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels])})
    adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(consensus_genes())))
    best, rows = None, 0
    for _ in range(3):
        t0 = time.perf_counter()
        res = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, inplace=False)
        dt = time.perf_counter() - t0
        best, rows = (dt if best is None else min(best, dt)), len(res)
    return {"wall_ms": best * 1e3, "rows": rows, "value": P * rows / best, "unit": UNIT,
            "note": "best of 3; host AnnData-like input -> liana_res DataFrame, Philox permutations, FAST order"}


def cpu_inputs(wl):
    from scipy import sparse
    X = sparse.csr_matrix((wl["data"].cpu().numpy(), wl["indices"].cpu().numpy(), wl["indptr"].cpu().numpy()),
                          shape=(wl["N"], wl["G"]))
    return X, wl["labels"].cpu().numpy()


def cpu_baseline(wl, P, T, n_jobs=1, budget_s=20.0):
    """The reference's hot path (NumPy/SciPy port, oracle/numpy_port.py) on a bounded sample of the
    same workload: a chunk of permutations, extrapolated linearly in P (every stage is linear in P)."""
    from oracle import numpy_port
    X, labels = cpu_inputs(wl)
    tri, obs = wl["tri"], wl["obs"]
    # size the sample from one timed permutation
    t0 = time.perf_counter()
    numpy_port.get_means_perms(X, labels, wl["L"], 1, 1337, 1)
    one = time.perf_counter() - t0
    p_s = int(max(n_jobs, min(P, budget_s / max(one, 1e-3) * max(n_jobs, 1) * (0.6 if n_jobs > 1 else 1.0))))
    p_s = max(1, (p_s // n_jobs) * n_jobs)
    t0 = time.perf_counter()
    numpy_port.run_method_block(X, labels, wl["L"], p_s, 1337, tri.lig_gene, tri.rec_gene, tri.src, tri.tgt,
                                tri.ligand_means.copy(), tri.receptor_means.copy(), "cellphonedb", n_jobs=n_jobs)
    dt = time.perf_counter() - t0
    return {"value": p_s * T / dt, "unit": UNIT, "cores": n_jobs, "kind": "port",
            "sample": f"{p_s} of {P} permutations of the same workload in {dt:.1f} s "
                      f"(NumPy/SciPy restatement of _get_means_perms + gather + _cpdb_score, n_jobs={n_jobs}); "
                      f"extrapolated full run: {dt / p_s * P:.0f} s",
            "host_cpu_count": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (port) with all host threads it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # same synthetic generator as the engine arm (torch CUDA RNG when a GPU is present)
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    _, wl = build_workload(args.workload, device, use_engine=False)
    P, T = args.n_perms, wl["T"]
    n_jobs = max(1, min(os.cpu_count() or 1, 16))
    budget = max(5.0, min(args.cpu_budget, 60.0))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(wl, P, T, n_jobs=n_jobs, budget_s=budget)
        if i >= args.warmup:
            vals.append(last["value"])
        if i == 0 and args.warmup + args.steps > 2:
            budget = max(5.0, budget / 2)
    value = float(np.mean(vals))
    last["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": P * T / value * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 sums / f64 compare", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "cells": wl["N"], "genes_in_resource": wl["G"],
                       "clusters": wl["L"], "nnz": wl["nnz"], "n_perms": P, "triples_kept": T, "expr_prop": 0.1},
            "cpu_baseline": last,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly ONE JSON line: everything any library prints to file descriptor 1 (NCCL's
    # "NCCL version ..." banner, torchrun notices) is sent to stderr, and the line is written to the saved fd
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--order", default="fast", choices=["fast", "exact"])
    ap.add_argument("--n-perms", type=int, default=1000)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-exact", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-cellchat", action="store_true")
    ap.add_argument("--no-stress", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3        # timing rule: at least 3 warm-up steps
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")   # joblib workers import oracle/
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
