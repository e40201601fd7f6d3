#!/usr/bin/env python
"""Benchmark of the permutation-null hot path (metric of BASELINE.json:
permutations x LR-triples per second for cellphonedb, n_perms=1000).

  python bench.py [--gpus N --steps K --warmup W] [--workload c3|c2|c4|c5|tiny] [--order fast|exact]
  python bench.py --impl reference ...      # the reference's OWN hot path (unmodified modules) on the box's host cores

The synthetic input is the FULL matrix the config names (e.g. configs[2]: 500,000 cells x 30,000 genes, 8 % density,
50 clusters; the 1,877 genes of the consensus resource sit at shuffled columns among the background genes), held as a
host SciPy CSR inside an AnnData-like object - what a liana user has.

A "step" = one full pass of the hot path (label shuffles -> per-cluster means -> score -> exceedance counts for all
P permutations and all T kept triples).
  value   device time of a step with the inputs resident in HBM (the matrix prepared by the product's own input
          path: staged, subset to the resource's genes and committed on the device), CUDA events, L2 flushed between steps
  e2e     the same metric through the user-facing plugin call `li.mt.cellphonedb(adata, groupby, n_perms=P)` on the
          host AnnData -> result DataFrame: input checks, H2D of the full matrix, gene subset, statistics, triple
          resolution, permutations, p-values, sorted frame - every copy inside the timed region
  e2e_cabi_subset   round 1's e2e: the C ABI fed with pinned host buffers of the gene-subset matrix
Multi-GPU: one process per GPU (torchrun), permutations sharded by global id, one NCCL all-reduce of the uint32
counters per step (strong scaling: total work fixed).  `counts_sha256` (of the all-reduced counters) is identical
for every GPU count.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells, genes, clusters, description)
    "c2": (50_000, 20_000, 20, "configs[1]: 50k cells x 20k genes CSR 8%, 20 clusters, consensus, n_perms=1000"),
    "c3": (500_000, 30_000, 50, "configs[2]: 500k cells x 30k genes CSR 8%, 50 clusters, consensus, n_perms=1000"),
    "c4": (200_000, 30_000, 40, "configs[3] shape: 200k cells x 30k genes, 40 clusters, n_perms=1000"),
    "c5": (10_000, 30_000, 30, "configs[4] one sample: 10k cells x 30k genes, 30 clusters, n_perms=1000"),
    "tiny": (3_000, 2_500, 8, "smoke-size workload for CI: 3k cells x 2.5k genes, 8 clusters"),
}
METRIC = "permutations x LR-triples per second (cellphonedb, n_perms=1000)"
UNIT = "perm*triples/s"
L2_NOTE = "L2 flushed (256 MB write) before every timed step"


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


# ---------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------
def gene_layout(G_full, seed=1339):
    """Names of the G_full columns: the consensus resource's 1,877 subunit symbols + background genes, shuffled."""
    from liana_py_b200.testing._synth import _gene_names
    names = _gene_names(G_full)
    order = np.random.default_rng(seed).permutation(G_full)
    return names[order]


def gen_full(name, device, seed_data=1337, seed_labels=1338):
    """The config's FULL synthetic matrix as torch tensors on `device` (CUDA, or CPU for the smoke size)."""
    from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
    N, G_full, L, desc = WORKLOADS[name]
    indptr, indices, data, labels = synthetic_csr_torch(N, G_full, L, 0.08, seed_data, seed_labels, device=device,
                                                        chunk=16384 if G_full <= 30_000 else 8192)
    return dict(name=name, desc=desc, N=N, G_full=G_full, L=L, indptr=indptr, indices=indices, data=data, labels=labels,
                var_names=gene_layout(G_full), nnz_full=int(indices.numel()))


def host_adata(full):
    """The user's object: AnnData-like, X = host SciPy CSR float32 (pageable NumPy buffers), obs['cluster'] categorical."""
    import pandas as pd
    from scipy import sparse
    from liana_py_b200.testing._anndata_lite import AnnDataLite
    X = sparse.csr_matrix((full["data"].cpu().numpy(), full["indices"].cpu().numpy(), full["indptr"].cpu().numpy()),
                          shape=(full["N"], full["G_full"]))
    cats = [f"c{i:02d}" for i in range(full["L"])]
    obs = pd.DataFrame({"cluster": pd.Categorical.from_codes(full["labels"].cpu().numpy().astype(np.int64), categories=cats)})
    return AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(full["var_names"])))


def subset_torch(full, genes):
    """Gene subset of the full matrix with torch ops (reference arm: no engine code on its path): the columns named in
    `genes` (sorted), renumbered in that order.  Returns a host SciPy CSR with sorted indices."""
    import torch
    from scipy import sparse
    dev = full["indices"].device
    pos = {g: i for i, g in enumerate(genes)}
    col_map = torch.tensor([pos.get(n, -1) for n in full["var_names"]], dtype=torch.int32, device=dev)
    N = full["N"]
    counts = torch.diff(full["indptr"].long())
    new_counts = torch.zeros(N, dtype=torch.int64, device=dev)
    idx_parts, val_parts = [], []
    step = max(1, N // 16)
    for r0 in range(0, N, step):               # row blocks: the row-id expansion of 1.2 G entries is done piecewise
        r1 = min(N, r0 + step)
        e0, e1 = int(full["indptr"][r0]), int(full["indptr"][r1])
        m = col_map[full["indices"][e0:e1].long()]
        keep = m >= 0
        rows = torch.repeat_interleave(torch.arange(r0, r1, device=dev), counts[r0:r1])
        new_counts.index_add_(0, rows[keep], torch.ones(int(keep.sum()), dtype=torch.int64, device=dev))
        idx_parts.append(m[keep])
        val_parts.append(full["data"][e0:e1][keep])
    indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(new_counts, 0)
    X = sparse.csr_matrix((torch.cat(val_parts).cpu().numpy(), torch.cat(idx_parts).cpu().numpy(),
                           indptr.to(torch.int32).cpu().numpy()), shape=(N, len(genes)))
    X.sort_indices()
    return X


def workload_config(wl, P, world):
    """The `config` object of the JSON line - identical for the engine arm and the reference arm."""
    return {"workload": f"{wl['name']}: {wl['desc']}", "cells": wl["N"], "genes": wl["G_full"], "nnz_full": wl["nnz_full"],
            "genes_in_resource": wl["G"], "clusters": wl["L"], "nnz": wl["nnz"], "n_perms": P,
            "triples_kept": wl["T"], "triples_max": wl["T_max"], "expr_prop": 0.1,
            "parallelism": f"perm-shard x{world}", "l2": L2_NOTE}


def sha_counts(arr) -> str:
    return hashlib.sha256(np.ascontiguousarray(np.asarray(arr), dtype="<u4").tobytes()).hexdigest()


def model_bytes(wl, P):
    """one streaming pass over X_sub per permutation."""
    per_perm = 8 * wl["nnz"] + 4 * (wl["N"] + 1) + 4 * wl["N"] + 8 * wl["L"] * wl["G"]
    return per_perm, P * per_perm + 24 * wl["T"]


def load_static_profile(key):
    """Numbers that come from a committed ncu capture, not from this run (labelled as such in the line)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    ent = d.get(key)
    if ent is None:
        return None
    if not isinstance(ent, dict):
        ent = {"dram_bytes_per_launch": ent}
    ent = dict(ent)
    ent["source"] = "static: profiles/traffic.json (ncu --set full capture, not measured by this run)"
    return ent


# ---------------------------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import liana_py_b200 as li
    from liana_py_b200 import _engine as E
    from liana_py_b200.method import _pipeline as PL
    from liana_py_b200.method._dist import DistContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun (--nproc-per-node N) for --gpus N > 1")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_bound = li.mt.bind_to_gpu_numa(local) if world > 1 else False      # before the host matrix is allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    dctx = DistContext.from_torch() if world > 1 else None
    P = args.n_perms
    order = E.ORDER_FAST if args.order == "fast" else E.ORDER_EXACT

    # ---- the config's full matrix, on the host like a user's AnnData; prepared by the product's own input path ----
    t0 = time.perf_counter()
    full = gen_full(args.workload, device)
    adata = host_adata(full)
    t_gen = time.perf_counter() - t0
    for k in ("indptr", "indices", "data"):
        full[k] = None                                     # the device copy of the generator is not an input of anything
    torch.cuda.empty_cache()
    eng = PL.get_engine(local)
    eng.set_stream(torch.cuda.current_stream(device).cuda_stream)
    state = PL.prepare(adata, "cluster", "consensus", None, None, None, 5, False, None, False, device=local)
    tri, _ = state.resolved(0.1, False)
    obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    T = int(len(obs))
    N, G, nnz = eng.matrix_info()
    wl = dict(name=args.workload, desc=full["desc"], N=N, G=G, G_full=full["G_full"], nnz_full=full["nnz_full"], L=full["L"],
              nnz=nnz, T=T, T_max=int(state.tables.lig_sub.shape[0]) * full["L"] ** 2, tri=tri, obs=obs)
    lo, hi = (P * rank) // world, (P * (rank + 1)) // world
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
    run_stats = eng.run_stats() if args.order == "fast" else {"genes_scored": G, "nnz_scored": nnz, "chunks_run": 0, "nnz": nnz}
    ms_per_step = total_ms / args.steps
    value = P * T / (ms_per_step * 1e-3)
    result_counts = counts.clone()
    counts_sha = sha_counts(result_counts.cpu().numpy())

    # ---- round 1's e2e: the C ABI with pinned host buffers of the gene-subset matrix ----
    sub_ip, sub_ix, sub_dt = eng.export_matrix_csr()
    labels_sub = np.ascontiguousarray(state.prep.labels, dtype=np.int32)
    host = {"indptr": torch.from_numpy(sub_ip).pin_memory(), "indices": torch.from_numpy(sub_ix).pin_memory(),
            "data": torch.from_numpy(sub_dt).pin_memory(), "labels": torch.from_numpy(labels_sub).pin_memory()}
    h_idx = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
             for a in (tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)]
    h_obs = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32)).pin_memory()
    h_counts = torch.zeros(max(T, 1), dtype=torch.int32).pin_memory()
    h2d_sub = sum(int(t.numel() * t.element_size()) for t in list(host.values()) + h_idx + [h_obs])
    d2h_sub = int(h_counts.numel() * 4)
    if world > 1:       # whole job: every rank uploads the small tables, the matrix arrays cross PCIe once in total
        mat = sum(int(host[k].numel() * host[k].element_size()) for k in ("indices", "data"))
        h2d_sub = world * (h2d_sub - mat) + mat
        d2h_sub *= world

    def step_cabi():
        if world > 1:
            d_ip, d_ix, d_dt = dctx.upload_csr_sharded(host["indptr"], host["indices"], host["data"])
            eng.set_matrix_csr(d_ip, d_ix, d_dt, N, G)
        else:
            eng.set_matrix_csr(host["indptr"], host["indices"], host["data"], N, G)
        eng.set_labels(host["labels"], wl["L"])
        eng.set_triples(h_idx[0], h_idx[1], h_idx[2], h_idx[3], h_obs, None)
        eng.run(hi - lo, seed=1337, perm_offset=lo, order=order, scores=E.SCORE_CPDB, counts_out=(counts, None))
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        h_counts.copy_(counts, non_blocking=False)

    def wall(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        s = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.MAX)
        return float(s.item())

    e2e_steps = max(1, min(args.steps, 5))
    cabi_s = wall(step_cabi, e2e_steps, max(3, args.warmup))
    assert torch.equal(h_counts.to(device), result_counts), "C-ABI e2e counts differ from the resident run"
    del host

    # ---- e2e: the user-facing plugin call on the host AnnData (full matrix) ----
    keep = {}

    def step_plugin():
        keep["res"] = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, inplace=False, device=local,
                                        dist=dctx, order=args.order)

    plugin_s = wall(step_plugin, max(1, min(args.steps, 3)), 2)
    res = keep["res"]
    assert len(res) == T, (len(res), T)
    got = np.rint(res["cellphone_pvals"].values * P).astype(np.int64)
    assert np.array_equal(got, result_counts.cpu().numpy().astype(np.int64)[res.index.values]), "plugin p-values differ from the resident run"
    X = adata.X
    n_rows_plugin = int(len(res))
    # whole job: the matrix crosses PCIe once (every rank stages its own block of rows), the small tables once per rank
    small_h2d = int(4 * N + 20 * T + 12 * wl["L"] * G)
    h2d_plugin = int(X.indptr.nbytes + X.indices.nbytes + X.data.nbytes) + world * small_h2d
    d2h_plugin = world * int(4 * T + 45 * T + 4 * T + 8 * wl["L"] * G + 8 * full["G_full"])

    # ---- configs[3] / configs[4]: the other two calls BASELINE.json names, through the plugin API, every rank taking part ----
    extra = {}
    if not args.no_configs34:
        n_rows_plugin = int(len(res))
        adata = res = None
        keep.clear()
        extra["configs3_rank_aggregate"] = _guarded("configs3_rank_aggregate", lambda: configs3_rank_aggregate(device, local, dctx, world, P))
        extra["configs4_by_sample"] = _guarded("configs4_by_sample", lambda: configs4_by_sample(device, local, dctx, world, P))
        if world == 1:      # the engine holds another matrix now: put the main workload back for the sections below
            adata = host_adata(gen_full(args.workload, device))
            state = PL.prepare(adata, "cluster", "consensus", None, None, None, 5, False, None, False, device=local)
            tri, _ = state.resolved(0.1, False)
            obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
            wl["tri"], wl["obs"] = tri, obs
            eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_gbs, peak_src = peaks()
    per_perm_bytes, total_bytes = model_bytes(wl, P)
    kern_avg_ms = kern_ms / max(kern_n, 1)
    perms_per_launch = (hi - lo) / max(tm["main_kernel_launches"], 1)
    achieved_gbs = per_perm_bytes * perms_per_launch / (kern_avg_ms * 1e-3) / 1e9
    static = load_static_profile(f"{args.workload}_{args.order}")
    traffic = static.get("dram_bytes_per_launch") if static else None
    kernel_name = "fast2_means_kernel" if args.order == "fast" else "exact_means_kernel"
    hbm_model = {"achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                 "peak_source": peak_src, "algorithmic_bytes_per_perm": per_perm_bytes,
                 "note": "streaming model of SURVEY.md 8(d): one pass over the gene-subset matrix per permutation "
                         "(8 B per stored entry of ALL its genes)"}
    if args.order == "fast":
        # The fast kernel does not stream the matrix once per permutation: it re-uses an L2-resident label tile for 512
        # permutations and (round 2) accumulates only the genes some kept triple reads, so against the streaming model
        # it "exceeds" the DRAM peak.  What bounds it is the shared-memory pipe: every (stored entry, permutation)
        # update is one LDS + one STS of a 32-lane conflict-free wavefront (2 wavefronts per update-warp of 32
        # updates: the hard floor, 16 updates/clk/SM), plus 1/4 wavefront of label load and 1/4 of staged entry reads
        # per update-warp (Q = 4) = 2.5 wavefronts -> 12.8 updates/clk/SM.
        sm_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        upd = perms_per_launch * run_stats["nnz_scored"]
        upd_clk_sm = upd / (kern_avg_ms * 1e-3) / sm_clk / 148
        hbm_model["note"] += ("; this kernel re-uses L2-resident label tiles across permutations and skips unreferenced genes, "
                              "so the model fraction exceeds 1 - it is kept for comparability, the binding resource is the shared-memory pipe")
        roofline = {"bound": "smem_pipe", "kernel": kernel_name, "achieved": upd_clk_sm, "peak": 12.8,
                    "unit": "bin updates/clk/SM", "frac": upd_clk_sm / 12.8,
                    "peak_hard_floor": 16.0, "frac_hard_floor": upd_clk_sm / 16.0,
                    "peak_source": "1 shared-memory wavefront per clock per SM; 2.5 wavefronts per 32 updates (LDS + STS + 0.25 label + 0.25 "
                                   "staged entries) = 12.8, 2 wavefronts (LDS + STS alone) = 16",
                    "updates_per_launch": upd, "sm_clock_mhz": sm_clk / 1e6, "sm_count": 148,
                    "kernel_ms_per_launch": kern_avg_ms, "perms_per_launch": perms_per_launch,
                    "traffic": traffic, "traffic_detail": static,
                    "physical_dram_frac": (traffic / (kern_avg_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
                    "hbm_streaming_model": hbm_model}
    else:
        roofline = {"bound": "hbm", "kernel": kernel_name, **hbm_model, "traffic": traffic, "traffic_detail": static,
                    "perms_per_launch": perms_per_launch, "kernel_ms_per_launch": kern_avg_ms}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 sums / f64 compare", "data": "synthetic",
        "config": workload_config(wl, P, world),
        "engine_config": {"order": args.order, "perm_source": "philox", "perm_batch": tm["perm_batch"],
                          "genes_scored": run_stats["genes_scored"], "nnz_scored": run_stats["nnz_scored"],
                          "note": "a run that exports no cube accumulates only the genes some kept triple reads; "
                                  "counts are bit-identical to accumulating all of them (tests)"},
        "counts_sha256": counts_sha,
        "e2e": {"value": P * T / plugin_s, "unit": UNIT, "h2d_bytes_per_step": h2d_plugin, "d2h_bytes_per_step": d2h_plugin,
                "ms_per_step": plugin_s * 1e3, "rows": n_rows_plugin,
                "call": "li.mt.cellphonedb(adata, groupby, use_raw=False, n_perms=1000, inplace=False) on a host AnnData-like "
                        "object holding the FULL matrix (SciPy CSR, pageable memory) -> sorted liana_res DataFrame"},
        "e2e_cabi_subset": {"value": P * T / cabi_s, "unit": UNIT, "h2d_bytes_per_step": h2d_sub, "d2h_bytes_per_step": d2h_sub,
                            "ms_per_step": cabi_s * 1e3,
                            "call": "C ABI (set_matrix_csr / set_labels / set_triples / run) with pinned host buffers of the gene-subset matrix"},
        "gpu_launches": launches,
        "clocks": clocks,
        "stages_ms": {k: tm[k] for k in ("setup_ms", "means_ms", "score_ms", "total_ms")},
        "roofline": roofline,
        "setup_s": {"generate_and_copy_to_host": t_gen, "process_bound_to_gpu_numa_node": bool(numa_bound)},
        **extra,
    }
    restore = lambda: eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)    # noqa: E731
    if world == 1 and not args.no_stress:
        line["stress_expr_prop0"] = _guarded("stress_expr_prop0", lambda: stress_variant(state, P, order, counts, timed))
        restore()
    if world == 1 and not args.no_cellchat:
        line["cellchat_null"] = _guarded("cellchat_null", lambda: cellchat_null(state, adata, wl, P, timed, cpu=not args.no_cpu))
        restore()
    if world == 1 and args.order == "fast" and not args.no_exact:
        def exact_section():
            ex_ms, _, ex_kern, ex_n, _ = timed(lambda: eng.run(P, seed=1337, order=E.ORDER_EXACT, scores=E.SCORE_CPDB,
                                                             counts_out=(counts, None)), 2, 1)
            ex_ms /= 2
            ex_static = load_static_profile(f"{args.workload}_exact")
            return {"ms_per_step": ex_ms, "value": P * T / (ex_ms * 1e-3), "kernel": "exact_means_kernel",
                    "roofline": {"bound": "hbm", "frac": per_perm_bytes * P / (ex_kern / 2 * 1e-3) / 1e9 / peak_gbs,
                                 "peak": peak_gbs, "traffic_detail": ex_static},
                    "note": "bit-identical-to-reference summation order (validation mode); every gene is computed"}
        line["exact_order_mode"] = _guarded("exact_order_mode", exact_section)
    if world == 1 and not args.no_dropin:
        line["dropin_call_subset"] = _guarded("dropin_call_subset", lambda: dropin_subset(eng, state, wl, P, local))
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = _guarded("cpu_baseline", lambda: cpu_baseline(subset_inputs(eng, state), wl, P, n_jobs=1,
                                                                            budget_s=args.cpu_budget))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _guarded(section, fn):
    """A secondary section must not cost the line its primary numbers: an exception is printed to stderr and recorded
    as {"error": ...} under the section's key (all sections run on every rank that takes part, so ranks fail alike)."""
    try:
        return fn()
    except Exception as e:      # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"section": section, "error": f"{type(e).__name__}: {e}"[:300]}


def _wall_max(fn, world, device, repeats=1):
    """Best wall time of `fn` over `repeats` runs (all ranks take part in every run), max over ranks."""
    import torch
    import torch.distributed as dist
    best = None
    for _ in range(repeats):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t = torch.tensor([best], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def configs3_rank_aggregate(device, local, dctx, world, P):
    """BASELINE.json configs[3]: `li.mt.rank_aggregate` (cellphonedb permutation null + connectome / logfc / natmi / sca
    scores + RRA consensus) on 200k cells x 30k genes x 40 clusters, n_perms=1000 - wall time of the plugin call on the
    host AnnData (full matrix), permutations (and, for N > 1, the staging of the matrix) sharded over the ranks."""
    import liana_py_b200 as li
    full = gen_full("c4", device)
    adata = host_adata(full)
    for k in ("indptr", "indices", "data"):
        full[k] = None

    def call():
        return li.mt.rank_aggregate(adata, groupby="cluster", use_raw=False, n_perms=P, inplace=False, device=local, dist=dctx)

    call()                                                   # warm-up (caches of the resource tables, CUDA buffers)
    wall, res = _wall_max(call, world, device, repeats=2)
    return {"call": "li.mt.rank_aggregate(adata, groupby, use_raw=False, n_perms=1000, inplace=False)", "wall_ms": wall * 1e3,
            "rows": int(len(res)), "cells": full["N"], "genes": full["G_full"], "clusters": full["L"], "nnz_full": full["nnz_full"],
            "n_gpus": world, "value": P * len(res) / wall, "unit": UNIT,
            "frame_sha256": hashlib.sha256(np.ascontiguousarray(res["magnitude_rank"].values).tobytes()
                                           + np.ascontiguousarray(res["cellphone_pvals"].values).tobytes()).hexdigest(),
            "note": "best of 2 after a warm-up; host AnnData-like input (full matrix) -> consensus frame; max over ranks"}


def by_sample_adata(device, n_samples=96, cells=10_000, genes=5_000, clusters=30):
    """configs[4] input: `n_samples` samples x 10k cells x 30 clusters (sample-specific seeds 1337 + s) in
    ONE AnnData-like object; 5,000 genes (the 1,877 resource genes + background) - BASELINE.json does not name a gene
    count for this config, and 96 x 10k x 30k genes would be 30 GB of host CSR per rank."""
    import pandas as pd
    import torch
    from scipy import sparse
    from liana_py_b200.testing._anndata_lite import AnnDataLite
    from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
    ips, ixs, dts, labs, base = [], [], [], [], 0
    for smp in range(n_samples):
        ip, ix, dt, lab = synthetic_csr_torch(cells, genes, clusters, 0.08, 1337 + smp, 2338 + smp, device=device)
        ips.append(ip[1:].to(torch.int64) + base)
        base += int(ip[-1])
        ixs.append(ix); dts.append(dt); labs.append(lab)
    indptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=device)] + ips)
    idt = np.int32 if base < 2 ** 31 else np.int64
    X = sparse.csr_matrix((torch.cat(dts).cpu().numpy(), torch.cat(ixs).cpu().numpy().astype(idt, copy=False),
                           indptr.cpu().numpy().astype(idt)), shape=(n_samples * cells, genes))
    obs = pd.DataFrame({
        "cluster": pd.Categorical.from_codes(torch.cat(labs).cpu().numpy().astype(np.int64), categories=[f"c{i:02d}" for i in range(clusters)]),
        "sample": pd.Categorical.from_codes(np.repeat(np.arange(n_samples), cells), categories=[f"S{i:02d}" for i in range(n_samples)])})
    return AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(gene_layout(genes)))), base


def configs4_by_sample(device, local, dctx, world, P):
    """BASELINE.json configs[4]: `li.mt.cellphonedb.by_sample` over 96 synthetic samples x 10k cells x 30 clusters,
    samples sharded over the ranks; the ONE frame the API promises (all samples, names materialised) on every rank."""
    import liana_py_b200 as li
    adata, nnz = by_sample_adata(device)
    n_samples = len(adata.obs["sample"].cat.categories)

    def call():
        return li.mt.cellphonedb.by_sample(adata, "sample", groupby="cluster", use_raw=False, n_perms=P, inplace=False,
                                           device=local, dist=dctx)

    def call_rank0():
        return li.mt.cellphonedb.by_sample(adata, "sample", groupby="cluster", use_raw=False, n_perms=P, inplace=False,
                                           device=local, dist=dctx, gather=0)

    warm = adata[np.asarray(adata.obs["sample"].cat.codes) < min(n_samples, 2 * world)].copy()
    for g in ((True, 0, False) if world > 1 else (True,)):          # every collective the timed calls use has run once
        li.mt.cellphonedb.by_sample(warm, "sample", groupby="cluster", use_raw=False, n_perms=P, inplace=False, device=local,
                                    dist=dctx, gather=g)
    out = {"call": "li.mt.cellphonedb.by_sample(adata, 'sample', groupby='cluster', use_raw=False, n_perms=1000, inplace=False)",
           "samples": n_samples, "cells_per_sample": 10_000, "genes": int(adata.shape[1]), "clusters": 30, "nnz_full": int(nnz),
           "n_gpus": world, "unit": UNIT}
    if world > 1:
        # the frame of all samples on ONE rank (what a single-process caller of the reference gets) ...
        wall0, res0 = _wall_max(call_rank0, world, device, repeats=1)
        out["wall_ms"] = wall0 * 1e3
        # ... and on EVERY rank (SPMD default): `world` copies of the 28 M-row frame are assembled side by side on the host
        wall, res = _wall_max(call, world, device, repeats=1)
        out["wall_ms_frame_on_every_rank"] = wall * 1e3
        # ... and with every rank keeping the frame of ITS samples (gather=False: no exchange, to be written out per rank)
        wall_own, _ = _wall_max(lambda: li.mt.cellphonedb.by_sample(adata, "sample", groupby="cluster", use_raw=False, n_perms=P,
                                                                    inplace=False, device=local, dist=dctx, gather=False),
                                world, device, repeats=1)
        out["wall_ms_frames_stay_on_their_ranks"] = wall_own * 1e3
    else:
        wall, res = _wall_max(call, world, device, repeats=1)
        out["wall_ms"] = wall * 1e3
    rows = int(len(res))
    out.update({"ms_per_sample": out["wall_ms"] / n_samples, "rows": rows, "value": P * rows / (out["wall_ms"] * 1e-3),
                "frame_sha256": hashlib.sha256(np.ascontiguousarray(res["cellphone_pvals"].values).tobytes()
                                               + np.ascontiguousarray(res["lr_means"].values).tobytes()).hexdigest(),
                "note": "one timed run after a warm-up on a few samples; samples sharded over the ranks, numeric columns gathered over "
                        "NCCL, the name columns built once on the consumer; `wall_ms`: the frame of all samples on rank 0 (gather=0; "
                        "N = 1: the plain call), `wall_ms_frame_on_every_rank`: the default (gather=True); max over ranks"})
    return out


def stress_variant(state, P, order, counts, timed):
    """the same workload with expr_prop=0, i.e. every (source, target, interaction) row reaches the
    score function (T = T_max) and every gene of the resource is read: the score + count stage grows ~15x."""
    import torch
    from liana_py_b200 import _engine as E
    from liana_py_b200.method import _pipeline as PL
    eng = state.engine
    tri, _ = state.resolved(0.0, False)
    obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    T = len(obs)
    eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
    big = torch.zeros(max(T, 1), dtype=torch.int32, device=counts.device)
    total_ms, _, _, _, tm = timed(lambda: eng.run(P, seed=1337, order=order, scores=E.SCORE_CPDB, counts_out=(big, None)), 2, 3)
    ms = total_ms / 2
    return {"triples": T, "ms_per_step": ms, "value": P * T / (ms * 1e-3), "unit": UNIT, "run_stats": eng.run_stats(),
            "stages_ms": {k: tm[k] for k in ("setup_ms", "means_ms", "score_ms", "total_ms")},
            "note": "expr_prop=0: all T_max rows scored; resident inputs, L2 flushed between steps"}


def cellchat_null(state, adata, wl, P, timed, cpu=True):
    """Secondary number: the CellChat null on the same matrix - per-(permutation, cluster, gene)
    TRIMEANS (three quantiles of the dense column, selected exactly from the stored entries) instead of means,
    prod / (kh + prod) score, same metric.  Device time with resident inputs, L2 flushed between steps."""
    from liana_py_b200.method import _pipeline as PL, _tables
    eng = state.engine
    # scipy's sparse max over the prepared matrix (non-empty genes of the kept cells): the implicit zeros take part
    mat_max = np.float32(max(adata.X.data.max(), 0.0))
    obs_tm = eng.observed_trimean(mat_max)
    tri = _tables.resolve_triples(state.tables, obs_tm, state.props, 0.1, engine=eng)
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
           "counts_sha256": sha_counts(keep["out"]["cellchat"]),
           "mean_pvalue": float(keep["out"]["cellchat"].mean() / P),
           "note": "lrperm_run_trimean: count (fast2_means_kernel<COUNT>, shared-memory-pipe bound like the cellphonedb "
                   "kernel; referenced genes only) + prefix + select + finalize + score; bit-identical to the reference for identical permutations"}
    if cpu:
        from oracle import numpy_port
        Xs, labels = subset_inputs(eng, state)[:2]
        t0 = time.perf_counter()
        numpy_port.get_trimean_perms(Xs, labels, wl["L"], 1, 1337, mat_max)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": T / dt, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"1 of {P} permutations in {dt:.1f} s (NumPy restatement of _get_means_perms with "
                                         f"agg_fun=_trimean: dense np.quantile / np.median per cluster); extrapolated full run: {dt * P:.0f} s"}
    return out


_SUBSET_CACHE = {}


def subset_inputs(eng, state):
    """(gene-subset SciPy CSR, labels, gene names) of the prepared matrix, for the CPU baselines."""
    if "v" not in _SUBSET_CACHE:
        from scipy import sparse
        ip, ix, dt = eng.export_matrix_csr()
        n, g, _ = eng.matrix_info()
        X = sparse.csr_matrix((dt, ix, ip), shape=(n, g))
        X.sort_indices()
        _SUBSET_CACHE["v"] = (X, np.asarray(state.prep.labels), np.asarray(state.prep.var_names))
    return _SUBSET_CACHE["v"]


def dropin_subset(eng, state, wl, P, device):
    """Wall time of `li.mt.cellphonedb(adata, groupby, n_perms=P)` on an AnnData holding ONLY the resource's genes (the
    north_star's boundary: "already subset to the resource's ligand/receptor genes"); round 1's `dropin_call`."""
    import pandas as pd
    import liana_py_b200 as li
    from liana_py_b200.testing._anndata_lite import AnnDataLite
    X, labels, names = subset_inputs(eng, state)
    obs = pd.DataFrame({"cluster": pd.Categorical([f"c{v:02d}" for v in labels])})
    adata = AnnDataLite(X=X, obs=obs, var=pd.DataFrame(index=pd.Index(names)))
    best, rows = None, 0
    for _ in range(3):
        t0 = time.perf_counter()
        res = li.mt.cellphonedb(adata, groupby="cluster", use_raw=False, n_perms=P, inplace=False, device=device)
        dt = time.perf_counter() - t0
        best, rows = (dt if best is None else min(best, dt)), len(res)
    return {"wall_ms": best * 1e3, "rows": rows, "value": P * rows / best, "unit": UNIT,
            "note": "best of 3; host AnnData-like input of the 1,877 resource genes -> liana_res DataFrame, Philox permutations, FAST order"}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own code on the host cores
# ---------------------------------------------------------------------------------------------------------------
_REF_CACHE = {}


def cpu_baseline(inputs, wl, P, n_jobs=1, budget_s=20.0):
    """The reference's hot path (`_liana_pipe.py:402-425`: _get_means_perms -> _get_mat_idx -> gather -> _cpdb_score) on
    a bounded sample of the same workload.  Runs the UNMODIFIED reference modules (oracle/_ref, kind "reference") when
    they are present, else the NumPy/SciPy restatement (kind "port").  A chunk of permutations and a slice of the rows
    are timed and scaled linearly (every stage is linear in P except _get_mat_idx, which is linear in T)."""
    X, labels, names = inputs
    tri, T = wl["tri"], wl["T"]
    from oracle import ref_hotpath
    if ref_hotpath.available():
        cache = _REF_CACHE.setdefault((id(X), n_jobs), {})
        if "inputs" not in cache:           # once per process: the stand-in objects, the sample size, `_get_mat_idx` (P-independent)
            adata, lr_res = ref_hotpath.make_inputs(X, labels, wl["L"], names, tri)
            idx_rows = int(min(T, 20_000))
            _, _, t1 = ref_hotpath.run_block(adata, lr_res, max(1, n_jobs), 1337, n_jobs=n_jobs, idx_rows=idx_rows)     # sizes the sample (and starts the workers)
            cache.update(inputs=(adata, lr_res), idx=(t1["get_mat_idx"], t1["get_mat_idx_rows"]),
                         per_perm=(t1["get_means_perms"] + t1["gather"] + t1["score"]) / max(1, n_jobs))
        adata, lr_res = cache["inputs"]
        p_s = int(max(n_jobs, min(P, budget_s / max(cache["per_perm"], 1e-4))))
        p_s = max(1, (p_s // max(n_jobs, 1)) * max(n_jobs, 1))
        # perm_stats + scores of the chunk (3 x P_chunk x T x 8 B) must stay below 25 % of RAM
        try:
            import psutil
            ram = psutil.virtual_memory().total
            p_s = int(max(1, min(p_s, 0.25 * ram / (24.0 * max(T, 1)))))
        except Exception:
            pass
        t0 = time.perf_counter()
        _, _, t = ref_hotpath.run_block(adata, lr_res, p_s, 1337, n_jobs=n_jobs, idx_reuse=cache["idx"])
        dt = time.perf_counter() - t0
        cache["per_perm"] = (t["get_means_perms"] + t["gather"] + t["score"]) / max(p_s, 1)
        full_s = ref_hotpath.full_run_seconds(t, p_s, P, T)
        return {"value": P * T / full_s, "unit": UNIT, "cores": n_jobs, "kind": "reference", "extrapolated": True,
                "measured_perms": p_s, "measured_seconds": dt, "stage_seconds": t, "full_run_seconds": full_s,
                "sample": f"{p_s} of {P} permutations of the same workload through the reference's own _get_means_perms (n_jobs={n_jobs}) "
                          f"+ gather + _cpdb_score in {dt:.1f} s, and _get_mat_idx on {t['get_mat_idx_rows']} of {T} rows in {t['get_mat_idx']:.1f} s "
                          f"(measured once per process: it does not depend on the permutations); "
                          f"scaled linearly in P (and in T for _get_mat_idx): full run {full_s:.0f} s",
                "host_cpu_count": os.cpu_count()}
    from oracle import numpy_port
    t0 = time.perf_counter()
    numpy_port.get_means_perms(X, labels, wl["L"], 1, 1337, 1)
    one = time.perf_counter() - t0
    p_s = int(max(n_jobs, min(P, budget_s / max(one, 1e-3) * max(n_jobs, 1) * (0.6 if n_jobs > 1 else 1.0))))
    p_s = max(1, (p_s // n_jobs) * n_jobs)
    t0 = time.perf_counter()
    numpy_port.run_method_block(X, labels, wl["L"], p_s, 1337, tri.lig_gene, tri.rec_gene, tri.src, tri.tgt,
                                tri.ligand_means.copy(), tri.receptor_means.copy(), "cellphonedb", n_jobs=n_jobs)
    dt = time.perf_counter() - t0
    return {"value": p_s * T / dt, "unit": UNIT, "cores": n_jobs, "kind": "port", "extrapolated": True,
            "measured_perms": p_s, "measured_seconds": dt, "full_run_seconds": dt / p_s * P,
            "sample": f"{p_s} of {P} permutations of the same workload in {dt:.1f} s "
                      f"(NumPy/SciPy restatement of _get_means_perms + gather + _cpdb_score, n_jobs={n_jobs}; oracle/_ref absent); "
                      f"extrapolated full run: {dt / p_s * P:.0f} s",
            "host_cpu_count": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path with all the host threads it can use.
    No engine code anywhere on this arm: the matrix comes from the same torch generator, its gene subset from torch ops,
    the observed statistics from the C oracle, the kept rows from the NumPy table code."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from liana_py_b200.method import _pipeline as PL, _tables
    from liana_py_b200.resource import select_resource
    from liana_py_b200.testing._synth import consensus_genes
    from oracle import c_oracle
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    full = gen_full(args.workload, device)
    genes = consensus_genes()
    X = subset_torch(full, genes)
    nonempty = np.asarray(X.sum(axis=0)).ravel() != 0          # empty genes are removed by prep_check_adata (`_pre.py:122-127`)
    if not nonempty.all():
        X, genes = X[:, np.flatnonzero(nonempty)].tocsr(), genes[nonempty]
    labels = full["labels"].cpu().numpy()
    for k in ("indptr", "indices", "data"):
        full[k] = None
    means, stored, sizes = c_oracle.observed(X, labels, full["L"])
    props = stored / sizes[:, None].astype(np.float64)
    tables = _tables.build_resource_tables(select_resource("consensus"), genes)
    if not np.array_equal(tables.entities, genes):            # genes only named by interactions that lost a subunit: not in the subset
        sel = np.flatnonzero(np.isin(genes, tables.entities))
        X, genes, means, props = X[:, sel].tocsr(), genes[sel], means[:, sel], props[:, sel]
    tri = _tables.resolve_triples(tables, means, props, 0.1)
    obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
    wl = dict(name=args.workload, desc=full["desc"], N=full["N"], G=len(genes), G_full=full["G_full"], nnz_full=full["nnz_full"],
              L=full["L"], nnz=int(X.nnz), T=int(len(obs)), T_max=int(tables.lig_sub.shape[0]) * full["L"] ** 2, tri=tri, obs=obs)
    P, T = args.n_perms, wl["T"]
    n_jobs = max(1, min(os.cpu_count() or 1, 32))
    # every step is a bounded sample of the workload; the whole `--steps K --warmup W` run is sized to ~2 minutes of sampling
    n_iter = max(1, args.warmup + args.steps)
    budget = max(2.0, min(args.cpu_budget, 100.0 / n_iter))
    vals, last = [], None
    for i in range(n_iter):
        last = cpu_baseline((X, labels, genes), wl, P, n_jobs=n_jobs, budget_s=budget)
        if i >= args.warmup:
            vals.append(last["value"])
    value = float(np.mean(vals))
    last["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": P * T / value * 1e3, "extrapolated": True,
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 sums / f64 compare", "data": "synthetic",
            "config": workload_config(wl, P, args.gpus),
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
    ap.add_argument("--no-configs34", action="store_true", help="skip the configs[3] / configs[4] sections")
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
