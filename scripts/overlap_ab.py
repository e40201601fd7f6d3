"""configs[2] shape with the consensus triples (genes no triple reads are skipped): the two-stream pipeline (label slab of
batch b+1 and finalize / score of batch b-1 concurrent with the main kernel) against the serial single-stream schedule.
   python scripts/overlap_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200 import _engine as E
from liana_py_b200.method import _pipeline as PL, _tables
from liana_py_b200.resource import select_resource
from liana_py_b200.testing._synth import consensus_genes
from liana_py_b200.testing._synth_gpu import synthetic_csr_torch
N, L = {"c3": (500_000, 50), "c2": (50_000, 20), "c4": (200_000, 40)}[sys.argv[1] if len(sys.argv) > 1 else "c3"]
genes = consensus_genes(); G = len(genes)
indptr, indices, data, labels = synthetic_csr_torch(N, G, L, 0.08)
eng = E.Engine(0); eng.set_matrix_csr(indptr, indices, data, N, G); eng.set_labels(labels, L)
means, stored, sizes = eng.observed()
tri = _tables.resolve_triples(_tables.build_resource_tables(select_resource("consensus"), genes), means, stored / sizes[:, None].astype(np.float64), 0.1)
obs = PL.observed_cpdb(tri.ligand_means, tri.receptor_means)
eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs, None)
counts = torch.zeros(len(obs), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for P in (500, 250):
    for overlap, pb, sc, mode in ((1, 0, 0, 2), (1, 256, 0, 2), (1, 512, 0, 2), (1, 0, 0, 2), (1, 256, 0, 2), (1, 128, 0, 2), (1, 512, 0, 2)):
        eng.set_option("overlap", overlap); eng.set_tuning(pb, 0, 0); eng.set_option("slab_ctas_per_sm", sc); eng.set_option("slab_mode", mode)
        ts = []
        for it in range(12):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(P, seed=1337, order=E.ORDER_FAST, scores=E.SCORE_CPDB, counts_out=(counts, None)); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        c = counts.cpu().numpy()
        if ref is None: ref = c
        tm = eng.timings(); st = eng.run_stats()
        upd = st["nnz_scored"] * P / (tm["main_kernel_ms"] * 1e-3) / 148 / 1.965e9
        print(f"P={P} overlap={overlap} batch={pb or 512} slab_ctas_per_sm={sc} slab_mode={mode}: {np.median(ts[2:]):.3f} ms (setup {tm['setup_ms']:.2f} main {tm['main_kernel_ms']:.2f} means {tm['means_ms']:.2f} "
              f"score {tm['score_ms']:.2f}; main kernel {upd:.2f} updates/clk/SM) same={np.array_equal(c, ref) if P == 1000 else '-'}", flush=True)
