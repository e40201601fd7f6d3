"""TEST INFRASTRUCTURE - never imported by the product.

A stand-in for `liana_py_b200._engine.Engine` whose every device call is answered by the CPU oracle (`oracle/`) or by
NumPy / SciPy.  With it installed, the drop-in layer (`liana_py_b200/method/`: input contract, resource tables, row
resolution, score functions, consensus, frame assembly, by_sample, MuData, supplementary columns) runs end to end on a
machine without a GPU, and `tests/test_dropin_cpu.py` compares its frames with the frames the reference produced
(tests/golden).  The GPU suite runs the very same comparisons through the CUDA library."""
import numpy as np
from scipy.stats import rankdata

from liana_py_b200 import _engine as E
from liana_py_b200.method import _aggregate, _tables
from oracle import c_oracle, numpy_port
from test_host_logic import SciPyStage


class OracleEngine(SciPyStage):
    _h = object()              # "open" for _pipeline.get_engine

    def __init__(self):
        self.N = self.G = self.L = self.T = 0

    # -- matrix / labels / triples ----------------------------------------------------------
    def commit_matrix(self, col_map, n_genes_out, n_cells_out):
        col_map = np.asarray(col_map)
        src = np.flatnonzero(col_map >= 0)
        inv = np.empty(int(n_genes_out), dtype=np.int64)
        inv[col_map[src]] = src
        rows = np.arange(self.X.shape[0]) if self.keep is None else np.flatnonzero(self.keep)
        Xc = self.X[rows][:, inv].tocsr().astype(np.float32)
        Xc.sort_indices()
        self.Xc, self.N, self.G = Xc, int(n_cells_out), int(n_genes_out)
        assert Xc.shape == (self.N, self.G)

    def export_matrix_csr(self, device=None):
        ip, ix, dt = self.Xc.indptr.astype(np.int32), self.Xc.indices.astype(np.int32), self.Xc.data.astype(np.float32)
        if device is None:
            return ip, ix, dt
        import torch
        return torch.from_numpy(ip), torch.from_numpy(ix), torch.from_numpy(dt)

    def set_matrix_csr(self, indptr, indices, data, n_cells, n_genes):
        from scipy import sparse
        to_np = lambda a: a.numpy() if hasattr(a, "numpy") else np.asarray(a)      # noqa: E731
        Xc = sparse.csr_matrix((to_np(data), to_np(indices), to_np(indptr)), shape=(int(n_cells), int(n_genes)))
        Xc.sort_indices()
        self.Xc, self.N, self.G = Xc, int(n_cells), int(n_genes)

    def set_labels(self, labels, n_clusters):
        self.labels, self.L = np.ascontiguousarray(labels, dtype=np.int32), int(n_clusters)

    def set_triples(self, lig, rec, src, tgt, obs_cpdb=None, obs_gmean=None):
        self.tri = tuple(np.ascontiguousarray(a, dtype=np.int32) for a in (lig, rec, src, tgt))
        self.obs_c, self.obs_g, self.T = obs_cpdb, obs_gmean, len(self.tri[0])

    # -- observed statistics ------------------------------------------------------------------
    def observed(self):
        means, stored, sizes = c_oracle.observed(self.Xc, self.labels, self.L)
        return means, stored.astype(np.int32), sizes.astype(np.int32)

    def observed_trimean(self, mat_max):
        return c_oracle.trimean_cube(self.Xc, self.labels, self.L, None, mat_max)[0]

    def cluster_moments(self, logbase=float(np.e)):
        outs = [np.zeros((self.L, self.G)) for _ in range(3)]
        X64 = self.Xc.astype(np.float64)
        E1 = X64.copy()
        E1.data = np.power(float(logbase), X64.data) - 1.0
        for c in range(self.L):
            rows = np.flatnonzero(self.labels == c)
            outs[0][c] = np.asarray(X64[rows].sum(axis=0)).ravel()
            outs[1][c] = np.asarray(X64[rows].multiply(X64[rows]).sum(axis=0)).ravel()
            outs[2][c] = np.asarray(E1[rows].sum(axis=0)).ravel()
        return tuple(outs)

    # -- permutation nulls --------------------------------------------------------------------
    def _perms(self, P, seed, perm_offset, perm_idx):
        if perm_idx is not None:
            return np.asarray(perm_idx, dtype=np.int64)
        return c_oracle.philox_perms(self.N, int(P), int(perm_offset), int(seed)).astype(np.int64)

    def run(self, n_perms, *, seed=1337, perm_offset=0, perm_idx=None, order=E.ORDER_EXACT, scores=E.SCORE_CPDB,
            gmean_mode=E.GMEAN_PRODUCT, return_cube=False, counts_out=None):
        out = {}
        if n_perms == 0 or self.T == 0:
            cube = np.zeros((0, self.L, self.G), dtype=np.float32)
        else:
            cube = c_oracle.cube(self.Xc, self.labels, self.L, self._perms(n_perms, seed, perm_offset, perm_idx))
        for flag, name, obs, method in ((E.SCORE_CPDB, "cpdb", self.obs_c, "cellphonedb"), (E.SCORE_GMEAN, "gmean", self.obs_g, "geometric_mean")):
            if scores & flag:
                with np.errstate(divide="ignore", invalid="ignore"):
                    cnt = numpy_port.counts_from_cube(cube, *self.tri, obs, method) if len(cube) else np.zeros(self.T)
                out[name] = np.asarray(cnt).astype(np.uint32)
        if return_cube:
            out["cube"] = cube
        return out

    def run_trimean(self, n_perms, mat_max, *, obs_prob=None, kh=0.5, seed=1337, perm_offset=0, perm_idx=None, return_cube=False):
        assert kh == numpy_port.KH
        cube = c_oracle.trimean_cube(self.Xc, self.labels, self.L, self._perms(n_perms, seed, perm_offset, perm_idx), mat_max)
        out = {"cube": cube} if return_cube else {}
        if obs_prob is not None:
            out["cellchat"] = np.asarray(numpy_port.counts_cellchat(cube, *self.tri, obs_prob)).astype(np.uint32)
        return out

    # -- host-logic kernels -------------------------------------------------------------------
    def resolve_triples(self, lig_sub, rec_sub, stat, props, expr_prop, pair_mask=None, first_subunit=False, return_all=False):
        tables = _tables.ResourceTables(None, None, np.asarray(lig_sub), np.asarray(rec_sub), None)
        t = _tables.resolve_triples(tables, stat, props, expr_prop, pair_mask, return_all, engine=None)
        return dict(inter=t.inter, src=t.src, tgt=t.tgt, lig_gene=t.lig_gene, rec_gene=t.rec_gene, lig_stat=t.ligand_means,
                    rec_stat=t.receptor_means, lig_props=t.ligand_props, rec_props=t.receptor_props, keep=t.keep)

    def rank_average(self, values, negate=False):
        v = np.asarray(values, dtype=np.float64)
        return rankdata(-v if negate else v, method="average").astype(np.float64)

    def rra(self, rmat, col_max=None):
        rmat = np.asarray(rmat, dtype=np.float64)
        return _aggregate.robust_rank_aggregate(rmat) if rmat.shape[0] else np.zeros(0)

    def argsort(self, values, descending=False):
        v = np.asarray(values, dtype=np.float64)
        return np.argsort(-v if descending else v, kind="stable").astype(np.int32)
