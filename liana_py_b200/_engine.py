"""ctypes binding of the C ABI in include/lrperm.h (the library built from csrc/lrperm_api.cu).

This is the only way the Python layer reaches the GPU: there is no CPU fallback.  If the
shared library is missing or no CUDA device is present, construction raises loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# LRPERM_LIB: another build of the same ABI (e.g. the bounds-checking debug build, `python -m liana_py_b200.build --debug`)
LIB_PATH = os.environ.get("LRPERM_LIB") or os.path.join(_PKG_DIR, "_lib", "liblrperm.so")

PERMS_INDEX, PERMS_PHILOX = 0, 1
ORDER_EXACT, ORDER_FAST = 0, 1
SCORE_CPDB, SCORE_GMEAN = 1, 2
GMEAN_PRODUCT, GMEAN_LITERAL = 0, 1

# every symbol include/lrperm.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "lrperm_abi_version", "lrperm_last_error", "lrperm_create", "lrperm_destroy", "lrperm_set_stream",
    "lrperm_set_matrix_csr", "lrperm_set_labels", "lrperm_set_triples", "lrperm_observed", "lrperm_run",
    "lrperm_philox_perms", "lrperm_last_timings", "lrperm_set_tuning", "lrperm_set_option",
    "lrperm_cluster_moments", "lrperm_stage_matrix_csr", "lrperm_commit_matrix",
    "lrperm_run_trimean", "lrperm_observed_trimean", "lrperm_trimean_timings", "lrperm_staged_max",
    "lrperm_rank_average", "lrperm_rra", "lrperm_resolve_triples", "lrperm_fetch_triples",
    "lrperm_argsort", "lrperm_staged_cluster_stats", "lrperm_stage_matrix_dense",
    "lrperm_run_stats", "lrperm_matrix_info", "lrperm_export_matrix_csr", "lrperm_debug_bounds", "lrperm_staged_info",
)

_lib = None


class EngineError(RuntimeError):
    pass


def load_library():
    """Load liblrperm.so (built in-tree by ``__graft_entry__.build()`` / ``liana_py_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"CUDA extension not built: {LIB_PATH} is missing. Run `python -m liana_py_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the permutation engine.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.lrperm_abi_version.restype = C.c_int
    lib.lrperm_last_error.restype = C.c_char_p
    lib.lrperm_last_error.argtypes = [vp]
    lib.lrperm_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.lrperm_destroy.argtypes = [vp]
    lib.lrperm_destroy.restype = None
    lib.lrperm_set_stream.argtypes = [vp, vp]
    lib.lrperm_set_tuning.argtypes = [vp, i32, i32, i32]
    lib.lrperm_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.lrperm_cluster_moments.argtypes = [vp, C.c_double, vp, vp, vp, C.c_int]
    lib.lrperm_stage_matrix_csr.argtypes = [vp, i64, i64, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    lib.lrperm_stage_matrix_dense.argtypes = [vp, i64, i64, vp, C.c_int, i64, C.c_int, vp, vp, vp, vp, vp]
    lib.lrperm_commit_matrix.argtypes = [vp, vp, i64]
    lib.lrperm_set_matrix_csr.argtypes = [vp, i64, i64, vp, C.c_int, vp, vp, C.c_int]
    lib.lrperm_set_labels.argtypes = [vp, vp, i32, C.c_int]
    lib.lrperm_set_triples.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.lrperm_observed.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.lrperm_run.argtypes = [vp, i64, i64, u64, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               vp, vp, C.c_int, vp, C.c_int]
    lib.lrperm_philox_perms.argtypes = [vp, i64, i64, i64, u64, vp, C.c_int]
    lib.lrperm_last_timings.argtypes = [vp, vp]
    lib.lrperm_run_trimean.argtypes = [vp, i64, i64, u64, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_double,
                                       vp, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.lrperm_observed_trimean.argtypes = [vp, C.c_double, vp, C.c_int]
    lib.lrperm_trimean_timings.argtypes = [vp, vp]
    lib.lrperm_staged_max.argtypes = [vp, vp, vp]
    lib.lrperm_rank_average.argtypes = [vp, i64, vp, C.c_int, vp, C.c_int]
    lib.lrperm_rra.argtypes = [vp, i64, i32, vp, vp, vp, C.c_int]
    lib.lrperm_argsort.argtypes = [vp, i64, vp, C.c_int, vp, C.c_int]
    lib.lrperm_staged_cluster_stats.argtypes = [vp, vp, i32, vp, vp]
    lib.lrperm_resolve_triples.argtypes = [vp, i32, i32, i32, vp, vp, vp, C.c_int, vp, C.c_double, vp, C.c_int, C.c_int, vp]
    lib.lrperm_fetch_triples.argtypes = [vp] + [vp] * 10
    lib.lrperm_run_stats.argtypes = [vp, vp]
    lib.lrperm_matrix_info.argtypes = [vp, vp]
    lib.lrperm_export_matrix_csr.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.lrperm_debug_bounds.argtypes = [vp, vp, vp]
    lib.lrperm_staged_info.argtypes = [vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _ptr(x):
    """(address, on_device, keepalive) for a numpy array / torch tensor / None."""
    if x is None:
        return None, 0, None
    if _is_torch(x):
        if not x.is_contiguous():
            x = x.contiguous()
        return C.c_void_p(x.data_ptr()), int(x.is_cuda), x
    return C.c_void_p(x.ctypes.data), 0, x


class Engine:
    """One engine = one CUDA device + one stream (see include/lrperm.h)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.lrperm_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = self._lib.lrperm_last_error(None).decode()
            self._h = None
            raise EngineError(f"lrperm_create failed (status {rc}): {msg}")
        self.device = int(device)
        self.N = self.G = self.L = self.T = 0

    # -- plumbing --------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.lrperm_last_error(self._h).decode()
            if rc == 1:
                raise ValueError(msg)
            raise EngineError(f"status {rc}: {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lrperm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        self._check(self._lib.lrperm_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_option(self, name: str, value: int):
        self._check(self._lib.lrperm_set_option(self._h, name.encode(), int(value)))

    def set_tuning(self, fast_perm_batch: int = 0, fast_chunk_nnz: int = 0, fast_tile_mib: int = 0):
        self._check(self._lib.lrperm_set_tuning(self._h, int(fast_perm_batch), int(fast_chunk_nnz), int(fast_tile_mib)))

    # -- inputs ----------------------------------------------------------------------------
    def set_matrix_csr(self, indptr, indices, data, n_cells: int, n_genes: int):
        """CSR float32 N x G.  numpy (host) or torch CUDA tensors (device) - all three alike."""
        if _is_torch(indptr):
            import torch
            is64 = indptr.dtype == torch.int64
            assert indptr.dtype in (torch.int32, torch.int64) and indices.dtype == torch.int32 and data.dtype == torch.float32
            assert indptr.is_cuda == indices.is_cuda == data.is_cuda
        else:
            indptr = np.ascontiguousarray(indptr)
            if indptr.dtype not in (np.int32, np.int64):
                indptr = indptr.astype(np.int64)
            is64 = indptr.dtype == np.int64
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            data = np.ascontiguousarray(data, dtype=np.float32)
        p0, dev, k0 = _ptr(indptr)
        p1, _, k1 = _ptr(indices)
        p2, _, k2 = _ptr(data)
        self._check(self._lib.lrperm_set_matrix_csr(self._h, int(n_cells), int(n_genes), p0, int(is64), p1, p2, dev))
        # pinned host buffers are uploaded asynchronously (include/lrperm.h): keep them alive until the next matrix
        self._matrix_keepalive = (k0, k1, k2)
        self.N, self.G = int(n_cells), int(n_genes)

    def set_matrix(self, X):
        """SciPy CSR matrix (canonical format: sorted, no duplicate entries per row)."""
        from scipy import sparse
        if not sparse.isspmatrix_csr(X):
            X = sparse.csr_matrix(X)
        if not X.has_canonical_format:
            X = X.copy()
            X.sum_duplicates()
        self.set_matrix_csr(X.indptr, X.indices, X.data, X.shape[0], X.shape[1])

    def stage_matrix(self, X, row_keep=None):
        """Upload the caller's FULL SciPy CSR matrix (float32) and return its statistics:
        dict(col_sums float64 [G], kept_total, n_nonfinite, n_empty_rows, n_unordered).  n_unordered > 0: some row is not
        strictly increasing in its column indices (unsorted or duplicate entries) - canonicalise (`sum_duplicates`) and
        stage again.  Follow with commit_matrix."""
        indptr = np.ascontiguousarray(X.indptr)
        if indptr.dtype not in (np.int32, np.int64):
            indptr = indptr.astype(np.int64)
        indices = np.ascontiguousarray(X.indices, dtype=np.int32)
        data = np.ascontiguousarray(X.data, dtype=np.float32)
        keep = None if row_keep is None else np.ascontiguousarray(row_keep, dtype=np.uint8)
        n, g = X.shape
        col_sums = np.zeros(g, dtype=np.float64)
        total, bad, empty = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
        self._check(self._lib.lrperm_stage_matrix_csr(
            self._h, int(n), int(g), C.c_void_p(indptr.ctypes.data), int(indptr.dtype == np.int64),
            C.c_void_p(indices.ctypes.data), C.c_void_p(data.ctypes.data), 0,
            None if keep is None else C.c_void_p(keep.ctypes.data), C.c_void_p(col_sums.ctypes.data),
            C.byref(total), C.byref(bad), C.byref(empty)))
        return dict(col_sums=col_sums, kept_total=total.value, n_nonfinite=bad.value, n_empty_rows=empty.value,
                    n_unordered=self.staged_info()["n_unordered"])

    def stage_matrix_device(self, indptr, indices, data, shape, row_keep=None):
        """`stage_matrix` for CSR arrays that already live on the device (CUDA torch tensors: indptr int32 / int64,
        indices int32, data float32; row_keep uint8 or None).  The staged matrix may hold more than 2^31 entries."""
        import torch
        if not (indptr.is_cuda and indices.is_cuda and data.is_cuda):
            raise ValueError("stage_matrix_device takes CUDA tensors")
        if indptr.dtype not in (torch.int32, torch.int64) or indices.dtype != torch.int32 or data.dtype != torch.float32:
            raise ValueError("indptr must be int32 / int64, indices int32, data float32")
        indptr, indices, data = indptr.contiguous(), indices.contiguous(), data.contiguous()
        keep = None if row_keep is None else row_keep.to(device=data.device, dtype=torch.uint8).contiguous()
        n, g = int(shape[0]), int(shape[1])
        col_sums = np.zeros(g, dtype=np.float64)
        total, bad, empty = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
        torch.cuda.current_stream().synchronize()            # the engine runs on its own stream
        self._check(self._lib.lrperm_stage_matrix_csr(
            self._h, n, g, C.c_void_p(indptr.data_ptr()), int(indptr.dtype == torch.int64),
            C.c_void_p(indices.data_ptr()), C.c_void_p(data.data_ptr()), 1,
            None if keep is None else C.c_void_p(keep.data_ptr()), C.c_void_p(col_sums.ctypes.data),
            C.byref(total), C.byref(bad), C.byref(empty)))
        return dict(col_sums=col_sums, kept_total=total.value, n_nonfinite=bad.value, n_empty_rows=empty.value,
                    n_unordered=self.staged_info()["n_unordered"])

    def stage_matrix_dense(self, X, row_keep=None):
        """`stage_matrix` for a dense 2-D array (NumPy float32 / float64, or a CUDA torch tensor of those types): the
        CSR conversion (`csr_matrix(X)`: entries != 0, column order) happens on the device."""
        on_device = _is_torch(X)
        if on_device:
            if X.dim() != 2 or str(X.dtype) not in ("torch.float32", "torch.float64") or not X.is_cuda:
                raise ValueError("dense device input must be a 2-D CUDA tensor of float32 or float64")
            if X.stride(1) != 1 and X.shape[1] > 1:
                X = X.contiguous()
            ptr, stride, is64 = X.data_ptr(), (X.stride(0) if X.shape[0] > 1 else X.shape[1]), X.dtype.itemsize == 8
            if row_keep is not None:
                import torch
                keep = torch.as_tensor(np.ascontiguousarray(row_keep, dtype=np.uint8), device=X.device)
                keep_ptr = keep.data_ptr()
        else:
            X = np.asarray(X)
            if X.ndim != 2:
                raise ValueError("dense input must be 2-D")
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float32)
            if not X.flags.c_contiguous:
                X = np.ascontiguousarray(X)
            ptr, stride, is64 = X.ctypes.data, X.shape[1], X.dtype == np.float64
            keep = None if row_keep is None else np.ascontiguousarray(row_keep, dtype=np.uint8)
            keep_ptr = None if keep is None else keep.ctypes.data
        if row_keep is None:
            keep_ptr = None
        n, g = int(X.shape[0]), int(X.shape[1])
        col_sums = np.zeros(g, dtype=np.float64)
        total, bad, empty = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
        if on_device:
            import torch
            torch.cuda.current_stream().synchronize()        # the engine runs on its own stream
        self._check(self._lib.lrperm_stage_matrix_dense(
            self._h, n, g, C.c_void_p(ptr), int(is64), int(max(stride, g)), int(on_device),
            None if keep_ptr is None else C.c_void_p(keep_ptr), C.c_void_p(col_sums.ctypes.data),
            C.byref(total), C.byref(bad), C.byref(empty)))
        return dict(col_sums=col_sums, kept_total=total.value, n_nonfinite=bad.value, n_empty_rows=empty.value,
                    n_unordered=self.staged_info()["n_unordered"])

    def staged_info(self) -> dict:
        """dict(n_cells, n_genes, nnz, n_unordered) of the staged matrix (`lrperm_staged_info`)."""
        vals = [C.c_int64(0) for _ in range(4)]
        self._check(self._lib.lrperm_staged_info(self._h, *[C.byref(v) for v in vals]))
        return dict(zip(("n_cells", "n_genes", "nnz", "n_unordered"), (int(v.value) for v in vals)))

    def staged_max(self):
        """(largest stored value over the kept rows of the staged matrix as np.float32, stored entries in them)."""
        mx, cnt = C.c_float(0.0), C.c_int64(0)
        self._check(self._lib.lrperm_staged_max(self._h, C.byref(mx), C.byref(cnt)))
        return np.float32(mx.value), int(cnt.value)

    def staged_cluster_stats(self, row_label, n_clusters: int):
        """(sum_x, sum_sq) float64 [L] over all stored entries of the staged matrix per cluster (row_label -1 = dropped)."""
        lab = np.ascontiguousarray(row_label, dtype=np.int32)
        sx, sq = np.zeros(n_clusters, dtype=np.float64), np.zeros(n_clusters, dtype=np.float64)
        self._check(self._lib.lrperm_staged_cluster_stats(self._h, C.c_void_p(lab.ctypes.data), int(n_clusters),
                                                          C.c_void_p(sx.ctypes.data), C.c_void_p(sq.ctypes.data)))
        return sx, sq

    def commit_matrix(self, col_map, n_genes_out: int, n_cells_out: int):
        """Make the staged matrix the engine's matrix: rows flagged in row_keep, columns with col_map >= 0."""
        col_map = np.ascontiguousarray(col_map, dtype=np.int32)
        self._check(self._lib.lrperm_commit_matrix(self._h, C.c_void_p(col_map.ctypes.data), int(n_genes_out)))
        self.N, self.G = int(n_cells_out), int(n_genes_out)

    def set_labels(self, labels, n_clusters: int):
        if not _is_torch(labels):
            labels = np.ascontiguousarray(labels, dtype=np.int32)
        p, dev, _k = _ptr(labels)
        self._check(self._lib.lrperm_set_labels(self._h, p, int(n_clusters), dev))
        self.L = int(n_clusters)

    def set_triples(self, ligand_idx, receptor_idx, source_idx, target_idx, obs_cpdb=None, obs_gmean=None):
        arrs = [ligand_idx, receptor_idx, source_idx, target_idx]
        if not _is_torch(ligand_idx):
            arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in arrs]
            obs_cpdb = None if obs_cpdb is None else np.ascontiguousarray(obs_cpdb, dtype=np.float32)
            obs_gmean = None if obs_gmean is None else np.ascontiguousarray(obs_gmean, dtype=np.float32)
        T = int(arrs[0].shape[0])
        ptrs = [_ptr(a) for a in arrs]
        pc, _, kc = _ptr(obs_cpdb)
        pg, _, kg = _ptr(obs_gmean)
        self._check(self._lib.lrperm_set_triples(self._h, T, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0], pc, pg,
                                                 ptrs[0][1]))
        self.T = T

    # -- outputs ---------------------------------------------------------------------------
    def observed(self):
        """(means float32 [L,G], stored_counts int32 [L,G], cluster_sizes int32 [L])."""
        means = np.empty((self.L, self.G), dtype=np.float32)
        stored = np.empty((self.L, self.G), dtype=np.int32)
        sizes = np.empty(self.L, dtype=np.int32)
        self._check(self._lib.lrperm_observed(self._h, C.c_void_p(means.ctypes.data), C.c_void_p(stored.ctypes.data),
                                              C.c_void_p(sizes.ctypes.data), 0))
        return means, stored, sizes

    def cluster_moments(self, logbase: float = float(np.e)):
        """(sum_x, sum_sq, sum_expm1) float64 [L,G]: per-cluster sums of x, x^2 and logbase^x - 1."""
        outs = [np.empty((self.L, self.G), dtype=np.float64) for _ in range(3)]
        self._check(self._lib.lrperm_cluster_moments(self._h, float(logbase), *(C.c_void_p(o.ctypes.data) for o in outs), 0))
        return tuple(outs)

    def run(self, n_perms: int, *, seed: int = 1337, perm_offset: int = 0, perm_idx=None, order: int = ORDER_EXACT,
            scores: int = SCORE_CPDB, gmean_mode: int = GMEAN_PRODUCT, return_cube: bool = False,
            counts_out=None):
        """Run the permutation null.  perm_idx given -> the reference's own index array
        (LRPERM_PERMS_INDEX); None -> Philox on device.  Returns dict with 'cpdb' / 'gmean' uint32
        count vectors (numpy, or the torch CUDA tensors passed in counts_out) and optionally 'cube'."""
        P = int(n_perms)
        if perm_idx is not None:
            src = PERMS_INDEX
            if not _is_torch(perm_idx):
                perm_idx = np.ascontiguousarray(perm_idx)
                if perm_idx.dtype not in (np.int32, np.int64):
                    perm_idx = perm_idx.astype(np.int64)
                is64 = perm_idx.dtype == np.int64
            else:
                import torch
                is64 = perm_idx.dtype == torch.int64
            if tuple(perm_idx.shape) != (P, self.N):
                raise ValueError(f"perm_idx must have shape ({P}, {self.N}), got {tuple(perm_idx.shape)}")
        else:
            src, is64 = PERMS_PHILOX, 0
        pp, pdev, _kp = _ptr(perm_idx)
        out = {}
        on_dev = 0
        cc = cg = None
        if counts_out is not None:
            cc, cg = counts_out
            on_dev = 1
        else:
            if scores & SCORE_CPDB:
                cc = np.zeros(self.T, dtype=np.uint32)
            if scores & SCORE_GMEAN:
                cg = np.zeros(self.T, dtype=np.uint32)
        pc, _, _kc = _ptr(cc if scores & SCORE_CPDB else None)
        pg, _, _kg = _ptr(cg if scores & SCORE_GMEAN else None)
        cube = np.empty((P, self.L, self.G), dtype=np.float32) if return_cube else None
        pcube, _, _kcube = _ptr(cube)
        self._check(self._lib.lrperm_run(self._h, P, int(perm_offset), C.c_uint64(int(seed) & (2 ** 64 - 1)), src,
                                         pp, int(is64), pdev, int(order), int(scores), int(gmean_mode),
                                         pc, pg, on_dev, pcube, 0))
        if scores & SCORE_CPDB:
            out["cpdb"] = cc
        if scores & SCORE_GMEAN:
            out["gmean"] = cg
        if return_cube:
            out["cube"] = cube
        return out

    # -- CellChat: trimean null -------------------------------------------------------------
    @staticmethod
    def trimean_reciprocals(mat_max):
        """(recip32, recip) the way SciPy evaluates the reference's two divisions by the float32 `mat_max`
        (`adata.X /= norm_factor` -> `data *= 1.0 / norm`, float32; `temp.X / mat_max` -> float64 matrix times
        `1 / mat_max`): NumPy scalar semantics decide whether `1 / np.float32` is float32 (NumPy >= 2) or float64."""
        recip = 1.0 / np.float32(mat_max)
        return float(np.float32(recip)), float(recip)

    def observed_trimean(self, mat_max) -> np.ndarray:
        """float64 [L,G]: `_trimean(temp.X / mat_max)` per cluster (`_liana_pipe.py:307-308`)."""
        out = np.empty((self.L, self.G), dtype=np.float64)
        _, recip = self.trimean_reciprocals(mat_max)
        self._check(self._lib.lrperm_observed_trimean(self._h, recip, C.c_void_p(out.ctypes.data), 0))
        return out

    def run_trimean(self, n_perms: int, mat_max, *, obs_prob=None, kh: float = 0.5, seed: int = 1337, perm_offset: int = 0,
                    perm_idx=None, return_cube: bool = False):
        """CellChat null.  Returns dict with 'cellchat' uint32 counts (if obs_prob is given; needs set_triples)
        and optionally 'cube' float64 [P,L,G]."""
        P = int(n_perms)
        if perm_idx is not None:
            src = PERMS_INDEX
            perm_idx = np.ascontiguousarray(perm_idx)
            if perm_idx.dtype not in (np.int32, np.int64):
                perm_idx = perm_idx.astype(np.int64)
            is64 = perm_idx.dtype == np.int64
            if tuple(perm_idx.shape) != (P, self.N):
                raise ValueError(f"perm_idx must have shape ({P}, {self.N}), got {tuple(perm_idx.shape)}")
        else:
            src, is64 = PERMS_PHILOX, 0
        pp, pdev, _kp = _ptr(perm_idx)
        recip32, _ = self.trimean_reciprocals(mat_max)
        counts = None
        if obs_prob is not None:
            obs_prob = np.ascontiguousarray(obs_prob, dtype=np.float64)
            if obs_prob.shape != (self.T,):
                raise ValueError(f"obs_prob must have shape ({self.T},)")
            counts = np.zeros(self.T, dtype=np.uint32)
        po, _, _ko = _ptr(obs_prob)
        pc, _, _kc = _ptr(counts)
        cube = np.empty((P, self.L, self.G), dtype=np.float64) if return_cube else None
        pcube, _, _kcube = _ptr(cube)
        self._check(self._lib.lrperm_run_trimean(self._h, P, int(perm_offset), C.c_uint64(int(seed) & (2 ** 64 - 1)), src,
                                                 pp, int(is64), pdev, C.c_float(recip32), float(kh), po, 0, pc, 0, pcube, 0))
        out = {}
        if counts is not None:
            out["cellchat"] = counts
        if return_cube:
            out["cube"] = cube
        return out

    def trimean_timings(self) -> dict:
        buf = (C.c_float * 6)()
        self._check(self._lib.lrperm_trimean_timings(self._h, C.cast(buf, C.c_void_p)))
        return {"count_ms": buf[0], "prefix_ms": buf[1], "select_ms": buf[2], "finalize_ms": buf[3],
                "walked_warps": int(buf[4]), "select_warps": int(buf[5])}

    # -- complexes / expr_prop filter ---------------------------------------------------------
    def resolve_triples(self, lig_sub, rec_sub, stat, props, expr_prop, pair_mask=None, first_subunit=False, return_all=False):
        """Kept (source, target, interaction) rows with resolved subunits (include/lrperm.h).  Returns a dict of
        host arrays: inter, src, tgt, lig_gene, rec_gene (int32), lig_stat, rec_stat (dtype of `stat`), lig_props,
        rec_props (float64), keep (bool)."""
        lig_sub = np.ascontiguousarray(lig_sub, dtype=np.int32)
        rec_sub = np.ascontiguousarray(rec_sub, dtype=np.int32)
        if stat.dtype not in (np.float32, np.float64):
            stat = stat.astype(np.float64)
        stat = np.ascontiguousarray(stat)
        props = np.ascontiguousarray(props, dtype=np.float64)
        assert stat.shape == (self.L, self.G) and props.shape == (self.L, self.G)
        mask = None if pair_mask is None else np.ascontiguousarray(pair_mask, dtype=np.uint8)
        n = C.c_int64(0)
        self._check(self._lib.lrperm_resolve_triples(
            self._h, int(lig_sub.shape[0]), int(lig_sub.shape[1]), int(rec_sub.shape[1]), C.c_void_p(lig_sub.ctypes.data),
            C.c_void_p(rec_sub.ctypes.data), C.c_void_p(stat.ctypes.data), int(stat.dtype == np.float64),
            C.c_void_p(props.ctypes.data), float(expr_prop), None if mask is None else C.c_void_p(mask.ctypes.data),
            int(bool(first_subunit)), int(bool(return_all)), C.byref(n)))
        T = int(n.value)
        out = {k: np.empty(T, dtype=np.int32) for k in ("inter", "src", "tgt", "lig_gene", "rec_gene")}
        out.update({k: np.empty(T, dtype=stat.dtype) for k in ("lig_stat", "rec_stat")})
        out.update({k: np.empty(T, dtype=np.float64) for k in ("lig_props", "rec_props")})
        keep = np.empty(T, dtype=np.uint8)
        order = ("inter", "src", "tgt", "lig_gene", "rec_gene", "lig_stat", "rec_stat", "lig_props", "rec_props")
        self._check(self._lib.lrperm_fetch_triples(self._h, *(C.c_void_p(out[k].ctypes.data) for k in order),
                                                   C.c_void_p(keep.ctypes.data)))
        out["keep"] = keep.astype(bool)
        return out

    # -- consensus ---------------------------------------------------------------------------
    def rank_average(self, values, negate: bool = False) -> np.ndarray:
        """scipy.stats.rankdata(-values if negate else values, method='average') as float64 (no NaN in `values`)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty(v.shape[0], dtype=np.float64)
        self._check(self._lib.lrperm_rank_average(self._h, int(v.shape[0]), C.c_void_p(v.ctypes.data), int(bool(negate)),
                                                  C.c_void_p(out.ctypes.data), 0))
        return out

    def argsort(self, values, descending: bool = False) -> np.ndarray:
        """Stable argsort of a float column on the device (no NaN in `values`)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty(v.shape[0], dtype=np.int32)
        self._check(self._lib.lrperm_argsort(self._h, int(v.shape[0]), C.c_void_p(v.ctypes.data), int(bool(descending)),
                                             C.c_void_p(out.ctypes.data), 0))
        return out

    def rra(self, rmat, col_max=None) -> np.ndarray:
        """RobustRankAggregate rho per row of a rank matrix [n_rows, n_cols] (`_aggregate.py:110-181`); `col_max` =
        the column maxima when the caller has them (an axis-0 reduction of a row-major matrix is slow in NumPy)."""
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        if col_max is not None:
            col_max = np.ascontiguousarray(col_max, dtype=np.float64)
        else:
            col_max = np.ascontiguousarray(r.max(axis=0)) if r.shape[0] else np.ones(r.shape[1])
        out = np.empty(r.shape[0], dtype=np.float64)
        self._check(self._lib.lrperm_rra(self._h, int(r.shape[0]), int(r.shape[1]), C.c_void_p(r.ctypes.data),
                                         C.c_void_p(col_max.ctypes.data), C.c_void_p(out.ctypes.data), 0))
        return out

    def philox_perms(self, n_cells: int, n_perms: int, seed: int = 1337, perm_offset: int = 0) -> np.ndarray:
        out = np.empty((int(n_perms), int(n_cells)), dtype=np.int32)
        self._check(self._lib.lrperm_philox_perms(self._h, int(n_cells), int(n_perms), int(perm_offset),
                                                  C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_void_p(out.ctypes.data), 0))
        return out

    def debug_bounds(self):
        """(violation bits, (first bit, first value)) of the kernels' own bounds checks since the last call - debug build
        only (`LRPERM_LIB=.../liblrperm_debug.so`); the production library raises EngineError."""
        flags, first = C.c_uint32(0), (C.c_int64 * 2)()
        self._check(self._lib.lrperm_debug_bounds(self._h, C.byref(flags), C.cast(first, C.c_void_p)))
        return int(flags.value), (int(first[0]), int(first[1]))

    def run_stats(self) -> dict:
        """Work of the last FAST run: genes / stored entries actually accumulated (include/lrperm.h)."""
        buf = (C.c_int64 * 4)()
        self._check(self._lib.lrperm_run_stats(self._h, C.cast(buf, C.c_void_p)))
        return {"genes_scored": int(buf[0]), "nnz_scored": int(buf[1]), "chunks_run": int(buf[2]), "nnz": int(buf[3])}

    def matrix_info(self):
        """(cells, genes, stored entries) of the engine's matrix."""
        buf = (C.c_int64 * 3)()
        self._check(self._lib.lrperm_matrix_info(self._h, C.cast(buf, C.c_void_p)))
        return int(buf[0]), int(buf[1]), int(buf[2])

    def export_matrix_csr(self, device=None):
        """The engine's matrix as CSR.  device=None: NumPy arrays (indptr, indices, data); a torch CUDA device (possibly
        another GPU than the engine's): torch tensors there, copied device to device (peer copy over NVLink)."""
        n, g, nnz = self.matrix_info()
        if device is None:
            indptr = np.empty(n + 1, dtype=np.int32)
            indices, data = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.float32)
            self._check(self._lib.lrperm_export_matrix_csr(self._h, C.c_void_p(indptr.ctypes.data), C.c_void_p(indices.ctypes.data),
                                                           C.c_void_p(data.ctypes.data), 0))
            return indptr, indices, data
        import torch
        indptr = torch.empty(n + 1, dtype=torch.int32, device=device)
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)[:nnz]
        data = torch.empty(max(nnz, 1), dtype=torch.float32, device=device)[:nnz]
        torch.cuda.synchronize(device)
        self._check(self._lib.lrperm_export_matrix_csr(self._h, C.c_void_p(indptr.data_ptr()), C.c_void_p(indices.data_ptr()),
                                                       C.c_void_p(data.data_ptr()), 1))
        return indptr, indices, data

    def timings(self) -> dict:
        buf = (C.c_float * 8)()
        self._check(self._lib.lrperm_last_timings(self._h, C.cast(buf, C.c_void_p)))
        return {"setup_ms": buf[0], "means_ms": buf[1], "score_ms": buf[2], "total_ms": buf[3], "launches": int(buf[4]),
                "main_kernel_ms": buf[5], "main_kernel_launches": int(buf[6]), "perm_batch": int(buf[7])}
