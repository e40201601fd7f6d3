"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/lrperm_oracle.c (see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "lrperm_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "lrperm_philox.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _csr_parts(X):
    indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(X.indices, dtype=np.int32)
    data = np.ascontiguousarray(X.data, dtype=np.float32)
    return indptr, indices, data


def cube(X, labels, n_clusters, perm_idx, threads: int | None = None) -> np.ndarray:
    """float32 cube [P, L, G]; perm_idx int64 [P, N] (the reference's rng.permutation draws)."""
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    N, G = X.shape
    indptr, indices, data = _csr_parts(X)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int64)
    P = perm_idx.shape[0]
    assert perm_idx.shape == (P, N)
    out = np.empty((P, n_clusters, G), dtype=np.float32)
    rc = lib().oracle_cube(C.c_int64(N), C.c_int64(G), _p(indptr, C.c_int64), _p(indices, C.c_int32),
                           _p(data, C.c_float), _p(labels, C.c_int32), C.c_int32(n_clusters), C.c_int64(P),
                           _p(perm_idx, C.c_int64), _p(out, C.c_float))
    if rc:
        raise RuntimeError(f"oracle_cube failed rc={rc}")
    return out


def observed(X, labels, n_clusters):
    N, G = X.shape
    indptr, indices, data = _csr_parts(X)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    means = np.empty((n_clusters, G), dtype=np.float32)
    stored = np.empty((n_clusters, G), dtype=np.int32)
    sizes = np.empty(n_clusters, dtype=np.int32)
    rc = lib().oracle_observed(C.c_int64(N), C.c_int64(G), _p(indptr, C.c_int64), _p(indices, C.c_int32),
                               _p(data, C.c_float), _p(labels, C.c_int32), C.c_int32(n_clusters),
                               _p(means, C.c_float), _p(stored, C.c_int32), _p(sizes, C.c_int32))
    if rc:
        raise RuntimeError(f"oracle_observed failed rc={rc}")
    return means, stored, sizes


def counts_cpdb(cube_f32, lig, rec, src, tgt, obs) -> np.ndarray:
    P, L, G = cube_f32.shape
    cube_f32 = np.ascontiguousarray(cube_f32, dtype=np.float32)
    lig, rec, src, tgt = (np.ascontiguousarray(a, dtype=np.int32) for a in (lig, rec, src, tgt))
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    T = len(lig)
    out = np.zeros(T, dtype=np.uint32)
    rc = lib().oracle_counts_cpdb(C.c_int64(P), C.c_int32(L), C.c_int64(G), _p(cube_f32, C.c_float), C.c_int64(T),
                                  _p(lig, C.c_int32), _p(rec, C.c_int32), _p(src, C.c_int32), _p(tgt, C.c_int32),
                                  _p(obs, C.c_float), _p(out, C.c_uint32))
    if rc:
        raise RuntimeError(f"oracle_counts_cpdb failed rc={rc}")
    return out


def philox_perms(n_cells, n_perms, perm_offset, seed) -> np.ndarray:
    out = np.empty((n_perms, n_cells), dtype=np.int32)
    lib().oracle_philox_perms(C.c_int64(n_cells), C.c_int64(n_perms), C.c_int64(perm_offset), C.c_uint64(seed),
                              _p(out, C.c_int32))
    return out


def philox_inverse(n_cells, perm_id, seed) -> np.ndarray:
    out = np.empty(n_cells, dtype=np.int32)
    lib().oracle_philox_inverse(C.c_int64(n_cells), C.c_int64(perm_id), C.c_uint64(seed), _p(out, C.c_int32))
    return out


def trimean_cube(X, labels, n_clusters, perm_idx, mat_max, threads: int | None = None) -> np.ndarray:
    """float64 trimean cube [P, L, G].  perm_idx int64 [P, N]: the null (float32 values x * fl32(1/mat_max));
    perm_idx None: the observed statistic [1, L, G] (float64 values x * (1/mat_max)).  The reciprocal is
    evaluated here with NumPy scalar semantics, like SciPy does inside the reference's two divisions."""
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    N, G = X.shape
    indptr, indices, data = _csr_parts(X)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    recip = 1.0 / np.float32(mat_max)
    if perm_idx is None:
        P, pp, observed = 1, None, 1
    else:
        perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int64)
        P, pp, observed = perm_idx.shape[0], _p(perm_idx, C.c_int64), 0
        assert perm_idx.shape == (P, N)
    out = np.empty((P, n_clusters, G), dtype=np.float64)
    rc = lib().oracle_trimean_cube(C.c_int64(N), C.c_int64(G), _p(indptr, C.c_int64), _p(indices, C.c_int32),
                                   _p(data, C.c_float), _p(labels, C.c_int32), C.c_int32(n_clusters), C.c_int64(P),
                                   pp, C.c_int(observed), C.c_float(float(np.float32(recip))), C.c_double(float(recip)),
                                   _p(out, C.c_double))
    if rc:
        raise RuntimeError(f"oracle_trimean_cube failed rc={rc}")
    return out


def counts_cellchat(cube_f64, lig, rec, src, tgt, obs, kh=0.5) -> np.ndarray:
    P, L, G = cube_f64.shape
    cube_f64 = np.ascontiguousarray(cube_f64, dtype=np.float64)
    lig, rec, src, tgt = (np.ascontiguousarray(a, dtype=np.int32) for a in (lig, rec, src, tgt))
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    T = len(lig)
    out = np.zeros(T, dtype=np.uint32)
    rc = lib().oracle_counts_cellchat(C.c_int64(P), C.c_int32(L), C.c_int64(G), _p(cube_f64, C.c_double), C.c_int64(T),
                                      _p(lig, C.c_int32), _p(rec, C.c_int32), _p(src, C.c_int32), _p(tgt, C.c_int32),
                                      _p(obs, C.c_double), C.c_double(kh), _p(out, C.c_uint32))
    if rc:
        raise RuntimeError(f"oracle_counts_cellchat failed rc={rc}")
    return out
