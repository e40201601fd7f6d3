"""Synthetic inputs of the shapes named in BASELINE.json (recipe: SURVEY.md §8d).

Genes: the `consensus` resource's subunit symbols (so the real resource, complexes and the
gene-subset step are exercised) plus background genes ``BG{i}``.  Per-gene expression
probability ``d_g ~ Beta(0.6, 0.6(1-d)/d)`` (mean d), optionally modulated per (cluster, gene)
by ``LogNormal(0, 0.5)``; values ``log1p(Gamma(1.5, 2.0))`` float32; labels uniform over L.
CSR, sorted indices, no explicit zeros, int32 indices.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
from scipy import sparse

from ._anndata_lite import AnnDataLite

_RES_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "resource")


def consensus_genes() -> np.ndarray:
    """Sorted distinct subunit symbols of the bundled consensus resource (1,877 genes)."""
    res = pd.read_csv(os.path.join(_RES_DIR, "consensus.csv"))
    genes = set()
    for col in ("ligand", "receptor"):
        for entry in res[col].astype(str):
            genes.update(entry.split("_"))
    return np.array(sorted(genes))


def _gene_names(n_genes_total: int) -> np.ndarray:
    cg = consensus_genes()
    if n_genes_total <= len(cg):
        return cg[:n_genes_total]
    bg = np.array([f"BG{i:05d}" for i in range(n_genes_total - len(cg))])
    return np.concatenate([cg, bg])


def synthetic_csr(n_cells: int, n_genes: int, n_clusters: int, density: float = 0.08,
                  seed_data: int = 1337, seed_labels: int = 1338, cluster_effect: bool = True,
                  chunk: int = 8192):
    """Return (X csr float32 [n_cells, n_genes], labels int32 [n_cells])."""
    rng = np.random.default_rng(seed_data)
    rng_l = np.random.default_rng(seed_labels)
    labels = rng_l.integers(0, n_clusters, size=n_cells).astype(np.int32)
    d_g = rng.beta(0.6, 0.6 * (1.0 - density) / density, size=n_genes)
    if cluster_effect:
        eff = rng.lognormal(0.0, 0.5, size=(n_clusters, n_genes))
        d_cg = np.clip(d_g[None, :] * eff, 0.0, 1.0).astype(np.float32)
    else:
        d_cg = np.repeat(d_g[None, :], n_clusters, axis=0).astype(np.float32)
    indptr = [np.zeros(1, dtype=np.int64)]
    indices = []
    data = []
    total = 0
    for s in range(0, n_cells, chunk):
        e = min(n_cells, s + chunk)
        u = rng.random((e - s, n_genes), dtype=np.float32)
        mask = u < d_cg[labels[s:e]]
        r, c = np.nonzero(mask)
        cnt = np.bincount(r, minlength=e - s)
        indptr.append(total + np.cumsum(cnt))
        total += len(c)
        indices.append(c.astype(np.int32))
        data.append(np.log1p(rng.gamma(1.5, 2.0, size=len(c))).astype(np.float32))
    indptr = np.concatenate(indptr)
    indices = np.concatenate(indices) if indices else np.zeros(0, np.int32)
    data = np.concatenate(data) if data else np.zeros(0, np.float32)
    idt = np.int32 if total < 2 ** 31 else np.int64
    X = sparse.csr_matrix((data, indices, indptr.astype(idt)), shape=(n_cells, n_genes))
    return X, labels


def make_synthetic_adata(n_cells: int, n_genes_total: int, n_clusters: int, density: float = 0.08,
                         seed_data: int = 1337, seed_labels: int = 1338,
                         cluster_effect: bool = True, groupby: str = "cluster",
                         n_samples: int = 0) -> AnnDataLite:
    """AnnData-like object with consensus gene symbols + background genes."""
    X, labels = synthetic_csr(n_cells, n_genes_total, n_clusters, density, seed_data,
                              seed_labels, cluster_effect)
    names = _gene_names(n_genes_total)
    rng = np.random.default_rng(seed_labels + 7)
    # shuffle gene order so that the alphabetical re-sort in preprocessing is exercised
    order = rng.permutation(n_genes_total)
    X = X[:, order].tocsr()
    X.sort_indices()
    var = pd.DataFrame(index=pd.Index(names[order]))
    cats = [f"c{i:02d}" for i in range(n_clusters)]
    obs = pd.DataFrame({groupby: pd.Categorical.from_codes(labels, categories=cats)},
                       index=pd.Index([f"cell{i}" for i in range(n_cells)]))
    if n_samples:
        obs["sample"] = pd.Categorical(
            [f"S{i:02d}" for i in rng.integers(0, n_samples, size=n_cells)])
    return AnnDataLite(X=X, obs=obs, var=var)
