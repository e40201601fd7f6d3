"""Vectorised construction of the tables the permutation kernels consume.

Replaces, with identical results, the reference's pandas pipeline between the per-cluster
statistics and the score function:

  explode_complexes              liana/resource/_reassemble_complexes.py:104-134
  filter_resource                liana/method/_pipe_utils/_pre.py:188-224
  _join_stats (L^2 merges)       liana/method/_pipe_utils/_common.py:5-39, _liana_pipe.py:310-321
  filter_reassemble_complexes    liana/resource/_reassemble_complexes.py:11-101
  _get_positions / _get_mat_idx  liana/method/_pipe_utils/_get_mean_perms.py:90-113  (O(T*G) there)

Key facts used: for a row (source s, target t, ligand complex LC, receptor
complex RC) the kept ligand subunit is the FIRST subunit (in the complex's '_' order) whose
mean in s is minimal, the receptor subunit likewise in t; the row survives if
min(all ligand-subunit props in s, all receptor-subunit props in t) >= expr_prop.
So everything factorises into [n_interactions, L] tables.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd


@dataclass
class ResourceTables:
    ligand_complex: np.ndarray     # [n_int] str
    receptor_complex: np.ndarray   # [n_int] str
    lig_sub: np.ndarray            # [n_int, max_l] gene index, -1 padded
    rec_sub: np.ndarray            # [n_int, max_r]
    entities: np.ndarray           # sorted gene names used by the kept interactions
    entity_index: np.ndarray = None    # [n_entities] int32: position of every kept gene in the resource's own entity list
    inter_index: np.ndarray = None     # [n_int] int32: position of every kept interaction in the resource


@dataclass
class ExplodedResource:
    """A resource split into subunits ONCE (the Python string work), independent of any dataset."""
    ligand_complex: np.ndarray     # [n_int] str, resource order
    receptor_complex: np.ndarray   # [n_int] str
    lig_all: np.ndarray            # [n_int, max_l] index into `entities`, -1 padded
    rec_all: np.ndarray            # [n_int, max_r]
    entities: np.ndarray           # sorted subunit symbols named anywhere in the resource


def explode_resource(resource: pd.DataFrame, sep: str = "_") -> ExplodedResource:
    lig = [str(x) for x in resource["ligand"].astype(str)]
    rec = [str(x) for x in resource["receptor"].astype(str)]
    lig_lists, rec_lists = [x.split(sep) for x in lig], [x.split(sep) for x in rec]
    ents = sorted({g for lst in lig_lists for g in lst} | {g for lst in rec_lists for g in lst})
    pos = {g: i for i, g in enumerate(ents)}
    n = len(lig)
    max_l = max((len(x) for x in lig_lists), default=1)
    max_r = max((len(x) for x in rec_lists), default=1)
    lig_all = np.full((n, max_l), -1, dtype=np.int32)
    rec_all = np.full((n, max_r), -1, dtype=np.int32)
    for i, (ls, rs) in enumerate(zip(lig_lists, rec_lists)):
        lig_all[i, :len(ls)] = [pos[g] for g in ls]
        rec_all[i, :len(rs)] = [pos[g] for g in rs]
    return ExplodedResource(np.array(lig, dtype=str), np.array(rec, dtype=str), lig_all, rec_all, np.array(ents, dtype=str))


def resource_entities(resource: pd.DataFrame, sep: str = "_") -> np.ndarray:
    """All subunit symbols named by the resource (np.union1d of exploded ligand / receptor)."""
    return explode_resource(resource, sep).entities


def tables_for_genes(ex: ExplodedResource, var_names: np.ndarray) -> ResourceTables:
    """Drop interactions with any subunit missing from `var_names` (filter_resource), then index the remaining
    subunits into the gene-subset column space (sorted intersection of resource entities and var_names,
    `_liana_pipe.py:161-164`).  Array operations only: `by_sample` calls this once per sample, because every sample
    has its own set of non-empty genes."""
    present = pd.Index(np.asarray(var_names, dtype=str)).get_indexer(ex.entities) >= 0 if len(ex.entities) else np.zeros(0, bool)
    present = np.append(present, True)                           # index -1 (padding) counts as present
    ok = present[ex.lig_all].all(axis=1) & present[ex.rec_all].all(axis=1)
    lig_all, rec_all = ex.lig_all[ok], ex.rec_all[ok]
    used = np.zeros(len(ex.entities) + 1, dtype=bool)
    used[lig_all.reshape(-1)] = True
    used[rec_all.reshape(-1)] = True
    used = used[:-1]                                             # the slot the -1 padding wrote to
    remap = np.append(np.cumsum(used, dtype=np.int64) - 1, -1).astype(np.int32)   # old entity -> kept entity; -1 stays -1
    lig_sub, rec_sub = remap[lig_all], remap[rec_all]
    # trailing all-padding columns go (widths = longest kept complex, as when built from the kept rows only)
    max_l = max(int((lig_sub >= 0).sum(axis=1).max(initial=1)), 1)
    max_r = max(int((rec_sub >= 0).sum(axis=1).max(initial=1)), 1)
    return ResourceTables(ex.ligand_complex[ok], ex.receptor_complex[ok], np.ascontiguousarray(lig_sub[:, :max_l]),
                          np.ascontiguousarray(rec_sub[:, :max_r]), ex.entities[used],
                          np.flatnonzero(used).astype(np.int32), np.flatnonzero(ok).astype(np.int32))


def build_resource_tables(resource: pd.DataFrame, var_names: np.ndarray, sep: str = "_"):
    """(ResourceTables for `var_names`) from a resource frame; see `explode_resource` / `tables_for_genes`."""
    return tables_for_genes(explode_resource(resource, sep), var_names)


def _resolve_side(sub: np.ndarray, means: np.ndarray, props: np.ndarray, first_subunit: bool):
    """For every (interaction, cluster): chosen subunit gene, its mean / prop, and the minimal prop
    over all subunits.  `means`, `props` are [L, G]."""
    n, m = sub.shape
    L = means.shape[0]
    pad = sub < 0
    safe = np.where(pad, 0, sub)
    mm = means[:, safe]                      # [L, n, m]
    pp = props[:, safe]
    mm = np.where(pad[None], np.inf, mm)
    pp = np.where(pad[None], np.inf, pp)
    if first_subunit:
        arg = np.zeros((L, n), dtype=np.int64)
    else:
        arg = np.argmin(mm, axis=2)          # first minimal subunit in '_' order
    ii = np.arange(n)[None, :]
    gene = safe[ii, arg]                     # [L, n]
    cl = np.arange(L)[:, None]
    return (gene.T.astype(np.int32), means[cl, gene].T, props[cl, gene].T, pp.min(axis=2).T)


@dataclass
class ResolvedTriples:
    inter: np.ndarray        # [T] interaction index
    src: np.ndarray          # [T] cluster index
    tgt: np.ndarray          # [T]
    lig_gene: np.ndarray     # [T] gene column of the resolved ligand subunit
    rec_gene: np.ndarray     # [T]
    ligand_means: np.ndarray     # [T] the complex-resolving statistic of the ligand subunit: cluster means
    receptor_means: np.ndarray   # [T] (float32) or, for CellChat, cluster trimeans (float64)
    ligand_props: np.ndarray     # [T] float64
    receptor_props: np.ndarray   # [T] float64
    keep: np.ndarray         # [T] bool (all True unless return_all_lrs)


def resolve_triples(tables: ResourceTables, means: np.ndarray, props: np.ndarray, expr_prop: float,
                    pair_mask: np.ndarray | None = None, return_all_lrs: bool = False, engine=None) -> ResolvedTriples:
    """means [L,G] = the statistic named by the method's `complex_cols` (float32 cluster means; float64 cluster
    trimeans for CellChat), props float64 [L,G], in the gene-subset column space.
    pair_mask [L,L] bool restricts (source, target) pairs (groupby_pairs).
    With an engine (the product path always passes one) the rows are built on the device
    (`lrperm_resolve_triples`); the NumPy formulation below is the checker of that kernel (CPU tests)."""
    if engine is not None:
        r = engine.resolve_triples(tables.lig_sub, tables.rec_sub, means, props, expr_prop, pair_mask,
                                   first_subunit=return_all_lrs, return_all=return_all_lrs)
        return ResolvedTriples(inter=r["inter"], src=r["src"], tgt=r["tgt"], lig_gene=r["lig_gene"], rec_gene=r["rec_gene"],
                               ligand_means=r["lig_stat"], receptor_means=r["rec_stat"], ligand_props=r["lig_props"],
                               receptor_props=r["rec_props"], keep=r["keep"])
    L = means.shape[0]
    # with return_all_lrs the reference de-duplicates subunit rows BEFORE the min-subunit step
    # (`_reassemble_complexes.py:52-57`), i.e. every complex is represented by its first subunits
    lg, lm, lp, lpmin = _resolve_side(tables.lig_sub, means, props, first_subunit=return_all_lrs)
    rg, rm, rp, rpmin = _resolve_side(tables.rec_sub, means, props, first_subunit=return_all_lrs)
    n = lg.shape[0]
    # min(all subunit props) >= expr_prop  <=>  both sides pass: two [n, L] masks instead of an [n, L, L] float cube;
    # the row order (per (source, target) pair with the source varying fastest, then interaction: the reference
    # concatenates per pair of `np.meshgrid(labels, labels).reshape(2, L*L).T`, `_liana_pipe.py:310-312`) is the
    # C order of a [target, source, interaction] mask, so one flatnonzero gives the rows already in place
    l_ok = np.ascontiguousarray((lpmin >= expr_prop).T)                   # [s, i]
    r_ok = np.ascontiguousarray((rpmin >= expr_prop).T)                   # [t, i]
    expressed = r_ok[:, None, :] & l_ok[None, :, :]                       # [t, s, i]
    mask_ts = None if pair_mask is None else np.ascontiguousarray(np.asarray(pair_mask, dtype=bool).T)   # [t, s]
    if return_all_lrs:
        sel = np.ones((L, L, n), dtype=bool) if mask_ts is None else np.broadcast_to(mask_ts[:, :, None], (L, L, n))
    else:
        sel = expressed if mask_ts is None else expressed & mask_ts[:, :, None]
    flat = np.flatnonzero(sel)
    keep = expressed.ravel()[flat] if return_all_lrs else np.ones(flat.shape[0], dtype=bool)
    st, i_idx = np.divmod(flat, n)
    t_idx, s_idx = np.divmod(st, L)
    li, ri = i_idx * L + s_idx, i_idx * L + t_idx                         # flat positions in the [n, L] side tables

    def side(a, pos, dtype):
        return np.take(np.ascontiguousarray(a).reshape(-1), pos).astype(dtype, copy=False)

    return ResolvedTriples(
        inter=i_idx.astype(np.int32), src=s_idx.astype(np.int32), tgt=t_idx.astype(np.int32),
        lig_gene=side(lg, li, np.int32), rec_gene=side(rg, ri, np.int32),
        ligand_means=side(lm, li, means.dtype), receptor_means=side(rm, ri, means.dtype),
        ligand_props=side(lp, li, np.float64), receptor_props=side(rp, ri, np.float64),
        keep=keep)
