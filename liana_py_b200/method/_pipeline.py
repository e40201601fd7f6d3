"""Host side of the drop-in: the equivalent of `liana_pipe` + `_run_method`
(`liana/method/sc/_liana_pipe.py:23-252, 363-452`) for the permutation methods, with the block
`:402-425` (cube -> gather -> score -> p-values) replaced by ONE call into the CUDA engine.

No CPU fallback: the per-cluster statistics and the permutation null always run on the GPU
through the C ABI (`liana_py_b200._engine`).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.stats import gmean as _scipy_gmean

from .. import _constants as K
from .. import _engine as E
from ..resource import handle_resource
from . import _pre, _tables
from ._dist import DistContext

_ENGINES = {}


def get_engine(device: int = 0) -> E.Engine:
    """One cached engine per CUDA device (raises EngineError if the library / GPU is missing)."""
    eng = _ENGINES.get(device)
    if eng is None or eng._h is None:
        eng = E.Engine(device)
        _ENGINES[device] = eng
    return eng


def resolve_devices(device, devices):
    """(`device` the input is prepared on, list of devices the null is split over or None)."""
    if devices is None:
        return int(device), None
    devices = [int(d) for d in devices]
    if not devices:
        raise ValueError("devices must name at least one CUDA device")
    return devices[0], (devices if len(devices) > 1 else None)


def observed_cpdb(ligand_means: np.ndarray, receptor_means: np.ndarray) -> np.ndarray:
    """float32 mean of the pair, zero where either side is zero (`_cellphonedb.py:25-27`)."""
    # `np.mean((a, b), axis=0)` of two float32 arrays is the float32 sum halved - written out, it skips the 2 x T copy
    # NumPy makes of the tuple (halving is exact, so the bits are the same)
    a, b = np.asarray(ligand_means), np.asarray(receptor_means)
    if a.dtype != np.float32 or b.dtype != np.float32:
        out = np.mean((a, b), axis=0)
    else:
        out = a + b
        out *= np.float32(0.5)
    out[(a == 0) | (b == 0)] = 0
    return out


def observed_gmean(ligand_means: np.ndarray, receptor_means: np.ndarray) -> np.ndarray:
    """scipy.stats.gmean of the pair in float32 (`_geometric_mean.py:22`); host side so that the
    threshold is bit-identical to the reference's NumPy math."""
    with np.errstate(divide="ignore"):
        return _scipy_gmean((ligand_means, receptor_means), axis=0)


def reference_perm_indices(n_cells: int, n_perms: int, seed: int) -> np.ndarray:
    """The reference's own draws: ONE default_rng(seed), `rng.permutation(arange(N))` P times in
    order (`_get_mean_perms.py:69-72,79`).  Used when perm_source='reference' (validation mode)."""
    chunks = [c for _, c in reference_perm_chunks(n_cells, n_perms, seed, 0, n_perms, max_bytes=1 << 62)]
    return chunks[0] if chunks else np.empty((0, n_cells), dtype=np.int32)


def reference_perm_chunks(n_cells: int, n_perms: int, seed: int, lo: int, hi: int, max_bytes: int = 256 << 20):
    """Draws [lo, hi) of the reference's permutation stream as (first draw, index array) chunks of at most `max_bytes`:
    the P x N index array of a large run never exists at once (500k cells x 1000 permutations would be 2 GB).  The
    stream is sequential (one PCG64 generator), so the draws before `lo` are generated and dropped."""
    rng = np.random.default_rng(seed=seed)
    idx = np.arange(n_cells)
    dtype = np.int32 if n_cells < 2 ** 31 else np.int64
    for _ in range(lo):
        rng.permutation(idx)
    per = max(1, int(max_bytes // max(n_cells * np.dtype(dtype).itemsize, 1)))
    for p0 in range(lo, hi, per):
        p1 = min(hi, p0 + per)
        out = np.empty((p1 - p0, n_cells), dtype=dtype)
        for p in range(p1 - p0):
            out[p] = rng.permutation(idx)
        yield p0, out


def default_perm_source() -> str:
    """Where the label shuffles come from when a call does not say: 'philox' (device, the production default) unless the
    environment variable LIANA_B200_PERM_SOURCE says 'reference' (the reference's own NumPy draws -> p-values identical
    to liana's for the same seed; validation runs)."""
    import os
    v = os.environ.get("LIANA_B200_PERM_SOURCE", "philox").strip().lower()
    if v not in ("philox", "reference"):
        raise ValueError("LIANA_B200_PERM_SOURCE must be 'philox' or 'reference'")
    return v


_TABLE_CACHE = {}
_EXPLODED_CACHE = {}


def exploded(res: pd.DataFrame):
    """The resource split into subunits (memoised on its content, see `_cached_tables`)."""
    hint = res.attrs.get("lrperm_resource_key")
    if hint is not None and hint[2] == len(res) and list(res.columns[:2]) == ["ligand", "receptor"]:
        rkey = hint                                   # a bundled table as `select_resource` handed it out
    else:
        # row hashes as bytes: sensitive to the ORDER of the interactions too (a plain sum would not be)
        rkey = hash(pd.util.hash_pandas_object(res[["ligand", "receptor"]], index=False).values.tobytes())
    ex = _EXPLODED_CACHE.get(rkey)
    if ex is None:
        if len(_EXPLODED_CACHE) >= 8:
            _EXPLODED_CACHE.pop(next(iter(_EXPLODED_CACHE)))
        ex = _EXPLODED_CACHE[rkey] = _tables.explode_resource(res)
    return rkey, ex


def _cached_tables(res: pd.DataFrame, var_names: np.ndarray):
    """(ResourceTables, all resource entities).  Two levels: the resource split into subunits (Python string work over
    4.6 k interactions, ~50 ms) is memoised on the resource's content; the tables for one gene set (array operations,
    < 1 ms) on (resource, gene names) - every sample of `by_sample` has its own set of non-empty genes."""
    rkey, ex = exploded(res)
    key = (rkey, len(var_names), hash(np.asarray(var_names, dtype=str).tobytes()))
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        if len(_TABLE_CACHE) >= 8:
            _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
        hit = _TABLE_CACHE[key] = (_tables.tables_for_genes(ex, var_names), ex.entities)
    return hit


class PipelineState:
    """Everything computed before any method-specific scoring (shared by rank_aggregate methods).

    means float32 [L,G], props float64 [L,G], sizes int32 [L] come from the engine's observed pass;
    the statistics only the non-permutation methods need are computed lazily from the engine's
    float64 per-(cluster, gene) moments (`lrperm_cluster_moments`)."""

    def __init__(self, prep, tables, engine, means, props, sizes, pair_mask, base=float(np.e), sizes_all=None):
        self.prep, self.tables, self.engine = prep, tables, engine
        self.means, self.props, self.sizes, self.pair_mask = means, props, sizes, pair_mask
        # cells without a label (NaN in `groupby`) stay in the matrix like in the reference: they form one more,
        # nameless cluster on the engine (last row of every [cluster, gene] table) that no interaction refers to, but
        # that takes part in the shuffles, in the whole-matrix statistics and in the "rest" of the log fold changes
        self.sizes_all = sizes if sizes_all is None else sizes_all
        self.engine_tables = None
        self.engine_labels = None            # (labels int32 [N], n_clusters) as set on the engine (`prepare`)
        self.base = float(base)
        self._moments = None
        self._cache = {}
        self._replicas = {}

    def scale_matrix_like_reference(self, mat_max):
        """Reference quirk, reproduced only on request (`reference_quirks=True` of a consensus): CellChat's null divides
        the SHARED matrix in place, `adata.X /= norm_factor` (`_get_mean_perms.py:43-44`), so every permutation method
        that runs after CellChat in the same consensus shuffles the scaled matrix (while its observed scores, computed
        before, stay unscaled).  SciPy evaluates the division as a float32 multiply by the reciprocal; so does this."""
        recip32, _ = E.Engine.trimean_reciprocals(mat_max)
        eng = self.engine
        n, g = eng.N, eng.G
        ip, ix, dt = eng.export_matrix_csr()
        eng.set_matrix_csr(ip, ix, (np.asarray(dt, dtype=np.float32) * np.float32(recip32)).astype(np.float32), n, g)
        eng.set_labels(*self.engine_labels)
        self._replicas = {}
        self.matrix_scaled_by = np.float32(recip32)

    def replicas(self, devices) -> list:
        """Engines of `devices` (CUDA indices) that hold this state's matrix and labels; the first entry is the state's
        own engine.  The committed matrix is exported once on its device and copied to the others device to device
        (peer copies over NVLink - nothing crosses PCIe again), each device then builds its own tables; all of that
        runs in one thread per device."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        devices = [int(d) for d in devices]
        if devices[0] != self.engine.device:
            raise ValueError("devices[0] must be the device the input was prepared on")
        if len(set(devices)) != len(devices):
            raise ValueError("devices must be distinct")
        missing = [d for d in devices[1:] if d not in self._replicas]
        if missing:
            n, g, _ = self.engine.matrix_info()
            src = self.engine.export_matrix_csr(torch.device("cuda", devices[0]))
            labels, L = self.engine_labels

            def make(d):
                dev = torch.device("cuda", d)
                with torch.cuda.device(dev):
                    ip, ix, dt = (t.to(dev, non_blocking=False) for t in src)
                    eng = get_engine(d)
                    eng.set_matrix_csr(ip, ix, dt, n, g)
                    eng.set_labels(labels, L)
                return d, eng

            with ThreadPoolExecutor(len(missing)) as pool:
                for d, eng in pool.map(make, missing):
                    self._replicas[d] = eng
        return [self.engine] + [self._replicas[d] for d in devices[1:]]

    def resolved(self, expr_prop, return_all_lrs, stat="means"):
        """(ResolvedTriples, base lr_res frame) for one filter setting; the methods of a consensus that share
        `complex_cols` and `expr_prop` share the tables (the reference re-runs `filter_reassemble_complexes` per
        method, `_liana_pipe.py:378-382`).  `stat` names the statistic the complexes are resolved by: 'means'
        (ligand_means / receptor_means) or 'trimean' (CellChat's complex_cols).  Callers must not mutate the frame."""
        key = ("tri", float(expr_prop), bool(return_all_lrs), stat)
        if key not in self._cache:
            by, props, mask = (self.means if stat == "means" else self.trimean()), self.props, self.pair_mask
            if self.engine_tables is not None:
                # unlabeled cells: the tables carry the engine's extra cluster, masked out of every (source, target) pair
                L = len(self.sizes)
                by = self.engine_tables[0] if stat == "means" else self.trimean(all_rows=True)
                props = self.engine_tables[1]
                mask = np.zeros((L + 1, L + 1), dtype=bool)
                mask[:L, :L] = True if self.pair_mask is None else self.pair_mask
            tri = _tables.resolve_triples(self.tables, by, props, expr_prop, mask, return_all_lrs, engine=self.engine)
            self._cache[key] = (tri, base_frame(self, tri, stat))
        return self._cache[key]

    def gene_cdf(self) -> np.ndarray:
        """scSeqComm's `_gene_cdf` per (cluster, gene), float64 [L,G] (`_liana_pipe.py:479-483`): P(N(cluster mean,
        cluster std / sqrt(n_c)) <= gene mean), 0 where the gene's cluster mean is 0."""
        if "cdf" not in self._cache:
            from scipy.stats import norm
            if self.prep.cluster_stats is None:
                raise RuntimeError("the input was prepared without cluster statistics (internal error)")
            mean, std = self.prep.cluster_stats
            scale = std / np.sqrt(self.sizes.astype(np.float64))
            with np.errstate(divide="ignore", invalid="ignore"):
                p = norm.cdf(self.means, loc=mean[:, None], scale=scale[:, None])
            p[self.means == 0] = 0
            self._cache["cdf"] = p
        return self._cache["cdf"]

    def trimean(self, all_rows: bool = False) -> np.ndarray:
        """Observed cluster trimeans of X / mat_max, float64 [L,G] (`_liana_pipe.py:307-308, 462-465`); `all_rows`
        keeps the engine's extra row for unlabeled cells."""
        if "trimean" not in self._cache:
            if self.prep.mat_max is None:
                raise RuntimeError("the input was prepared without mat_max (internal error)")
            self._cache["trimean"] = self.engine.observed_trimean(self.prep.mat_max)
        return self._cache["trimean"] if all_rows else self._cache["trimean"][:len(self.sizes)]

    def _get_moments(self):
        if self._moments is None:
            self._moments = self.engine.cluster_moments(self.base)
        return self._moments

    def zscores(self) -> np.ndarray:
        """Cluster means of the per-gene z-scored matrix, float32 [L,G] (`_liana_pipe.py:259-262, 303-304`:
        sc.pp.scale = (x - mean_g) / std_g with the unbiased variance, std 0 -> 1; the mean over a
        cluster's cells is (cluster mean - mean_g) / std_g)."""
        if "z" not in self._cache:
            sum_x, sum_sq, _ = self._get_moments()
            n = float(self.sizes_all.sum())
            mean = sum_x.sum(axis=0) / n
            var = (sum_sq.sum(axis=0) / n - mean ** 2) * (n / (n - 1.0))
            std = np.sqrt(np.maximum(var, 0.0))
            std[std == 0] = 1.0
            L = len(self.sizes)
            cl_mean = sum_x[:L] / self.sizes[:, None].astype(np.float64)
            self._cache["z"] = ((cl_mean - mean[None, :]) / std[None, :]).astype(np.float32)
        return self._cache["z"]

    def logfc(self) -> np.ndarray:
        """1-vs-rest log2FC of base^x - 1 ("normcounts"), float32 [L,G] (`_liana_pipe.py:341-360`)."""
        if "fc" not in self._cache:
            _, _, sum_e = self._get_moments()
            n_c = self.sizes.astype(np.float64)[:, None]
            n = float(self.sizes_all.sum())
            total = sum_e.sum(axis=0)[None, :]                  # unlabeled cells belong to every cluster's "rest"
            sum_e = sum_e[:len(self.sizes)]
            subj = sum_e / n_c
            with np.errstate(invalid="ignore", divide="ignore"):
                rest = (total - sum_e) / (n - n_c)
            self._cache["fc"] = (np.log2(subj + 1.0) - np.log2(rest + 1.0)).astype(np.float32)
        return self._cache["fc"]

    def means_sums(self, side: str) -> np.ndarray:
        """NATMI denominators (`_liana_pipe.py:185-190, 337-338`): for a ligand row (target t, gene g) the sum
        of the gene's cluster means over all sources paired with t; for a receptor row (source s, gene g) the
        sum over all targets paired with s.  float32 [L,G], accumulated in cluster order with the
        compensated float32 summation pandas' groupby-sum uses."""
        key = "sum_" + side
        if key not in self._cache:
            L = self.means.shape[0]
            mask = np.ones((L, L), dtype=bool) if self.pair_mask is None else self.pair_mask   # [source, target]
            if side == "receptor":
                mask = mask.T                                                                   # [target, source]
            out = np.zeros_like(self.means)
            comp = np.zeros_like(self.means)
            for j in range(L):                 # j = the cluster being added; rows of `out` = the fixed side
                sel = mask[j]                  # ligand: mask[source j, target t]; receptor: mask.T[target j, source s]
                y = self.means[j][None, :] - comp
                t = out + y
                new_comp = (t - out) - y
                out = np.where(sel[:, None], t, out)
                comp = np.where(sel[:, None], new_comp, comp)
            self._cache[key] = out
        return self._cache[key]


def prepare(adata, groupby, resource_name, resource, interactions, groupby_pairs, min_cells, use_raw, layer,
            verbose, device=0, base=float(np.e), want_max=False, want_cluster_stats=False,
            mdata_kwargs=None, dist: DistContext | None = None) -> PipelineState:
    res = handle_resource(interactions=interactions, resource=resource, resource_name=resource_name, verbose=verbose)
    if _pre.is_mudata(adata):                    # `_liana_pipe.py:126-129`
        adata = _pre.mdata_to_anndata(adata, **(mdata_kwargs or {}), verbose=verbose)
        use_raw, layer = False, None
    groupby_subset = None
    if groupby_pairs is not None:
        if not (K.SOURCE in groupby_pairs.columns) | (K.TARGET in groupby_pairs.columns):
            raise AssertionError(f"{K.SOURCE} and {K.TARGET} must be in groupby_pairs")
        groupby_subset = np.union1d(groupby_pairs[K.SOURCE].unique(), groupby_pairs[K.TARGET].unique())
    eng = get_engine(device)
    prep = _pre.prepare_input(adata, groupby=groupby, min_cells=min_cells, engine=eng, groupby_subset=groupby_subset,
                              use_raw=use_raw, layer=layer, verbose=verbose, want_max=want_max,
                              want_cluster_stats=want_cluster_stats, dist=dist)
    tables, all_entities = _cached_tables(res, prep.var_names)
    _pre.assert_covered(all_entities, prep.var_names, verbose=verbose)
    # subset to the resource's genes (sorted intersection, `_liana_pipe.py:161-164`): a column map for the
    # staged matrix instead of SciPy fancy indexing + index sorting
    pos = np.searchsorted(prep.var_names, tables.entities)
    col_map = np.full(prep.n_genes_in, -1, dtype=np.int32)
    col_map[prep.var_cols[pos]] = np.arange(len(tables.entities), dtype=np.int32)
    if verbose:
        print(f"Generating ligand-receptor stats for {prep.n_cells} samples and {len(tables.entities)} features")
    L = len(prep.categories)
    if prep.n_cells_local is None:
        eng.commit_matrix(col_map, len(tables.entities), prep.n_cells)
    else:
        # every rank committed the cells of its own row block: the committed blocks (resource genes only, ~6 % of the
        # staged bytes) are all-gathered over NVLink and become every rank's matrix
        eng.commit_matrix(col_map, len(tables.entities), prep.n_cells_local)
        d_ip, d_ix, d_dt = dist.all_gather_csr(eng)
        eng.set_matrix_csr(d_ip, d_ix, d_dt, prep.n_cells, len(tables.entities))
    prep.var_names = tables.entities
    labels, L_eng = prep.labels, L
    if len(labels) and labels.min() < 0:          # cells without a label: one more cluster on the engine (see PipelineState)
        labels, L_eng = np.where(labels < 0, L, labels).astype(np.int32), L + 1
    eng.set_labels(labels, L_eng)
    engine_labels = (np.ascontiguousarray(labels, dtype=np.int32), L_eng)
    means_all, stored, sizes_all = eng.observed()
    props_all = stored / sizes_all[:, None].astype(np.float64)
    means, props, sizes = means_all[:L], props_all[:L], sizes_all[:L]
    pair_mask = None
    if groupby_pairs is not None:
        pos = {c: i for i, c in enumerate(prep.categories)}
        pair_mask = np.zeros((L, L), dtype=bool)
        for s, t in zip(groupby_pairs[K.SOURCE], groupby_pairs[K.TARGET]):
            if s in pos and t in pos:
                pair_mask[pos[s], pos[t]] = True
    state = PipelineState(prep, tables, eng, means, props, sizes, pair_mask, base=base, sizes_all=sizes_all)
    state.engine_labels = engine_labels
    if L_eng != L:
        state.engine_tables = (means_all, props_all)       # [L + 1, G]: what the engine's row resolution indexes
    return state


_CHUNK_POOL = None


def _chunk_pool():
    """Second small pool for the row chunks of `_take_names` (which itself runs inside `_frame_pool` tasks)."""
    global _CHUNK_POOL
    if _CHUNK_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _CHUNK_POOL = ThreadPoolExecutor(max_workers=max(1, min(16, os.cpu_count() or 1)), thread_name_prefix="lrperm-take")
    return _CHUNK_POOL


def _take_names(names, codes):
    """names[codes] as a pandas string array: one vectorised take on the Index's backing array (10x faster
    than fancy-indexing a NumPy unicode array and letting pandas re-infer every string).  Millions of rows (the frame
    of `by_sample`) are taken in row chunks on a thread pool - Arrow's take releases the GIL - and stay chunked: a
    pandas Arrow string column is a ChunkedArray anyway (1 s -> 0.2 s per column at 28 M rows on 8 threads)."""
    idx = names if isinstance(names, pd.Index) else pd.Index(np.asarray(names))
    arr = idx.array
    codes = np.asarray(codes)
    if len(codes) >= (1 << 21) and isinstance(arr, pd.arrays.ArrowStringArray) and hasattr(arr, "_pa_array"):
        try:
            import pyarrow as pa
            import pyarrow.compute as pc
            table = arr._pa_array
            table = table.combine_chunks() if isinstance(table, pa.ChunkedArray) else table
            parts = np.array_split(codes, max(1, min(64, len(codes) >> 20)))
            outs = list(_chunk_pool().map(lambda c: pc.take(table, pa.array(c)), parts))
            return type(arr)(pa.chunked_array(outs), dtype=arr.dtype)
        except Exception:        # an Arrow / pandas build without these entry points: the plain take below
            pass
    return arr.take(codes.astype(np.int64, copy=False))


def base_frame(state: PipelineState, tri: _tables.ResolvedTriples, stat: str = "means") -> pd.DataFrame:
    """lr_res with the reference's column names / dtypes / order for the kept rows.  The six name columns hold
    int32 CODES until `materialise_names` runs on the final (filtered, sorted) frame: selecting, sorting and
    concatenating integer columns is several times cheaper than moving strings around."""
    lig_stat, rec_stat = (K.LIGAND_MEANS, K.RECEPTOR_MEANS) if stat == "means" else (K.LIGAND_TRIMEAN, K.RECEPTOR_TRIMEAN)
    return pd.DataFrame({
        K.LIGAND: tri.lig_gene, K.LIGAND_COMPLEX: tri.inter,
        lig_stat: tri.ligand_means, K.LIGAND_PROPS: tri.ligand_props,
        K.RECEPTOR: tri.rec_gene, K.RECEPTOR_COMPLEX: tri.inter,
        rec_stat: tri.receptor_means, K.RECEPTOR_PROPS: tri.receptor_props,
        K.SOURCE: tri.src, K.TARGET: tri.tgt,
    }, copy=False)              # views of the resolved arrays (never written: run_method works on a copy)


def name_tables(state: PipelineState) -> dict:
    """The small lookup tables `materialise_names` needs (what travels with a code frame between ranks)."""
    names, cats, tb = state.prep.var_names, state.prep.categories, state.tables
    return {K.LIGAND: names, K.RECEPTOR: names, K.LIGAND_COMPLEX: tb.ligand_complex,
            K.RECEPTOR_COMPLEX: tb.receptor_complex, K.SOURCE: cats, K.TARGET: cats}


def global_name_tables(res: pd.DataFrame, categories) -> dict:
    """Name tables that do not depend on the sample / rank: every gene the RESOURCE names, the ligand / receptor
    complex of every interaction of the resource, every cluster of the whole object.  `by_sample` exchanges frames
    whose name columns hold codes into these tables (integers travel over NCCL; the strings are built once, on the
    consumer)."""
    _, ex = exploded(res)
    genes = pd.Index(ex.entities)
    return {K.LIGAND: genes, K.RECEPTOR: genes, K.LIGAND_COMPLEX: pd.Index(ex.ligand_complex),
            K.RECEPTOR_COMPLEX: pd.Index(ex.receptor_complex), K.SOURCE: pd.Index(categories),
            K.TARGET: pd.Index(categories)}


def global_code_maps(state: PipelineState, global_tables: dict) -> dict:
    """local code -> global code per name column: genes and interactions through the indices the resource tables
    carry (`ResourceTables.entity_index / inter_index`, no string is touched), clusters through one small lookup."""
    tb = state.tables
    cats = global_tables[K.SOURCE].get_indexer(pd.Index(np.asarray(state.prep.categories))).astype(np.int32)
    if (cats < 0).any():
        raise RuntimeError("a cluster of the sample is missing from the object's categories (internal error)")
    maps = {K.LIGAND: tb.entity_index, K.RECEPTOR: tb.entity_index, K.LIGAND_COMPLEX: tb.inter_index,
            K.RECEPTOR_COMPLEX: tb.inter_index, K.SOURCE: cats, K.TARGET: cats}
    # global codes in the narrowest unsigned type their table allows (1,877 genes, 4,624 interactions: 16 bits; clusters:
    # 8): 7 code columns are 11 instead of 28 of the 64 bytes a `by_sample` row carries between ranks
    return {c: np.asarray(m).astype(code_dtype(len(global_tables[c])), copy=False) for c, m in maps.items()}


def code_dtype(n_names: int):
    """Narrowest unsigned integer type that indexes a table of `n_names` entries (int32 beyond 16 bits)."""
    return np.uint8 if n_names <= 0x100 else np.uint16 if n_names <= 0x10000 else np.int32


def to_global_codes(columns: dict, maps: dict) -> dict:
    """Columns of a code frame (1-D arrays), the name columns re-coded into `global_name_tables` (a take per column from
    a table of a few thousand entries; NumPy releases the GIL for it, so the columns go through the thread pool)."""
    names = list(columns)

    def recode(c):
        v = columns[c]
        return maps[c].take(v) if c in maps else np.ascontiguousarray(v)

    if len(names) < 2 or len(columns[names[0]]) < 65536:
        return {c: recode(c) for c in names}
    return dict(zip(names, _frame_pool().map(recode, names)))


def sorted_code_columns(engine, frame: pd.DataFrame, by: str, ascending: bool) -> dict:
    """`sort_frame` for a frame that stays numeric: every column gathered once as a 1-D array (no pandas take over
    consolidated blocks)."""
    vals = np.asarray(frame[by].values, dtype=np.float64)
    if engine is None or len(vals) < 4096 or np.isnan(vals).any():
        frame = frame.sort_values(by=by, ascending=ascending)
        return {c: frame[c].values for c in frame.columns}
    order = engine.argsort(vals, descending=not ascending)
    cols = list(frame.columns)
    arrays = [frame[c].values for c in cols]
    return dict(zip(cols, _frame_pool().map(lambda a: a.take(order), arrays)))


def frame_from_codes(columns: dict, global_tables: dict) -> pd.DataFrame:
    """The result frame from gathered columns: code columns -> names (one vectorised take each, in the thread pool)."""
    cols = list(columns)

    def build(c):
        v = columns[c]
        return _take_names(global_tables[c], v) if c in global_tables else v

    done = list(_frame_pool().map(build, cols))
    return pd.DataFrame(dict(zip(cols, done)), copy=False)


def materialise_names(state_or_tables, frame: pd.DataFrame) -> pd.DataFrame:
    """Replace the code columns of `base_frame` by the gene / complex / cluster names (in place, same order)."""
    lookup = state_or_tables if isinstance(state_or_tables, dict) else name_tables(state_or_tables)
    for col, table in lookup.items():
        if col in frame.columns:
            frame[col] = _take_names(table, frame[col].values)
    return frame


def sort_frame(engine, frame: pd.DataFrame, by: str, ascending: bool) -> pd.DataFrame:
    """`liana_res.sort_values(by, ascending)` (`_liana_pipe.py:246-250`) with the argsort on the device; the order among
    equal values is the original row order (the reference's quicksort leaves it unspecified).  NaN keys go through
    pandas (NaN last)."""
    vals = np.asarray(frame[by].values, dtype=np.float64)
    if engine is None or len(vals) < 4096 or np.isnan(vals).any():
        return frame.sort_values(by=by, ascending=ascending)
    return frame.take(engine.argsort(vals, descending=not ascending))


_FRAME_POOL = None


def _frame_pool():
    global _FRAME_POOL
    if _FRAME_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _FRAME_POOL = ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1)), thread_name_prefix="lrperm-frame")
    return _FRAME_POOL


def finish_frame(state: PipelineState, frame: pd.DataFrame, by: str, ascending: bool) -> pd.DataFrame:
    """`sort_frame` + `materialise_names` in one pass over the columns: the row order comes from the device, every
    column is gathered once as a 1-D array (NumPy take / Arrow string take release the GIL, so the columns go through
    a small thread pool) and the frame is assembled from the finished columns - instead of a pandas `take` over the
    consolidated blocks followed by six string takes (70 -> 15 ms at 785 k rows)."""
    vals = np.asarray(frame[by].values, dtype=np.float64)
    if state.engine is None or len(vals) < 4096 or np.isnan(vals).any():
        return materialise_names(state, sort_frame(state.engine, frame, by, ascending))
    order = state.engine.argsort(vals, descending=not ascending)
    return pd.DataFrame(gather_columns(state, frame, order), index=pd.Index(order), copy=False)


def gather_columns(state: PipelineState, frame: pd.DataFrame, order: np.ndarray) -> dict:
    """Every column of a code frame gathered in `order` as a finished 1-D array (name columns materialised), through the
    thread pool - the column loop of `finish_frame`, usable while the engine is busy (it touches no device)."""
    lookup = name_tables(state)
    cols = list(frame.columns)
    arrays = [frame[c].values for c in cols]

    def gather(i):
        v = arrays[i].take(order)
        return _take_names(lookup[cols[i]], v) if cols[i] in lookup else v

    return dict(zip(cols, _frame_pool().map(gather, range(len(cols)))))


def add_method_columns(state: PipelineState, tri: _tables.ResolvedTriples, frame: pd.DataFrame, add_cols) -> pd.DataFrame:
    """Method-specific statistic columns of the rows in `tri` (the resolved min-subunit gene of every row
    carries its own statistics through `filter_reassemble_complexes`, `_reassemble_complexes.py:60-66`)."""
    frame = frame.copy(deep=False)           # new columns only; the cached base frame is not written
    for col in add_cols:
        if col == "ligand_zscores":
            frame[col] = state.zscores()[tri.src, tri.lig_gene]
        elif col == "receptor_zscores":
            frame[col] = state.zscores()[tri.tgt, tri.rec_gene]
        elif col == "ligand_logfc":
            frame[col] = state.logfc()[tri.src, tri.lig_gene]
        elif col == "receptor_logfc":
            frame[col] = state.logfc()[tri.tgt, tri.rec_gene]
        elif col == "ligand_means_sums":
            frame[col] = state.means_sums("ligand")[tri.tgt, tri.lig_gene]
        elif col == "receptor_means_sums":
            frame[col] = state.means_sums("receptor")[tri.src, tri.rec_gene]
        elif col == "mat_mean":
            frame[col] = np.float32(state.prep.mat_mean)
        elif col == K.MAT_MAX:
            frame[col] = np.float32(state.prep.mat_max)
        elif col == "ligand_cdf":
            frame[col] = state.gene_cdf()[tri.src, tri.lig_gene]
        elif col == "receptor_cdf":
            frame[col] = state.gene_cdf()[tri.tgt, tri.rec_gene]
        elif col in frame.columns:
            pass                                   # a key / complex / proportion column the frame already carries
        elif col == K.LIGAND_MEANS:                # only reachable as supplementary columns of CellChat
            frame[col] = state.means[tri.src, tri.lig_gene]
        elif col == K.RECEPTOR_MEANS:
            frame[col] = state.means[tri.tgt, tri.rec_gene]
        elif col in (K.LIGAND_TRIMEAN, K.RECEPTOR_TRIMEAN):
            if state.prep.mat_max is None:
                # `_get_lr` only computes trimeans when `mat_max` is among the requested columns
                # (`_liana_pipe.py:143-146, 307-308`); the reference then fails selecting the column
                raise KeyError(f"`{col}` needs `mat_max` among the supp_columns")
            frame[col] = state.trimean()[tri.src, tri.lig_gene] if col == K.LIGAND_TRIMEAN \
                else state.trimean()[tri.tgt, tri.rec_gene]
        elif col in ("ligand_pvals", "receptor_pvals"):
            raise NotImplementedError(f"`{col}` comes from scanpy's rank_genes_groups (`_liana_pipe.py:279-283`), "
                                      "which is outside the scope of the permutation engine")
        else:
            raise KeyError(f"`{col}` is not a column of the ligand-receptor statistics")
    return frame


def method_columns(method_add_cols, supp_columns) -> list:
    """`_add_cols + supp_columns` of `liana_pipe` (`_liana_pipe.py:111-114`): the method's own statistic columns
    followed by the caller's supplementary ones, once each (the reference takes `np.union1d` of them)."""
    out = list(method_add_cols)
    for col in (supp_columns or []):
        if col not in out:
            out.append(col)
    return out


def _shard_runner(state, dist, devices, n_perms, setup, run_range):
    """Shared plumbing of the two null entry points.  One engine: permutations [lo, hi) of this rank.  `devices`
    (several CUDA indices, one process): the permutation range is split over the devices' engines, one thread each
    (the C ABI calls release the GIL), and the integer counts are added on the host.  Ranks (`dist`): the uint32
    count vectors are summed with one all-reduce - identical results for any number of GPUs either way."""
    if devices is not None and len(devices) > 1:
        if dist is not None and dist.world_size > 1:
            raise ValueError("pass either `dist` (one process per GPU) or `devices` (one process, several GPUs)")
        from concurrent.futures import ThreadPoolExecutor
        engines = state.replicas(devices)
        n = len(engines)

        def work(k):
            import torch
            with torch.cuda.device(engines[k].device):
                setup(engines[k])
                return run_range(engines[k], (n_perms * k) // n, (n_perms * (k + 1)) // n)

        with ThreadPoolExecutor(n) as pool:
            parts = list(pool.map(work, range(n)))
        return {k: sum((p[k].astype(np.uint32) for p in parts[1:]), parts[0][k].astype(np.uint32)) for k in parts[0]}
    rank, world = (dist.rank, dist.world_size) if dist is not None else (0, 1)
    setup(state.engine)
    out = run_range(state.engine, (n_perms * rank) // world, (n_perms * (rank + 1)) // world)
    if dist is not None and world > 1:
        out = {k: dist.all_reduce_sum_u32(v) for k, v in out.items()}
    return out


def permutation_counts(state: PipelineState, tri, obs_cpdb, obs_gmean, n_perms, seed, perm_source, order,
                       gmean_mode, dist: DistContext | None = None, devices=None):
    """Exceedance counts for the kept rows.  Permutations are sharded by global id over ranks (`dist`) or over the
    devices of one process (`devices`); only the uint32 count vectors are combined (integer -> identical for any
    GPU count)."""
    scores = (E.SCORE_CPDB if obs_cpdb is not None else 0) | (E.SCORE_GMEAN if obs_gmean is not None else 0)
    perm_source = default_perm_source() if perm_source is None else perm_source
    if perm_source not in ("philox", "reference"):
        raise ValueError("perm_source must be 'philox' or 'reference'")
    if order is None:
        order = "exact" if perm_source == "reference" else "fast"
    if order not in ("exact", "fast"):
        raise ValueError("order must be 'exact', 'fast' or None")
    kw = dict(order=E.ORDER_EXACT if order == "exact" else E.ORDER_FAST, scores=scores,
              gmean_mode=E.GMEAN_LITERAL if gmean_mode == "literal" else E.GMEAN_PRODUCT)

    def setup(eng):
        eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt, obs_cpdb, obs_gmean)

    def run_range(eng, lo, hi):
        if perm_source == "philox":
            return eng.run(hi - lo, seed=seed, perm_offset=lo, **kw)
        # the reference's draws, a bounded chunk of the index array at a time; integer counts add up
        out = None
        for _, idx in reference_perm_chunks(state.prep.n_cells, n_perms, seed, lo, hi):
            part = eng.run(len(idx), seed=seed, perm_idx=idx, **kw)
            out = part if out is None else {k: out[k] + part[k] for k in out}
        return out if out is not None else eng.run(0, seed=seed, **kw)

    return _shard_runner(state, dist, devices, n_perms, setup, run_range)


KH = 0.5          # `cellchat._kh` (`liana/method/sc/_cellchat.py:53`)


def observed_cellchat(ligand_trimean: np.ndarray, receptor_trimean: np.ndarray, kh: float = KH) -> np.ndarray:
    """`_lr_probability` of the observed trimeans, float64 (`_cellchat.py:7-10, 29`); host side so that the
    threshold is bit-identical to the reference's NumPy math."""
    lr_prob = np.prod((ligand_trimean, receptor_trimean), axis=0)
    return lr_prob / (kh + lr_prob)


def permutation_counts_trimean(state: PipelineState, tri, obs_prob, n_perms, seed, perm_source, kh=KH,
                               dist: DistContext | None = None, devices=None):
    """CellChat exceedance counts for the kept rows (selection, not summation: no `order` applies).  Sharded over
    ranks / devices like `permutation_counts`."""
    perm_source = default_perm_source() if perm_source is None else perm_source
    if perm_source not in ("philox", "reference"):
        raise ValueError("perm_source must be 'philox' or 'reference'")
    mat_max = state.prep.mat_max

    def setup(eng):
        eng.set_triples(tri.lig_gene, tri.rec_gene, tri.src, tri.tgt)

    def run_range(eng, lo, hi):
        if perm_source == "philox":
            return eng.run_trimean(hi - lo, mat_max, obs_prob=obs_prob, kh=kh, seed=seed, perm_offset=lo)
        out = None
        for _, idx in reference_perm_chunks(state.prep.n_cells, n_perms, seed, lo, hi):
            part = eng.run_trimean(len(idx), mat_max, obs_prob=obs_prob, kh=kh, seed=seed, perm_idx=idx)
            out = part if out is None else {k: out[k] + part[k] for k in out}
        return out if out is not None else eng.run_trimean(0, mat_max, obs_prob=obs_prob, kh=kh, seed=seed)

    return _shard_runner(state, dist, devices, n_perms, setup, run_range)
