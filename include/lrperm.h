/*
 * lrperm.h - C ABI of the B200 permutation-null engine for ligand-receptor scoring.
 *
 * Drop-in boundary for the hot path of saezlab/liana-py (reference paths relative to
 * /root/reference):
 *
 *   liana/method/sc/_liana_pipe.py:402-425        _run_method: cube -> gather -> score
 *   liana/method/_pipe_utils/_get_mean_perms.py   _get_means_perms / _generate_perms_cube /
 *                                                 _permute_and_aggregate / _calculate_pvals
 *   liana/method/sc/_cellphonedb.py:9-30          _cpdb_score   (null side)
 *   liana/method/sc/_geometric_mean.py:6-25       _gmean_score  (null side)
 *   liana/method/sc/_liana_pipe.py:285-302        per-cluster observed means / props
 *   liana/method/sc/_cellchat.py:7-32, _liana_pipe.py:307-308, 394-397, 462-465
 *                                                 CellChat: trimean null + _lr_probability
 *
 * The reference is pure Python, so "the FFI a maintainer would bind" is ctypes; the stub is
 * in INTEGRATION.md.  All entry points take plain pointers and sizes.  Every function
 * returns an int status (LRPERM_OK == 0); nothing aborts or throws across the boundary.
 * `lrperm_last_error` returns a human-readable message for the last non-zero status.
 *
 * Pointers flagged `*_on_device != 0` are CUDA device pointers in the engine's device
 * (e.g. `torch.Tensor.data_ptr()`); otherwise they are host pointers and the engine does
 * the copies on its stream.  One engine = one device + one stream; not thread-safe.
 */
#ifndef LRPERM_H
#define LRPERM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRPERM_ABI_VERSION 5   /* 2: trimean (CellChat), staged max, consensus / resolve / argsort entry points, async pinned uploads; 3: dense staging; 4: run statistics, matrix export; 5: staged info */

/* status codes */
#define LRPERM_OK              0
#define LRPERM_ERR_INVALID     1   /* bad argument / call order */
#define LRPERM_ERR_CUDA        2   /* CUDA runtime error        */
#define LRPERM_ERR_OOM         3   /* device allocation failed  */
#define LRPERM_ERR_UNSUPPORTED 4   /* shape outside engine limits */

/* where the permutations come from (replaces `rng.permutation(idx)`, _get_mean_perms.py:79) */
#define LRPERM_PERMS_INDEX   0   /* caller supplies the reference's own index array perm_idx[P][N] */
#define LRPERM_PERMS_PHILOX  1   /* engine generates them on device from (seed, global perm id)    */

/* summation order of the per-(perm, cluster, gene) means (replaces _permute_and_aggregate :61-64) */
#define LRPERM_ORDER_EXACT   0   /* the reference's order: float32, sequential in position order -> bit-identical cube */
#define LRPERM_ORDER_FAST    1   /* float32, sequential in cell order per gene (<= 1e-5 rel. from the reference)      */

/* which null scores to count (bitmask) */
#define LRPERM_SCORE_CPDB    1   /* (a + b) / 2 >= lr_means        (_cellphonedb.py:25-28)    */
#define LRPERM_SCORE_GMEAN   2   /* gmean(a, b) >= lr_gmeans       (_geometric_mean.py:22-23) */

/* how the geometric-mean comparison is evaluated */
#define LRPERM_GMEAN_PRODUCT 0   /* a*b >= g*g in float64 (exact for float32 inputs)          */
#define LRPERM_GMEAN_LITERAL 1   /* exp((log a + log b)/2) >= g in float64, as scipy.stats.gmean */

typedef struct lrperm_engine lrperm_engine;

int         lrperm_abi_version(void);
const char *lrperm_last_error(const lrperm_engine *h);   /* h may be NULL: error of the last failed create */

/* Create / destroy an engine bound to CUDA device `device`. */
int  lrperm_create(int device, lrperm_engine **out);
void lrperm_destroy(lrperm_engine *h);

/* Use an existing CUDA stream (cudaStream_t passed as void*; NULL = the engine's own stream). */
int lrperm_set_stream(lrperm_engine *h, void *cuda_stream);

/*
 * The expression matrix the hot path sees: CSR float32, N cells x G genes, already subset to
 * the resource's genes (input contract: _pipe_utils/_pre.py:65-169, _liana_pipe.py:161-164).
 * indptr has N+1 entries of int32 (indptr_is_64 == 0) or int64; nnz must be < 2^31, G <= 65535.
 *
 * Host buffers in PINNED memory (cudaHostAlloc / cudaHostRegister) are uploaded asynchronously, cell tile by cell tile,
 * on a copy stream of the engine: the call returns at once, and the buffers must stay valid and unchanged until the
 * first call that consumes the matrix (lrperm_run, lrperm_observed, ...) has returned - as with cudaMemcpyAsync.
 * lrperm_run in FAST order then starts its first batch on the first tile while the others are still in flight.
 * The row pointers are validated before the call returns; a column out of range is reported by the consuming call.
 * Pageable host buffers and device buffers are consumed before the call returns.
 */
int lrperm_set_matrix_csr(lrperm_engine *h, int64_t n_cells, int64_t n_genes,
                          const void *indptr, int indptr_is_64,
                          const int32_t *indices, const float *data, int on_device);

/*
 * Input contract on device - replaces the matrix work of prep_check_adata
 * (liana/method/_pipe_utils/_pre.py:112-142: float32 CSR, empty-feature / empty-sample / non-finite
 * checks, the min_cells row filter :153-162, the alphabetical gene order :168) and the gene subset of
 * liana_pipe (liana/method/sc/_liana_pipe.py:161-164), which the reference does with SciPy copies.
 *
 * lrperm_stage_matrix_csr uploads the caller's FULL matrix (n_cells x n_genes, any column order within a
 * row; it may hold 2^31 stored entries or more - its row pointers are int64 on the device - as long as what
 * survives lrperm_commit_matrix stays below 2^31) and returns, over that matrix: col_sums[g] (float64, all rows; a gene is "empty" when
 * it is 0), the total over the rows flagged in row_keep (NULL = all rows; for mat_mean), the number of
 * non-finite stored values and of all-zero rows.  The outputs are HOST pointers (any may be NULL);
 * indptr / indices / data / row_keep follow `on_device`.
 * lrperm_commit_matrix then makes the staged matrix the engine's matrix, keeping the rows flagged in
 * row_keep and the columns with col_map[g] >= 0 renamed to col_map[g] in [0, n_genes_out) (host array,
 * at most one source per output column); it is equivalent to lrperm_set_matrix_csr on the filtered copy.
 */
int lrperm_stage_matrix_csr(lrperm_engine *h, int64_t n_cells, int64_t n_genes, const void *indptr, int indptr_is_64,
                            const int32_t *indices, const float *data, int on_device, const uint8_t *row_keep,
                            double *col_sums, double *kept_total, int64_t *n_nonfinite, int64_t *n_empty_rows);
/*
 * Shape of the staged matrix and n_unordered: the number of its stored entries whose column index is not greater than
 * its predecessor's in the same row, counted by the staging pass.  0 = every row strictly increasing, i.e. SciPy's
 * "canonical format" (sorted, no duplicate entries) - the property the reference's SciPy arithmetic relies on when it
 * sums a (cell, gene) entry once (`X[perm_idx]`, `.sum(0)`: liana/method/sc/_get_mean_perms.py:62-63).  The engine
 * itself accepts any column order but not duplicates: a caller that sees n_unordered > 0 sorts / sums duplicates on
 * its side (`csr_matrix.sum_duplicates`) and stages again - and does not need a pass of its own over every index
 * (SciPy's `has_canonical_format`: 0.5 ms per million entries) when it is 0.  Any output may be NULL.
 */
int lrperm_staged_info(lrperm_engine *h, int64_t *n_cells, int64_t *n_genes, int64_t *nnz, int64_t *n_unordered);
/*
 * The same staging step for a DENSE row-major matrix (adata.X as a 2-D array): replaces the `csr_matrix(X)` of
 * _choose_mtx_rep (liana/method/_pipe_utils/_pre.py:259-262) and the float32 cast (:117-119).  values = float32
 * (values_are_f64 == 0) or float64, n_cells rows of n_genes values, row_stride (in elements, >= n_genes) apart.  The
 * entries that compare != 0 in their own type become the stored entries (NaN included, -0.0 not), in column order,
 * rounded to float32.  The dense rows are on the device only for the conversion (host input: the whole matrix is
 * uploaded next to its CSR copy; LRPERM_ERR_OOM when that does not fit).  Outputs and row_keep as above.
 */
int lrperm_stage_matrix_dense(lrperm_engine *h, int64_t n_cells, int64_t n_genes, const void *values, int values_are_f64,
                              int64_t row_stride, int on_device, const uint8_t *row_keep,
                              double *col_sums, double *kept_total, int64_t *n_nonfinite, int64_t *n_empty_rows);
int lrperm_commit_matrix(lrperm_engine *h, const int32_t *col_map, int64_t n_genes_out);
/* Largest STORED value over the kept rows of the staged matrix (-inf if none) and the number of stored entries
 * in those rows: the inputs of CellChat's `mat_max = adata.X.max()` (liana/method/sc/_liana_pipe.py:143-146;
 * the implicit zeros take part in that maximum, the caller adds them).  Valid between stage and the next stage. */
int lrperm_staged_max(lrperm_engine *h, float *max_stored, int64_t *n_stored);
/* Per-cluster sum x and sum x^2 (float64) over ALL stored entries of the staged matrix, every gene: the inputs of
 * scSeqComm's `_cluster_stats` (cluster mean and standard deviation of the prepared matrix BEFORE the resource gene
 * subset, liana/method/sc/_liana_pipe.py:157-158, 466-476).  row_label[r] (host, int32, one per staged row) = cluster
 * of the row, -1 = row not kept.  Valid between stage and the next stage. */
int lrperm_staged_cluster_stats(lrperm_engine *h, const int32_t *row_label, int32_t n_clusters, double *sum_x, double *sum_sq);

/* Cluster code per cell (codes of obs['@label'] in .cat.categories order, _get_mean_perms.py:47-52). */
int lrperm_set_labels(lrperm_engine *h, const int32_t *labels, int32_t n_clusters, int on_device);

/*
 * The T rows of lr_res that reach the score function, as produced by _get_mat_idx
 * (_get_mean_perms.py:103-113): gene column of the resolved ligand / receptor subunit and
 * cluster index of source / target, plus the observed scores the null is compared against
 * (float32, widened to float64 in the comparison like np.greater_equal does).  Either
 * threshold pointer may be NULL if that score is not requested.
 */
int lrperm_set_triples(lrperm_engine *h, int64_t n_triples,
                       const int32_t *ligand_idx, const int32_t *receptor_idx,
                       const int32_t *source_idx, const int32_t *target_idx,
                       const float *obs_cpdb, const float *obs_gmean, int on_device);

/*
 * Observed per-cluster statistics (_liana_pipe.py:285-302, _common.py:41-42):
 * means[c*G+g] (float32, the same arithmetic as the cube with the identity permutation) and
 * stored_counts[c*G+g] (number of stored CSR entries, explicit zeros included) ; cluster
 * sizes n_c[c].  Any output pointer may be NULL.
 */
int lrperm_observed(lrperm_engine *h, float *means, int32_t *stored_counts, int32_t *cluster_sizes,
                    int out_on_device);

/*
 * Per-(cluster, gene) float64 moments of the observed matrix, inputs of the non-permutation scores
 * that rank_aggregate ranks next to the permutation p-values (_liana_pipe.py:259-273 z-scores via
 * sc.pp.scale, :341-360 1-vs-rest log2FC on `logbase`^x - 1):  sum_x[c*G+g], sum_sq[c*G+g] = sum x^2,
 * sum_expm1[c*G+g] = sum (logbase^x - 1) over the stored entries of cluster c.  Deterministic
 * (fixed-order segment partials).  Any output pointer may be NULL.
 */
int lrperm_cluster_moments(lrperm_engine *h, double logbase, double *sum_x, double *sum_sq, double *sum_expm1,
                           int out_on_device);

/*
 * Run `n_perms` permutations and accumulate exceedance counts.
 *   perm_source  LRPERM_PERMS_INDEX : perm_idx[p*N + j] = cell placed at position j in
 *                                     permutation p (int32 or int64), p = 0..n_perms-1
 *                LRPERM_PERMS_PHILOX: permutation p uses global id perm_offset + p and `seed`
 *   counts_*[t]  number of permutations whose null score >= observed score (uint32, T entries),
 *                OVERWRITTEN (not accumulated); NULL if that score is not requested
 *   cube_out     optional float32 [n_perms][L][G] (the reference's `perms` cube layout), or NULL
 */
int lrperm_run(lrperm_engine *h, int64_t n_perms, int64_t perm_offset, uint64_t seed,
               int perm_source, const void *perm_idx, int perm_idx_is_64, int perm_idx_on_device,
               int order_mode, int scores, int gmean_mode,
               uint32_t *counts_cpdb, uint32_t *counts_gmean, int counts_on_device,
               float *cube_out, int cube_on_device);

/*
 * CellChat null (liana/method/sc/_cellchat.py:7-32): the per-(permutation, cluster, gene) statistic is the
 * TRIMEAN (q25 + 2 median + q75) / 4 of the cluster's dense column of X / mat_max (_trimean,
 * _liana_pipe.py:462-465, called through _get_means_perms with norm_factor, :394-397 and
 * _get_mean_perms.py:43-44) and the null score is prod / (kh + prod) of the ligand and receptor trimeans,
 * compared with >= in float64.  The engine selects the six order statistics a trimean needs from the stored
 * entries (zeros accounted for arithmetically) and applies numpy's interpolation arithmetic: results are
 * bit-identical to the reference for identical permutations (no summation order is involved).
 *
 * SciPy evaluates the reference's two divisions differently, and both are followed:
 *   null      `adata.X /= norm_factor`  is  data *= (1.0 / norm_factor) in float32  -> pass recip32 = float32(1 / mat_max)
 *   observed  `temp.X / mat_max`        is  X.astype(float64) * (1 / mat_max)       -> pass recip    = (double)(1 / mat_max)
 * (`1 / np.float32` is a float32 under NumPy >= 2 and a float64 before: the caller evaluates it, see _engine.py).
 *
 * lrperm_run_trimean: permutations as in lrperm_run; the triples' index arrays come from lrperm_set_triples
 * (its thresholds may be NULL); obs_prob[T] = the observed lr_probs (float64); counts[T] (uint32, overwritten) =
 * number of permutations with null score >= observed, or NULL (then obs_prob may be NULL and no triples are
 * needed); cube_out = optional float64 [n_perms][L][G] trimean cube (the reference's `perms` layout).
 * lrperm_observed_trimean: trimean[c*G + g] (float64) of the unpermuted clusters (_liana_pipe.py:307-308).
 * Limits: n_cells < 2^28, at most 255 clusters.
 */
int lrperm_run_trimean(lrperm_engine *h, int64_t n_perms, int64_t perm_offset, uint64_t seed,
                       int perm_source, const void *perm_idx, int perm_idx_is_64, int perm_idx_on_device,
                       float recip32, double kh, const double *obs_prob, int obs_on_device,
                       uint32_t *counts, int counts_on_device, double *cube_out, int cube_on_device);
int lrperm_observed_trimean(lrperm_engine *h, double recip, double *trimean, int out_on_device);
/* Device time (ms) of the stages of the last trimean run, summed over batches: out[0] entry counting
 * (the dominant kernel), out[1] prefix, out[2] rank selection, out[3] interpolation; out[4] the number of
 * (chunk, permutation group) warps of the selection pass that had to re-walk their chunk, out[5] all of them. */
int lrperm_trimean_timings(lrperm_engine *h, float *out6);

/*
 * expr_prop filter, min-subunit resolution of the complexes and name -> index maps on the device
 * (filter_reassemble_complexes, liana/resource/_reassemble_complexes.py:11-101; _get_mat_idx,
 * liana/method/_pipe_utils/_get_mean_perms.py:90-113; _join_stats, _pipe_utils/_common.py:5-39), which the reference
 * does with pandas merges over L^2 x 5,780 exploded rows and an O(T x G) Python dictionary comprehension.
 *   lig_sub / rec_sub [n_inter][max_l | max_r]  gene column of every subunit of the ligand / receptor complex, in the
 *                                               complex's '_' order, -1 padded (host)
 *   stat [L][G]            the statistic named by the method's complex_cols: cluster means (float32) or CellChat's
 *                          cluster trimeans (float64, stat_is_f64 = 1); props [L][G] float64; pair_mask [L][L] or NULL
 *   first_subunit          1 = every complex is represented by its first subunit (return_all_lrs de-duplicates before
 *                          the min-subunit step, :52-57), else the FIRST subunit with the minimal statistic
 *   return_all             1 = rows that fail expr_prop are emitted too (keep = 0)
 * Rows come out in the reference's order (per (source, target) pair with the source varying fastest - target-major,
 * `np.meshgrid(labels, labels).reshape(2, L*L).T`, _liana_pipe.py:310-312 - then interaction).  lrperm_resolve_triples returns
 * the row count; lrperm_fetch_triples copies the columns to host arrays of that length (all required): interaction index, source / target cluster, resolved ligand / receptor gene column, their
 * statistic (float32 or float64 like `stat`) and props, and the expr_prop flag.
 */
int lrperm_resolve_triples(lrperm_engine *h, int32_t n_inter, int32_t max_l, int32_t max_r,
                           const int32_t *lig_sub, const int32_t *rec_sub, const void *stat, int stat_is_f64,
                           const double *props, double expr_prop, const uint8_t *pair_mask, int first_subunit,
                           int return_all, int64_t *n_rows);
int lrperm_fetch_triples(lrperm_engine *h, int32_t *inter, int32_t *src, int32_t *tgt, int32_t *lig_gene, int32_t *rec_gene,
                         void *lig_stat, void *rec_stat, double *lig_props, double *rec_props, uint8_t *keep);

/*
 * Consensus of method scores (liana/method/_pipe_utils/_aggregate.py:70-181), the step that turns the permutation
 * p-values and the other scores into `specificity_rank` / `magnitude_rank`:
 *   lrperm_rank_average  ranks[i] = scipy.stats.rankdata(negate ? -values : values, method='average')[i]   (:91-98);
 *                        float64, ties share the average rank, -0.0 ties with +0.0; values must not contain NaN
 *   lrperm_rra           RobustRankAggregate of Kolde et al. 2012 (:110-181): rmat[row][col] are ranks (row-major,
 *                        n_cols <= 16), col_max[col] their column maxima; rho[row] = min(1, n_cols * min_k
 *                        I_{r_(k)}(k, n_cols - k + 1)) with r = sorted(rmat[row] / col_max)
 * Host or device pointers (`on_device`), independent of the matrix / label state of the engine.
 */
int lrperm_rank_average(lrperm_engine *h, int64_t n, const double *values, int negate, double *ranks, int on_device);
/* order[k] = index of the k-th row when sorting by `values` (stable; descending = largest first): the final
 * `liana_res.sort_values(magnitude)` of liana_pipe (liana/method/sc/_liana_pipe.py:246-250).  No NaN in `values`. */
int lrperm_argsort(lrperm_engine *h, int64_t n, const double *values, int descending, int32_t *order, int on_device);
int lrperm_rra(lrperm_engine *h, int64_t n_rows, int32_t n_cols, const double *rmat, const double *col_max, double *rho, int on_device);

/* Export Philox permutation(s) as index arrays: out[p*N + j] = cell at position j (int32), i.e. the array
 * that reproduces a PHILOX run when passed back as LRPERM_PERMS_INDEX.  The device bijection rho_p (keyed by
 * seed and global permutation id, include/lrperm_philox.h) acts on CLUSTER-SORTED positions: position
 * order[k] (k-th position in stable label order) receives cell rho_p(k), so the label of a cell follows from
 * k = rho_p^-1(cell) and the cluster boundaries alone.  Requires the labels to be set; n_cells must equal
 * the engine's cell count. */
int lrperm_philox_perms(lrperm_engine *h, int64_t n_cells, int64_t n_perms, int64_t perm_offset,
                        uint64_t seed, int32_t *out, int out_on_device);

/* Device time (ms, CUDA events on the engine's stream) of the stages of the last lrperm_run:
 * out[0] label/permutation setup, out[1] per-cluster means (all kernels), out[2] score+count,
 * out[3] total, out[4] number of kernel launches, out[5] time of the dominant means kernel alone
 * (summed over batches), out[6] its launch count, out[7] permutations per batch. */
int lrperm_last_timings(lrperm_engine *h, float *out8);

/* Work statistics of the last lrperm_run in FAST order: out[0] genes accumulated, out[1] stored entries of those genes,
 * out[2] (gene, cell tile) chunks launched per permutation group, out[3] stored entries of the whole matrix.  With the
 * option "skip_unreferenced" (default 1) a run that exports no cube accumulates only the genes some triple of
 * lrperm_set_triples reads (the reference computes every gene of the subset matrix, _get_mean_perms.py:61-64, and then
 * gathers only those, _liana_pipe.py:416-419); the counts are identical either way. */
int lrperm_run_stats(lrperm_engine *h, int64_t *out4);

/* Debug build only (-DLRPERM_BOUNDS_CHECK, `python -m liana_py_b200.build --debug` -> liblrperm_debug.so; the production library
 * answers LRPERM_ERR_UNSUPPORTED): the hot kernels test every index they compute - shared-memory bin offsets, accumulator columns,
 * cells, chunk ranges, partial-sum slots, cube rows - against the engine's extents.  Returns and clears the violation bits since
 * the last call (0 = none; 1 bin offset, 2 cell, 4 chunk range, 8 slot, 16 column, 32 row extent, 64 cube row, 128 trimean plane) and,
 * in first2 (may be NULL), the bit and the value of the first violation.  Stands in for compute-sanitizer where that tool is closed. */
int lrperm_debug_bounds(lrperm_engine *h, uint32_t *flags, int64_t *first2);

/* Shape of the engine's matrix: out[0] cells, out[1] genes, out[2] stored entries. */
int lrperm_matrix_info(lrperm_engine *h, int64_t *out3);
/* Copy the engine's matrix (after lrperm_set_matrix_csr / lrperm_commit_matrix) out as CSR: indptr int32 [N+1], indices
 * int32 [nnz], data float32 [nnz]; the order of the entries inside a row is unspecified.  out_on_device != 0: device
 * pointers, which may belong to ANOTHER CUDA device (unified addressing: the copy then goes peer to peer over NVLink) -
 * this is how a committed matrix is replicated to the engines of other GPUs without crossing PCIe again. */
int lrperm_export_matrix_csr(lrperm_engine *h, int32_t *indptr, int32_t *indices, float *data, int out_on_device);

/* Tuning knobs of FAST mode (0 = engine default): permutations per batch (multiple of 128), stored
 * entries per gene chunk, and MiB of label slab one cell tile may span (kept L2 resident). */
int lrperm_set_tuning(lrperm_engine *h, int32_t fast_perm_batch, int32_t fast_chunk_nnz, int32_t fast_tile_mib);

/* Named engine options (experiments / A-B measurements; defaults are the production settings):
 *   "fast_variant"  2 (default) = fast2_means_kernel, 1 = fast_means_kernel (round-1 first version)
 *   "fast_q"        permutations per lane in FAST mode: 0 = auto, 2, 4 or 8
 *   "eager_csc"           1 (default) = lrperm_set_matrix_csr / lrperm_commit_matrix queue the CSC copy of FAST mode at once
 *   "score_groups_per_block"  32-permutation groups per block of the score kernels (0 = auto: the gathered cube slice fits L2)
 *   "slab0_by_tile"       1 (default) = the first batch's label slab is generated and consumed cell tile by cell tile
 *   "fast_first_batch"    permutations of a short first batch of the pipelined FAST path (0 = none, default)
 *   "trimean_chunk_nnz"   stored entries per chunk of the trimean path (0 = auto; multiple of 1024)
 *   "trimean_perm_batch"  permutations per batch of the trimean path (0 = auto; multiple of 128)
 *   "skip_unreferenced"   1 (default) = runs that export no cube skip the genes no triple reads (see lrperm_run_stats)
 *   "slab_ctas_per_sm"    CTAs per SM of the (persistent) label-slab kernel of the pipelined FAST path (0 = auto, default: a quarter of
 *                         the main kernel's resident CTAs): it runs next to the main kernel and must not flood the warp schedulers
 *   "slab_mode"           which label-slab launches use that persistent grid: 2 (default) all but the very first tile of a run,
 *                         1 all, 3 batches >= 1 only, 0 none (one CTA per block of cells, round 1's behaviour)
 *   "narrow_indices"      1 (default) = lrperm_stage_matrix_csr sends the column indices of a large pageable upload as 16 bits
 *                         (n_genes <= 65536): less host memory traffic and PCIe volume; same staged matrix  */
int lrperm_set_option(lrperm_engine *h, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* LRPERM_H */
