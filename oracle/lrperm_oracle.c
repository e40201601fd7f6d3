/*
 * lrperm_oracle.c - TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's permutation
 * cube arithmetic.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library; the product path never does.
 *
 * Parity status: PINNED.  tests/golden/ holds cubes, observed means and p-values produced by
 * the reference's own code (liana/method/_pipe_utils/_get_mean_perms.py executed unmodified in
 * the build container through oracle/ref_shim.py; generator: scripts/make_golden.py);
 * tests/test_oracle.py checks this file against them bit-for-bit.
 *
 * What is restated (paths relative to /root/reference):
 *   oracle_cube      _generate_perms_cube + _permute_and_aggregate
 *                    liana/method/_pipe_utils/_get_mean_perms.py:61-87 with agg_fun = np.mean
 *                    (liana/method/sc/_liana_pipe.py:394-400).  np.mean on a SciPy CSR matrix
 *                    is scipy.sparse _base.mean(axis=0): (X * float32(1/n)).sum(0), evaluated
 *                    by csc_matvecs as a sequential float32 accumulation over the rows in
 *                    order (scipy is a pinned dependency, not vendored: scipy 1.10.1,
 *                    /root/reference/poetry.lock:3574).  SURVEY.md Appendix B.
 *   oracle_observed  per-cluster means / stored-entry counts, _liana_pipe.py:285-302,
 *                    liana/method/_pipe_utils/_common.py:41-42.
 *   oracle_counts    _calculate_pvals (_get_mean_perms.py:117-142) with the CellPhoneDB null
 *                    score (liana/method/sc/_cellphonedb.py:28): mean of two float64 values.
 *                    (The geometric-mean null uses numpy's own exp/log and is restated in
 *                    oracle/numpy_port.py.)
 *   oracle_trimean_cube / oracle_counts_cellchat
 *                    the CellChat null: agg_fun = _trimean (liana/method/sc/_liana_pipe.py:394-397,
 *                    462-465) on X / norm_factor (float32 division, _get_mean_perms.py:43-44), and
 *                    _lr_probability (liana/method/sc/_cellchat.py:7-10).  np.quantile / np.median
 *                    are numpy (pinned 1.24.4, poetry.lock:2233; not vendored); restated from the
 *                    published algorithm (Hyndman & Fan method 7, numpy/lib/function_base.py _lerp)
 *                    WITHOUT densifying: the order statistics of a cluster's column are read off the
 *                    sorted stored values plus the implicit zeros.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "../include/lrperm_philox.h"

/* cube[p][c][g] float32.  perm_idx[p*N + j] = cell placed at position j. */
int oracle_cube(int64_t N, int64_t G, const int64_t *indptr, const int32_t *indices, const float *data,
                const int32_t *labels, int32_t L, int64_t P, const int64_t *perm_idx, float *cube)
{
    if (N < 0 || G <= 0 || L <= 0 || P < 0) return 1;
    int64_t *n_c = (int64_t *)calloc((size_t)L, sizeof(int64_t));
    float *inv = (float *)malloc(sizeof(float) * (size_t)L);
    if (!n_c || !inv) return 2;
    for (int64_t j = 0; j < N; ++j) {
        if (labels[j] < 0 || labels[j] >= L) { free(n_c); free(inv); return 1; }
        n_c[labels[j]]++;
    }
    for (int32_t c = 0; c < L; ++c) inv[c] = (float)(1.0 / (double)n_c[c]);   /* float64 division, one rounding */

    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t p = 0; p < P; ++p) {
        float *out = cube + (size_t)p * (size_t)L * (size_t)G;
        memset(out, 0, sizeof(float) * (size_t)L * (size_t)G);
        const int64_t *perm = perm_idx ? perm_idx + (size_t)p * (size_t)N : NULL;
        /* ascending position j: for a fixed (cluster, gene) this is ascending k in pos_c */
        for (int64_t j = 0; j < N; ++j) {
            const int32_t c = labels[j];
            const int64_t cell = perm ? perm[j] : j;
            float *row = out + (size_t)c * (size_t)G;
            const float w = inv[c];
            for (int64_t o = indptr[cell]; o < indptr[cell + 1]; ++o) {
                const float prod = data[o] * w;          /* fl32(x * inv_c) */
                row[indices[o]] = row[indices[o]] + prod; /* fl32(acc + prod); -ffp-contract=off */
            }
        }
    }
    free(n_c); free(inv);
    return 0;
}

/* means[c][g] (float32, identity permutation), stored[c][g] (stored CSR entries), sizes[c] */
int oracle_observed(int64_t N, int64_t G, const int64_t *indptr, const int32_t *indices, const float *data,
                    const int32_t *labels, int32_t L, float *means, int32_t *stored, int32_t *sizes)
{
    int rc = oracle_cube(N, G, indptr, indices, data, labels, L, 1, NULL, means);
    if (rc) return rc;
    memset(stored, 0, sizeof(int32_t) * (size_t)L * (size_t)G);
    memset(sizes, 0, sizeof(int32_t) * (size_t)L);
    for (int64_t j = 0; j < N; ++j) {
        const int32_t c = labels[j];
        sizes[c]++;
        for (int64_t o = indptr[j]; o < indptr[j + 1]; ++o) stored[(size_t)c * (size_t)G + indices[o]]++;
    }
    return 0;
}

/* CellPhoneDB exceedance counts from a float32 cube[p][c][g]:
 * count[t] = #{p : (a + b)/2 >= (double)obs[t]},  a = cube[p][src][lig], b = cube[p][tgt][rec] */
int oracle_counts_cpdb(int64_t P, int32_t L, int64_t G, const float *cube, int64_t T,
                       const int32_t *lig, const int32_t *rec, const int32_t *src, const int32_t *tgt,
                       const float *obs, uint32_t *counts)
{
    #pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        uint32_t n = 0;
        const double thr = (double)obs[t];
        for (int64_t p = 0; p < P; ++p) {
            const float *slab = cube + (size_t)p * (size_t)L * (size_t)G;
            const double a = (double)slab[(size_t)src[t] * (size_t)G + lig[t]];
            const double b = (double)slab[(size_t)tgt[t] * (size_t)G + rec[t]];
            if ((a + b) / 2.0 >= thr) ++n;
        }
        counts[t] = n;
    }
    return 0;
}

/* host evaluation of the Philox bijection (same header as the device code): out[p*N + j] */
int oracle_philox_perms(int64_t N, int64_t P, int64_t perm_offset, uint64_t seed, int32_t *out)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t p = 0; p < P; ++p) {
        lrperm_bij_t b = lrperm_bij_init(seed, (uint64_t)(perm_offset + p), (uint32_t)N);
        for (int64_t j = 0; j < N; ++j) out[(size_t)p * (size_t)N + j] = (int32_t)lrperm_bij_forward(&b, (uint32_t)j);
    }
    return 0;
}

int oracle_philox_inverse(int64_t N, int64_t perm_id, uint64_t seed, int32_t *out)
{
    lrperm_bij_t b = lrperm_bij_init(seed, (uint64_t)perm_id, (uint32_t)N);
    for (int64_t i = 0; i < N; ++i) out[i] = (int32_t)lrperm_bij_inverse(&b, (uint32_t)i);
    return 0;
}

/* ---- CellChat: trimean of every (permutation, cluster, gene) column ----------------------------------
 * Two precisions, as in the reference (SciPy scalar division semantics, scipy/sparse/_base.py, _data.py):
 *   null      adata.X /= norm_factor  ->  data *= (1.0 / norm): float32 values fl32(x * r32); np.quantile on the
 *             float32 matrix takes the neighbour difference in float32 and interpolates with float64 weights,
 *             np.median averages in float32 (_get_mean_perms.py:43-44, _liana_pipe.py:462-465)
 *   observed  temp.X / mat_max        ->  X.astype(float64) * (1 / mat_max): float64 values (double)x * r, and every
 *             step of _trimean in float64 (_liana_pipe.py:307-308)
 * `r` is whatever `1 / np.float32` evaluates to under the NumPy in use (float32 under NEP 50, float64 before);
 * the caller passes it, so the same code follows either. */
static int cmp_float_asc(const void *a, const void *b)
{
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* i-th smallest (0-based) raw value of a column made of `k` stored values (sorted ascending in v) and n - k implicit zeros */
static float column_order_stat(const float *v, int64_t k, int64_t n, int64_t n_neg, int64_t i)
{
    const int64_t n_zero = n - k;                      /* implicit zeros sit between the negatives and the rest */
    if (i < n_neg) return v[i];
    if (i < n_neg + n_zero) return 0.0f;
    return v[i - n_zero];
}

typedef struct { int observed; float r32; double r64; } tm_scale_t;

/* numpy _lerp(a, b, t): diff in the array dtype, weights float64 */
static double lerp_np(const tm_scale_t *sc, float xa, float xb, double t)
{
    double a, b, d;
    if (sc->observed) { a = (double)xa * sc->r64; b = (double)xb * sc->r64; d = b - a; }
    else { const float fa = xa * sc->r32, fb = xb * sc->r32; const float fd = fb - fa; a = fa; b = fb; d = fd; }
    double r = a + d * t;
    if (t >= 0.5) r = b - d * (1.0 - t);
    return r;
}

static double quantile_np(const tm_scale_t *sc, const float *v, int64_t k, int64_t n, int64_t n_neg, double q)
{
    const double virt = (double)(n - 1) * q;           /* method 'linear': (n - 1) * q */
    int64_t lo = (int64_t)floor(virt), hi = lo + 1;
    if (virt >= (double)(n - 1)) { lo = n - 1; hi = n - 1; }
    const double gamma = virt - floor(virt);
    return lerp_np(sc, column_order_stat(v, k, n, n_neg, lo), column_order_stat(v, k, n, n_neg, hi), gamma);
}

static double trimean_column(const tm_scale_t *sc, const float *sorted_vals, int64_t k, int64_t n)
{
    int64_t n_neg = 0;
    while (n_neg < k && sorted_vals[n_neg] < 0.0f) ++n_neg;
    const double q25 = quantile_np(sc, sorted_vals, k, n, n_neg, 0.25);
    const double q75 = quantile_np(sc, sorted_vals, k, n, n_neg, 0.75);
    const float xa = column_order_stat(sorted_vals, k, n, n_neg, (n & 1) ? n / 2 : n / 2 - 1);
    const float xb = column_order_stat(sorted_vals, k, n, n_neg, n / 2);
    double med2;                                       /* 2 * np.median: np.mean of the middle one / two values */
    if (sc->observed) med2 = 2.0 * (((double)xa * sc->r64 + (double)xb * sc->r64) / 2.0);
    else med2 = (double)(2.0f * ((xa * sc->r32 + xb * sc->r32) / 2.0f));
    return ((q25 + med2) + q75) / 4.0;
}

/* cube[p][c][g] float64 = trimean over the cells placed in cluster c by permutation p of column g.
 * perm_idx NULL + observed = 1: the observed statistic (identity, float64 path); otherwise the null path. */
int oracle_trimean_cube(int64_t N, int64_t G, const int64_t *indptr, const int32_t *indices, const float *data,
                        const int32_t *labels, int32_t L, int64_t P, const int64_t *perm_idx, int observed,
                        float recip32, double recip64, double *cube)
{
    if (N < 0 || G <= 0 || L <= 0 || P < 0) return 1;
    const tm_scale_t sc = {observed, recip32, recip64};
    const int64_t nnz = indptr[N];
    int64_t *cptr = (int64_t *)calloc((size_t)G + 1, sizeof(int64_t));
    int64_t *n_c = (int64_t *)calloc((size_t)L, sizeof(int64_t));
    int32_t *ccell = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    float *cval = (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1));
    if (!cptr || !n_c || !ccell || !cval) return 2;
    for (int64_t o = 0; o < nnz; ++o) cptr[indices[o] + 1]++;
    for (int64_t g = 0; g < G; ++g) cptr[g + 1] += cptr[g];
    {
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)G);
        memcpy(fill, cptr, sizeof(int64_t) * (size_t)G);
        for (int64_t i = 0; i < N; ++i)
            for (int64_t o = indptr[i]; o < indptr[i + 1]; ++o) {
                const int64_t d = fill[indices[o]]++;
                ccell[d] = (int32_t)i;
                cval[d] = data[o];
            }
        free(fill);
    }
    for (int64_t j = 0; j < N; ++j) n_c[labels[j]]++;

    #pragma omp parallel
    {
        int32_t *lab_cell = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
        float *bucket = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));   /* grouped by cluster */
        int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * ((size_t)L + 1));
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
        #pragma omp for schedule(dynamic, 1)
        for (int64_t p = 0; p < P; ++p) {
            const int64_t *perm = perm_idx ? perm_idx + (size_t)p * (size_t)N : NULL;
            for (int64_t j = 0; j < N; ++j) lab_cell[perm ? perm[j] : j] = labels[j];   /* position j holds cell perm[j] */
            for (int64_t g = 0; g < G; ++g) {
                memset(cnt, 0, sizeof(int64_t) * ((size_t)L + 1));
                for (int64_t o = cptr[g]; o < cptr[g + 1]; ++o) cnt[lab_cell[ccell[o]] + 1]++;
                for (int32_t c = 0; c < L; ++c) cnt[c + 1] += cnt[c];
                for (int32_t c = 0; c < L; ++c) fill[c] = cnt[c];
                for (int64_t o = cptr[g]; o < cptr[g + 1]; ++o) bucket[fill[lab_cell[ccell[o]]]++] = cval[o];
                for (int32_t c = 0; c < L; ++c) {
                    float *v = bucket + cnt[c];
                    const int64_t k = cnt[c + 1] - cnt[c];
                    qsort(v, (size_t)k, sizeof(float), cmp_float_asc);
                    cube[((size_t)p * (size_t)L + (size_t)c) * (size_t)G + (size_t)g] = trimean_column(&sc, v, k, n_c[c]);
                }
            }
        }
        free(lab_cell); free(bucket); free(cnt); free(fill);
    }
    free(cptr); free(n_c); free(ccell); free(cval);
    return 0;
}

/* CellChat exceedance counts from a float64 trimean cube[p][c][g]: prob = ab / (kh + ab) >= obs */
int oracle_counts_cellchat(int64_t P, int32_t L, int64_t G, const double *cube, int64_t T,
                           const int32_t *lig, const int32_t *rec, const int32_t *src, const int32_t *tgt,
                           const double *obs, double kh, uint32_t *counts)
{
    #pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        uint32_t n = 0;
        for (int64_t p = 0; p < P; ++p) {
            const double *slab = cube + (size_t)p * (size_t)L * (size_t)G;
            const double prod = slab[(size_t)src[t] * (size_t)G + lig[t]] * slab[(size_t)tgt[t] * (size_t)G + rec[t]];
            if (prod / (kh + prod) >= obs[t]) ++n;
        }
        counts[t] = n;
    }
    return 0;
}
