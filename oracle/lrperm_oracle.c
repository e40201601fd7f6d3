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
