/*
 * lrperm_philox.h - counter-based label shuffles for the permutation null.
 *
 * The reference draws P sequential permutations from ONE NumPy PCG64 stream
 * (`liana/method/_pipe_utils/_get_mean_perms.py:69,79`), which cannot be sharded over
 * GPUs.  Production mode replaces it with a stateless, order-free construction keyed by
 * (seed, global permutation id):
 *
 *   1. Philox4x32-10 (Salmon et al., SC'11) turns (seed, perm id) into four 32-bit words;
 *      round keys are those words plus a Weyl sequence (Philox's own key schedule).
 *   2. A variable-width Feistel network over k = ceil(log2 N) bits, whose round function is
 *      a multiply/xor-shift mixer built from the Philox multipliers, is a bijection on [0, 2^k).
 *   3. Cycle walking (re-apply until the value is < N) restricts it to a bijection on [0, N).
 *
 * lrperm_bij_forward(j) plays the role of the reference's `perm_idx[j]` (the cell placed at
 * position j); lrperm_bij_inverse(i) is the position cell i lands on.  Both are pure
 * functions, so any rank / thread can evaluate any element of any permutation.
 *
 * This header is plain C (usable from gcc for the host-side checker) and CUDA.
 */
#ifndef LRPERM_PHILOX_H
#define LRPERM_PHILOX_H

#include <stdint.h>

#if defined(__CUDACC__)
#define LRPERM_HD __host__ __device__ __forceinline__
#else
#define LRPERM_HD static inline
#endif

#define LRPERM_PHILOX_M0 0xD2511F53u
#define LRPERM_PHILOX_M1 0xCD9E8D57u
#define LRPERM_PHILOX_W0 0x9E3779B9u
#define LRPERM_PHILOX_W1 0xBB67AE85u
#define LRPERM_BIJ_ROUNDS 10

typedef struct {
    uint32_t k[4];      /* Philox output words for (seed, perm id) */
    uint32_t lbits;     /* bits in the high half */
    uint32_t rbits;     /* bits in the low half  */
    uint32_t n;         /* domain size N         */
} lrperm_bij_t;

LRPERM_HD uint32_t lrperm_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox4x32-10: counter c[4], key (k0,k1) -> out[4] */
LRPERM_HD void lrperm_philox4x32_10(const uint32_t c[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = lrperm_mulhi32(LRPERM_PHILOX_M0, x0), lo0 = LRPERM_PHILOX_M0 * x0;
        uint32_t hi1 = lrperm_mulhi32(LRPERM_PHILOX_M1, x2), lo1 = LRPERM_PHILOX_M1 * x2;
        uint32_t y0 = hi1 ^ x1 ^ k0;
        uint32_t y1 = lo1;
        uint32_t y2 = hi0 ^ x3 ^ k1;
        uint32_t y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        k0 += LRPERM_PHILOX_W0;
        k1 += LRPERM_PHILOX_W1;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

LRPERM_HD lrperm_bij_t lrperm_bij_init(uint64_t seed, uint64_t perm_id, uint32_t n) {
    lrperm_bij_t b;
    uint32_t c[4];
    c[0] = (uint32_t)perm_id; c[1] = (uint32_t)(perm_id >> 32); c[2] = 0x6c72706du; /* "lrpm" */ c[3] = 0u;
    lrperm_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32), b.k);
    /* at least a 4-bit domain: with 1-bit halves the family of round functions is too small and the joint
     * position of two cells is measurably non-uniform (tests/test_oracle.py); N <= 8 cycle-walks from 16 */
    uint32_t bits = 4;
    while (bits < 32 && ((uint32_t)1 << bits) < n) ++bits;
    b.lbits = bits / 2;
    b.rbits = bits - b.lbits;
    b.n = n;
    return b;
}

LRPERM_HD uint32_t lrperm_bij_f(uint32_t v, uint32_t key) {
    uint32_t h = (v ^ key) * LRPERM_PHILOX_M0;
    h ^= h >> 16;
    h *= LRPERM_PHILOX_M1;
    h ^= h >> 15;
    h *= 0x9E3779B1u;
    h ^= h >> 13;
    return h;
}

LRPERM_HD uint32_t lrperm_bij_rk(const lrperm_bij_t* b, int r) {
    return b->k[r & 3] + (uint32_t)r * LRPERM_PHILOX_W0;
}

/* one pass of the Feistel network on [0, 2^(lbits+rbits)) */
LRPERM_HD uint32_t lrperm_bij_enc(const lrperm_bij_t* b, uint32_t x) {
    const uint32_t mL = ((uint32_t)1 << b->lbits) - 1u, mR = ((uint32_t)1 << b->rbits) - 1u;
    uint32_t Lh = x >> b->rbits, Rh = x & mR;
    for (int r = 0; r < LRPERM_BIJ_ROUNDS; ++r) {
        if ((r & 1) == 0) Lh ^= lrperm_bij_f(Rh, lrperm_bij_rk(b, r)) & mL;
        else              Rh ^= lrperm_bij_f(Lh, lrperm_bij_rk(b, r)) & mR;
    }
    return (Lh << b->rbits) | Rh;
}

LRPERM_HD uint32_t lrperm_bij_dec(const lrperm_bij_t* b, uint32_t x) {
    const uint32_t mL = ((uint32_t)1 << b->lbits) - 1u, mR = ((uint32_t)1 << b->rbits) - 1u;
    uint32_t Lh = x >> b->rbits, Rh = x & mR;
    for (int r = LRPERM_BIJ_ROUNDS - 1; r >= 0; --r) {
        if ((r & 1) == 0) Lh ^= lrperm_bij_f(Rh, lrperm_bij_rk(b, r)) & mL;
        else              Rh ^= lrperm_bij_f(Lh, lrperm_bij_rk(b, r)) & mR;
    }
    return (Lh << b->rbits) | Rh;
}

/* cell placed at position j (the analogue of the reference's perm_idx[j]) */
LRPERM_HD uint32_t lrperm_bij_forward(const lrperm_bij_t* b, uint32_t j) {
    uint32_t x = lrperm_bij_enc(b, j);
    while (x >= b->n) x = lrperm_bij_enc(b, x);
    return x;
}

/* position that cell i lands on (inverse of lrperm_bij_forward) */
LRPERM_HD uint32_t lrperm_bij_inverse(const lrperm_bij_t* b, uint32_t i) {
    uint32_t x = lrperm_bij_dec(b, i);
    while (x >= b->n) x = lrperm_bij_dec(b, x);
    return x;
}

#endif /* LRPERM_PHILOX_H */
