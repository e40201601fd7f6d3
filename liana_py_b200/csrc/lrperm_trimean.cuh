// lrperm_trimean.cuh - sm_100a kernels of the CellChat null: the TRIMEAN of every
// (permutation, cluster, gene) column instead of its mean.
//
// Reference being reproduced (paths relative to /root/reference):
//   liana/method/sc/_liana_pipe.py:394-397      agg_fun = _trimean, norm_factor = mat_max
//   liana/method/sc/_liana_pipe.py:462-465      _trimean = (q25 + 2*median + q75) / 4 on the DENSE column
//   liana/method/_pipe_utils/_get_mean_perms.py:43-44, 61-64   X /= norm_factor ; agg_fun(perm_mat[mask], axis=0)
//   liana/method/sc/_liana_pipe.py:307-308      observed: _trimean(temp.X / mat_max)
//   liana/method/sc/_cellchat.py:7-32           _lr_probability = prod / (kh + prod), >= , count
// The reference densifies every cluster's rows and calls np.quantile / np.median per permutation.  Here the
// six order statistics a trimean needs are SELECTED from the stored entries (the implicit zeros are
// accounted for arithmetically), exactly - no summation-order question exists, results are bit-identical.
//
// Data: vsc_pairs[nnz] int2 {cell, float bits}: gene-major copy of X whose entries are ordered, inside a gene,
// positives DESCENDING, then explicit zeros, then negatives ASCENDING.  Order statistic i (0-based, ascending)
// of a column with n rows is: the (n - i)-th entry of the cluster met when walking the positives (rank from the
// top), or the (i + 1)-th met when walking the negatives (rank from the bottom), else 0.
//
// Per batch of permutations (label slab as in FAST mode):
//   1. count   fast2_means_kernel<COUNT>: entries per (chunk, cluster, permutation)            - the heavy pass
//   2. prefix  tm_prefix_kernel: exclusive prefix of the counts over the chunks of a (gene, sign) region; lists the
//              (chunk, permutation group) items in which a wanted rank falls (a few % of all)
//   3. select  tm_select_kernel: persistent warps re-walk the listed chunks with a per-(cluster, permutation)
//              countdown and store the value met at 0
//   4. final   tm_finalize_kernel: numpy's interpolation arithmetic on the selected values -> float64 cube
//   5. score   score_count_cellchat_kernel
#pragma once

#include "lrperm_kernels.cuh"

namespace lrperm {

constexpr int kTmDirBit = 30;                 // FastChunk.gene bit: 1 = negative region (ranks from the bottom)
constexpr uint32_t kTmNever = 0x7fffffffu;    // countdown state that no chunk can exhaust
constexpr int kTmPlanes = 6;                  // distinct order statistics a trimean can need

// 32-bit sort key inside a gene: positives descending, zeros, negatives ascending
__device__ __forceinline__ uint32_t tm_value_key(float v) {
    const uint32_t b = (uint32_t)__float_as_int(v);
    if (v > 0.0f) return 0x7fffffffu - b;                  // larger value -> smaller key; all < 0x80000000
    if (v < 0.0f) return 0xffffffffu - (b & 0x7fffffffu);  // larger magnitude -> smaller key; all > 0x80000000
    return 0x80000000u;                                    // +-0 (NaN never reaches here: rejected at staging)
}

// CSR -> value-sorted CSC, step 1: 64-bit key {column, value key}, payload {value bits, cell}
__global__ void vsc_keys_kernel(const int32_t* __restrict__ indptr, const int2* __restrict__ pairs, int32_t N, int32_t G,
                                unsigned long long* __restrict__ keys, unsigned long long* __restrict__ payload,
                                int32_t* __restrict__ col_count) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < N; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        for (int o = s + lane; o < e; o += 32) {
            const int2 pr = pairs[o];
            keys[o] = ((unsigned long long)(uint32_t)pr.x << 32) | tm_value_key(__int_as_float(pr.y));
            payload[o] = ((unsigned long long)(uint32_t)pr.y << 32) | (uint32_t)r;
            atomicAdd(&col_count[pr.x], 1);
        }
    }
}

// per gene: first entry that is not positive, first entry that is negative (both predicates are monotone in the
// sorted order)
__global__ void tm_bounds_kernel(const int2* __restrict__ vsc_pairs, const int32_t* __restrict__ ptr, int32_t G,
                                 int32_t* __restrict__ pos_end, int32_t* __restrict__ neg_begin) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int b = ptr[g], e = ptr[g + 1];
    int lo = b, hi = e;
    while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (__int_as_float(vsc_pairs[mid].y) > 0.0f) lo = mid + 1; else hi = mid; }
    pos_end[g] = lo;
    hi = e;
    while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (!(__int_as_float(vsc_pairs[mid].y) < 0.0f)) lo = mid + 1; else hi = mid; }
    neg_begin[g] = lo;
}

// identity "permutation": every cell keeps its own label (the observed statistic); column 0 of the slab
__global__ void slab_identity_kernel(const uint8_t* __restrict__ labels8s, int32_t N, uint8_t* __restrict__ slab, int32_t PB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) slab[(int64_t)i * PB] = labels8s[i];
}

// exclusive prefix of the per-chunk counts over the chunks (slots) of one region, per (cluster, permutation); and,
// while the running count passes a wanted rank of the cluster, the (slot, permutation group) pair is appended ONCE to
// the work list of the select pass (walk_flag de-duplicates; the list order is arbitrary, every item writes its
// own outputs).  item = (batch * n_slots + slot) * groups + group.
__global__ void tm_prefix_kernel(const float* __restrict__ counts /*[slot][L][PB]*/, const int32_t* __restrict__ region_ptr /*[R+1]*/,
                                 const uint8_t* __restrict__ region_dir /*[R]*/, const int32_t* __restrict__ tm_tgt /*[2][L][8]*/,
                                 int32_t R, int32_t L, int32_t PB, int32_t perms_per_group, int32_t n_valid,
                                 uint32_t item_base /* batch * n_slots * groups */, uint32_t* __restrict__ prefix /*[slot][L][PB]*/,
                                 uint32_t* __restrict__ walk_flag, uint32_t* __restrict__ walk_list, uint32_t* __restrict__ walk_count) {
    const int64_t total = (int64_t)R * L * PB;
    const int groups = PB / perms_per_group;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cp = i % ((int64_t)L * PB);
        const int r = (int)(i / ((int64_t)L * PB));
        const int c = (int)(cp / PB), p = (int)(cp % PB);
        const int32_t* tgt = tm_tgt + ((int)region_dir[r] * L + c) * 8;
        const int m = (p < n_valid) ? __ldg(tgt + 6) : 0;
        int k = 0;
        uint32_t next = m ? (uint32_t)__ldg(tgt) : 0xffffffffu;       // smallest wanted rank above the running count
        uint32_t run = 0;
        const int s0 = region_ptr[r], s1 = region_ptr[r + 1];
        const int group = p / perms_per_group;
        for (int sl = s0; sl < s1; ++sl) {
            const int64_t at = (int64_t)sl * L * PB + cp;
            prefix[at] = run;
            run += (uint32_t)__ldcs(&counts[at]);
            if (run >= next) {                                         // a wanted rank falls inside slot sl
                const uint32_t item = item_base + (uint32_t)sl * (uint32_t)groups + (uint32_t)group;
                if (walk_flag[item] == 0 && atomicExch(&walk_flag[item], 1u) == 0) walk_list[atomicAdd(walk_count, 1u)] = item;
                do { ++k; next = (k < m) ? (uint32_t)__ldg(tgt + k) : 0xffffffffu; } while (run >= next);
            }
        }
    }
}

// One warp = one chunk x 32*Q permutations, same decomposition and shared-memory layout as fast2_means_kernel.
// bins hold a 32-bit state per (cluster, permutation): (entries still to meet before the next wanted rank) << 3 |
// (index k of that rank in the cluster's ascending rank list).  Every entry of the chunk decrements the state of
// its cluster by 8; when the countdown part reaches 0 the entry's value IS the wanted order statistic.
template <int Q>
__device__ __forceinline__ uint32_t tm_label_byte(const typename LabelVec<Q>::type& lw, int q) {
    return (label_word<Q>(lw, q) >> (8 * (q & 3))) & 0xffu;
}

// A countdown reached zero: the value of this entry is the wanted order statistic k of (cluster, permutation);
// store it and re-arm the state with the distance to the cluster's next wanted rank.  Out of line (the caller is
// unrolled 2*B times); the Q states are re-read together, the rare body runs only in the lanes that hit.
template <int Q>
__device__ __noinline__ void tm_hit(unsigned char* bins, const int32_t* tgt_s, float* out_base, int64_t plane_stride, int32_t PB,
                                    uint32_t lane4, uint32_t lw, float v, int dir) {
    uint32_t* b[Q];
    uint32_t st[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) b[q] = reinterpret_cast<uint32_t*>(bins + __byte_perm(lw, lane4, 0x5504u | ((uint32_t)q << 4)) + q * 128);
#pragma unroll
    for (int q = 0; q < Q; ++q) st[q] = *b[q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (st[q] < 8u) {
            const int k = (int)st[q];
            const int c = (int)(((lw >> (8 * q)) & 0xffu) / (Q / 2));
            const int4 t4 = *reinterpret_cast<const int4*>(tgt_s + c * 8 + (k & 4));      // ranks k&4 .. k&4+3
            const int m = tgt_s[c * 8 + 6];
            const int plane = dir ? k : m - 1 - k;
            LRPERM_CHECK(plane >= 0 && plane < kTmPlanes && c >= 0 && c < g_dbg_dims.L, kDbgPlane, plane);
            out_base[(int64_t)plane * plane_stride + (int64_t)c * PB + q] = v;
            const int cur = (k & 3) == 0 ? t4.x : (k & 3) == 1 ? t4.y : (k & 3) == 2 ? t4.z : t4.w;
            const int nxt = (k & 3) == 0 ? t4.y : (k & 3) == 1 ? t4.z : (k & 3) == 2 ? t4.w : tgt_s[c * 8 + 4];
            *b[q] = (k + 1 < m) ? ((uint32_t)(nxt - cur) << 3) | (uint32_t)(k + 1) : kTmNever;
        }
    }
}

template <int Q, int B>
__global__ void __launch_bounds__(32)
tm_select_kernel(const int2* __restrict__ vsc_pairs, const FastChunk* __restrict__ chunks, const int32_t* __restrict__ chunk_of_slot,
                 const uint8_t* __restrict__ slab_all, int32_t PB, int32_t L, int32_t groups_per_batch, int32_t n_slots,
                 const float* __restrict__ counts_all, const uint32_t* __restrict__ prefix_all,
                 const int32_t* __restrict__ tm_tgt /*[2][L][8]: ranks k<6 ascending, [6] = number of ranks*/,
                 float* __restrict__ planes_all /*[6][batch][G][L][PB]*/, int64_t plane_stride,
                 int64_t slab_batch_stride, int64_t slot_batch_stride, int64_t plane_batch_stride, int32_t n_perms_total,
                 const uint32_t* __restrict__ walk_list, const uint32_t* __restrict__ walk_count, uint32_t* __restrict__ cursor,
                 unsigned long long* __restrict__ n_walked) {
    extern __shared__ __align__(16) unsigned char smem2[];
    using LV = typename LabelVec<Q>::type;
    static_assert(Q == 2 || Q == 4, "label byte = cluster * Q/2");
    const int lane = threadIdx.x;
    const int bins_bytes = L * Q * 128;
    unsigned char* bins = smem2;
    int* st_cell = reinterpret_cast<int*>(smem2 + bins_bytes);                 // [2][B]
    float* st_val = reinterpret_cast<float*>(smem2 + bins_bytes + 2 * B * 4);  // [2][B]
    int32_t* tgt_s = reinterpret_cast<int32_t*>(smem2 + bins_bytes + 4 * B * 4);   // [L][8]
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t n_items = *walk_count;
    if (blockIdx.x == 0 && lane == 0) atomicAdd(n_walked, (unsigned long long)n_items);     // diagnostic
    int tgt_dir = -1;

    // Persistent warp: the work list holds only the (batch, chunk, permutation group) items in which a wanted rank
    // falls (a few % of all); a lone walking warp is latency-bound, so every SM keeps as many of them resident as
    // its shared memory allows and they pull items until the list is empty.
    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(cursor, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const uint32_t item = walk_list[it];
        const int group = (int)(item % (uint32_t)groups_per_batch);
        const uint32_t bs = item / (uint32_t)groups_per_batch;
        const int slot = (int)(bs % (uint32_t)n_slots);
        const int batch = (int)(bs / (uint32_t)n_slots);
        const FastChunk ch = chunks[chunk_of_slot[slot]];
        const int dir = (ch.gene >> kTmDirBit) & 1;
        const int gene = ch.gene & ((1 << kTmDirBit) - 1);
        const uint8_t* slab = slab_all + batch * slab_batch_stride;
        const uint32_t* prefix = prefix_all + batch * slot_batch_stride;
        float* planes = planes_all + batch * plane_batch_stride;
        const int n_valid = min(PB, n_perms_total - batch * PB);     // permutations of this batch that exist
        __syncwarp();
        if (dir != tgt_dir) {
            for (int i = lane; i < L * 8; i += 32) tgt_s[i] = tm_tgt[dir * L * 8 + i];
            tgt_dir = dir;
        }
        __syncwarp();

        // ---- initial states from the prefix counts ----
        const int p_first = group * 32 * Q + lane * Q;
        {
            const int64_t base = (int64_t)slot * L * PB + p_first;
            uint32_t* st = reinterpret_cast<uint32_t*>(bins) + lane;
#pragma unroll 2
            for (int c = 0; c < L; ++c) {
                uint32_t pre[Q];
                if (Q == 4) {
                    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(prefix + base + (int64_t)c * PB));
                    pre[0] = a.x; pre[1] = a.y; pre[Q > 2 ? 2 : 0] = a.z; pre[Q > 2 ? 3 : 0] = a.w;
                } else {
                    const uint2 a = __ldcs(reinterpret_cast<const uint2*>(prefix + base + (int64_t)c * PB));
                    pre[0] = a.x; pre[1] = a.y;
                }
                const int m = tgt_s[c * 8 + 6];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    int k = 0;
                    while (k < m && (uint32_t)tgt_s[c * 8 + k] <= pre[q]) ++k;
                    uint32_t state = kTmNever;
                    if (k < m && p_first + q < n_valid) state = (((uint32_t)tgt_s[c * 8 + k] - pre[q]) << 3) | (uint32_t)k;
                    st[(c * Q + q) * 32] = state;
                }
            }
        }
        __syncwarp();

        const LV* __restrict__ slab_w = reinterpret_cast<const LV*>(slab) + (group * 32 + lane);
        const int64_t row_words = PB / Q;
        const int first_cell = __ldg(&vsc_pairs[ch.begin]).x;
        float* const out_base = planes + ((int64_t)gene * L) * PB + p_first;

        auto load_pair = [&](int k) -> int2 {
            return (lane < B && k + lane < ch.end) ? __ldg(&vsc_pairs[k + lane]) : make_int2(first_cell, 0);
        };
        auto stage = [&](int buf, int2 pr) {
            __syncwarp();
            if (lane < B) { st_cell[buf * B + lane] = pr.x; st_val[buf * B + lane] = __int_as_float(pr.y); }
            __syncwarp();
        };
        auto fetch_labels = [&](int buf, LV (&lw)[B]) {
#pragma unroll
            for (int u = 0; u < B; u += 4) {
                const int4 c4 = *reinterpret_cast<const int4*>(st_cell + buf * B + u);
                LRPERM_CHECK((unsigned)c4.x < (unsigned long long)g_dbg_dims.N && (unsigned)c4.y < (unsigned long long)g_dbg_dims.N &&
                             (unsigned)c4.z < (unsigned long long)g_dbg_dims.N && (unsigned)c4.w < (unsigned long long)g_dbg_dims.N, kDbgCell, c4.x);
                lw[u + 0] = __ldg(slab_w + (int64_t)c4.x * row_words);
                lw[u + 1] = __ldg(slab_w + (int64_t)c4.y * row_words);
                lw[u + 2] = __ldg(slab_w + (int64_t)c4.z * row_words);
                lw[u + 3] = __ldg(slab_w + (int64_t)c4.w * row_words);
            }
        };
        // one entry: decrement the Q states; rarely (<= 6 times per cluster and permutation) a countdown hits zero
        auto entry = [&](const LV& lw, float v, uint32_t dec) {
            uint32_t* b[Q];
            uint32_t s[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const uint32_t off = __byte_perm(label_word<Q>(lw, q), lane4, 0x5504u | ((uint32_t)(q & 3) << 4));
                LRPERM_CHECK((long long)off + q * 128 + 4 <= g_dbg_dims.L * Q * 128, kDbgBin, off);
                b[q] = reinterpret_cast<uint32_t*>(bins + off + q * 128);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) s[q] = *b[q];
#pragma unroll
            for (int q = 0; q < Q; ++q) s[q] -= dec;
#pragma unroll
            for (int q = 0; q < Q; ++q) *b[q] = s[q];
            uint32_t mn = s[0];
#pragma unroll
            for (int q = 1; q < Q; ++q) mn = min(mn, s[q]);
            // out of line: this lambda is unrolled 2*B times and the kernel must stay inside the instruction cache
            if (mn < 8u) tm_hit<Q>(bins, tgt_s, out_base, plane_stride, PB, lane4, (uint32_t)label_word<Q>(lw, 0), v, dir);
        };
        auto update = [&](int buf, const LV (&lw)[B], int k0) {
#pragma unroll
            for (int u = 0; u < B; u += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(st_val + buf * B + u);
                entry(lw[u + 0], v4.x, (k0 + u + 0 < ch.end) ? 8u : 0u);
                entry(lw[u + 1], v4.y, (k0 + u + 1 < ch.end) ? 8u : 0u);
                entry(lw[u + 2], v4.z, (k0 + u + 2 < ch.end) ? 8u : 0u);
                entry(lw[u + 3], v4.w, (k0 + u + 3 < ch.end) ? 8u : 0u);
            }
        };

        LV lwA[B], lwB[B];
        int2 pr = load_pair(ch.begin);
        stage(0, pr);
        fetch_labels(0, lwA);
        pr = load_pair(ch.begin + B);
        for (int k = ch.begin; k < ch.end; k += 2 * B) {
            stage(1, pr);
            pr = load_pair(k + 2 * B);
            const bool more1 = k + B < ch.end;
            if (more1) fetch_labels(1, lwB);
            update(0, lwA, k);
            if (!more1) break;
            stage(0, pr);
            pr = load_pair(k + 3 * B);
            if (k + 2 * B < ch.end) fetch_labels(0, lwA);
            update(1, lwB, k + B);
        }
    }
}

// numpy's arithmetic on the selected raw values (see oracle/lrperm_oracle.c `trimean_column`, which this mirrors):
//   null      values fl32(x * r32); neighbour difference in float32; interpolation weights and sums in float64;
//             median = float32 mean of the middle pair
//   observed  values (double)x * r64; everything in float64
struct TmRoles {            // per cluster: plane of q25 lo/hi, median lo/hi, q75 lo/hi; interpolation weights
    int32_t plane[6];
    float g25, g75;         // in {0, .25, .5, .75}: exact
};

template <bool OBSERVED>
__device__ __forceinline__ double tm_lerp(float xa, float xb, double t, float r32, double r64) {
    double a, b, d;
    if (OBSERVED) { a = __dmul_rn((double)xa, r64); b = __dmul_rn((double)xb, r64); d = __dsub_rn(b, a); }
    else { const float fa = __fmul_rn(xa, r32), fb = __fmul_rn(xb, r32); a = fa; b = fb; d = (double)__fsub_rn(fb, fa); }
    return (t >= 0.5) ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));
}

template <bool OBSERVED>
__global__ void tm_finalize_kernel(const float* __restrict__ planes, int64_t plane_stride, const TmRoles* __restrict__ roles,
                                   int32_t G, int32_t L, int32_t PB, float r32, double r64, double* __restrict__ cube, int32_t PP) {
    const int64_t total = (int64_t)G * L * PB;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % PB);
        const int64_t gc = i / PB;
        const int c = (int)(gc % L);
        const TmRoles ro = roles[c];
        float x[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) x[j] = __ldcs(&planes[(int64_t)ro.plane[j] * plane_stride + i]);
        const double q25 = tm_lerp<OBSERVED>(x[0], x[1], (double)ro.g25, r32, r64);
        const double q75 = tm_lerp<OBSERVED>(x[4], x[5], (double)ro.g75, r32, r64);
        double med2;
        if (OBSERVED) med2 = __dmul_rn(2.0, __dmul_rn(__dadd_rn(__dmul_rn((double)x[2], r64), __dmul_rn((double)x[3], r64)), 0.5));
        else med2 = (double)__fmul_rn(2.0f, __fmul_rn(__fadd_rn(__fmul_rn(x[2], r32), __fmul_rn(x[3], r32)), 0.5f));
        cube[gc * PP + p] = __dmul_rn(__dadd_rn(__dadd_rn(q25, med2), q75), 0.25);
    }
}

// CellChat null score + compare + count: prob = a*b / (kh + a*b) >= observed (float64 throughout,
// liana/method/sc/_cellchat.py:7-10, _get_mean_perms.py:133-137).  Same decomposition as score_count_kernel.
__global__ void __launch_bounds__(kScoreThreads)
score_count_cellchat_kernel(const double* __restrict__ cube, int32_t PP, int32_t n_perms_batch, int32_t groups_per_block,
                            const int32_t* __restrict__ rowA, const int32_t* __restrict__ rowB,
                            const double* __restrict__ thr, double kh, int64_t T, uint32_t* __restrict__ counts) {
    __shared__ int32_t s_ra[kScoreTile], s_rb[kScoreTile];
    __shared__ double s_t[kScoreTile];
    __shared__ uint32_t s_c[kScoreTile];
    const int64_t t0 = (int64_t)blockIdx.x * kScoreTile;
    const int nt = (int)min((int64_t)kScoreTile, T - t0);
    for (int i = threadIdx.x; i < kScoreTile; i += blockDim.x) {
        const bool ok = i < nt;
        s_ra[i] = ok ? rowA[t0 + i] : 0;
        s_rb[i] = ok ? rowB[t0 + i] : 0;
        s_t[i] = ok ? thr[t0 + i] : 0.0;
        s_c[i] = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = kScoreThreads / 32;
    const int n_groups = (n_perms_batch + 31) >> 5;
    const int g_first = blockIdx.y * groups_per_block;           // permutation block: see score_count_kernel
    const int n_g = min(groups_per_block, n_groups - g_first);
    const int n_tb = (nt + 31) >> 5;
    for (int item = warp; item < n_g * n_tb; item += NW) {
        const int pg = g_first + item % n_g;
        const int tb = (item / n_g) * 32;
        const int p = pg * 32 + lane;
        const bool valid = p < n_perms_batch;
        {
            uint32_t mine = 0;
            const int tn = min(32, nt - tb);
            for (int u = 0; u < tn; ++u) {
                const int t = tb + u;
                const double a = __ldg(&cube[(int64_t)s_ra[t] * PP + p]);
                const double b = __ldg(&cube[(int64_t)s_rb[t] * PP + p]);
                const double prod = __dmul_rn(a, b);
                const bool ge = valid && (__ddiv_rn(prod, __dadd_rn(kh, prod)) >= s_t[t]);
                const uint32_t msk = __ballot_sync(0xffffffffu, ge);
                if (lane == u) mine = __popc(msk);
            }
            if (lane < tn) atomicAdd(&s_c[tb + lane], mine);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x) if (s_c[i]) atomicAdd(&counts[t0 + i], s_c[i]);
}

// cube[g][c][PP] (perm-minor, float64) -> out[p][c][g] (the reference's layout)
__global__ void cube_export_f64_kernel(const double* __restrict__ cube, int32_t G, int32_t L, int32_t PP,
                                       int32_t n_perms_batch, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int g0 = blockIdx.x * 32, p0 = blockIdx.y * 32, c = blockIdx.z;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int g = g0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (g < G && p < n_perms_batch) ? cube[((int64_t)g * L + c) * PP + p] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, g = g0 + threadIdx.x;
        if (g < G && p < n_perms_batch) out[((int64_t)p * L + c) * G + g] = tile[threadIdx.x][r];
    }
}

}  // namespace lrperm
