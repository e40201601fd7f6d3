// lrperm_kernels.cuh - sm_100a kernels of the permutation-null engine.
//
// Data layout in HBM:
//   csr_pairs[nnz]  int2 {column, float bits}      row-major (CSR) copy of X, 8 B per stored entry
//   csc_pairs[nnz]  int2 {cell,   float bits}      gene-major (CSC) copy, cells ascending per gene
//   order[N]        positions j sorted by cluster (stable), cl_start[L+1]
//   slab[N][PB]     uint8 permuted label of cell i under permutation p (FAST mode, L2 resident)
//   cube[G][L][PP]  float32 per-(gene, cluster) permuted means, permutation-minor (PP = padded batch)
//   counts[T]       uint32 exceedance counters
//
// Reference arithmetic being reproduced: SURVEY.md Appendix B /
// liana/method/_pipe_utils/_get_mean_perms.py:61-87 (cube), :117-142 (p-values),
// liana/method/sc/_cellphonedb.py:25-28, liana/method/sc/_geometric_mean.py:22-23.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lrperm_philox.h"

namespace lrperm {

// ------------------------------------------------------------------------------------------
// Bounds checks of our own (compute-sanitizer is closed on the B200 pool): compiled in only with
// -DLRPERM_BOUNDS_CHECK (`python -m liana_py_b200.build --debug` -> liblrperm_debug.so).  Every index the
// hot kernels compute - shared-memory bin offsets, accumulator columns, cells, chunk ranges, partial-sum slots,
// cube rows - is tested against the extents the host stored in `g_dbg_dims` before the run; a violation sets a
// bit in `g_dbg_flags` (and remembers the first offending value) instead of trapping, `lrperm_debug_bounds`
// returns and clears it.  The production library contains none of this.
// ------------------------------------------------------------------------------------------
struct DebugDims { long long N, G, L, nnz, n_slots, PB, T; };
#ifdef LRPERM_BOUNDS_CHECK
__device__ DebugDims g_dbg_dims;
__device__ unsigned int g_dbg_flags;
__device__ long long g_dbg_first[2];
__device__ __forceinline__ void dbg_fail(unsigned int bit, long long value) {
    if (atomicOr(&g_dbg_flags, bit) == 0u) { g_dbg_first[0] = (long long)bit; g_dbg_first[1] = value; }
}
#define LRPERM_CHECK(cond, bit, value) do { if (!(cond)) dbg_fail((bit), (long long)(value)); } while (0)
#else
#define LRPERM_CHECK(cond, bit, value) do { } while (0)
#endif
// bits: 1 bin offset, 2 cell index, 4 chunk range, 8 partial-sum slot, 16 accumulator column, 32 row extent, 64 cube row, 128 plane / cluster of the trimean select
constexpr unsigned int kDbgBin = 1, kDbgCell = 2, kDbgChunk = 4, kDbgSlot = 8, kDbgColumn = 16, kDbgRow = 32, kDbgCubeRow = 64, kDbgPlane = 128;

constexpr int kExactWarps = 8;       // warps per CTA in the exact-order kernel
constexpr int kFastWarps = 8;        // warps per CTA in the fast kernel
constexpr int kScoreThreads = 256;   // threads per CTA in the score kernel
constexpr int kScoreTile = 256;      // triples per CTA in the score kernel

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ------------------------------------------------------------------------------------------
// CSR packing + validation
// ------------------------------------------------------------------------------------------
// One warp per row.  Besides packing {column, value}, the row's entries are re-ordered so that every
// group of 32 consecutive entries (one warp-wide accumulate in exact_means_kernel) spreads over the
// shared-memory banks (column & 31) as evenly as possible: entries are ranked in bank-sorted order
// and dealt round-robin over the row's 32-entry groups, so a bank with c entries appears at most
// ceil(c / groups) (+1) times per group instead of ~3.4 times for random columns.  The order of the
// stored entries inside a row is irrelevant for the arithmetic (one entry per column and row).
__global__ void pack_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                const float* __restrict__ data, int32_t row0, int32_t row1, int32_t G,
                                int2* __restrict__ pairs, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int r = row0 + warp; r < row1; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        const int n = e - s;
        if (n <= 0) continue;
        // pass 1: entries per bank (lane b holds the count of bank b)
        int cnt = 0;
        for (int o = s; o < e; o += 32) {
            const bool valid = o + lane < e;
            int c = valid ? indices[o + lane] : 0;
            if (valid && (c < 0 || c >= G)) { atomicOr(err, 1); c = 0; }
            const int bank = c & 31;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const unsigned m = __ballot_sync(0xffffffffu, valid && bank == b);
                if (lane == b) cnt += __popc(m);
            }
        }
        // exclusive prefix over banks -> first bank-sorted index of every bank
        int start = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, start, d);
            if (lane >= d) start += t;
        }
        start -= cnt;
        const int nfull = n >> 5, rem = n & 31;
        const int ngroups = nfull + (rem ? 1 : 0);
        const int phase1 = rem * ngroups;       // entries dealt while the partial (last) group fills up
        // pass 2: place
        int run = 0;                            // entries of bank `lane` seen in earlier chunks
        for (int o = s; o < e; o += 32) {
            const bool valid = o + lane < e;
            int c = valid ? indices[o + lane] : 0;
            if (c < 0 || c >= G) c = 0;
            const float v = valid ? data[o + lane] : 0.0f;
            const int bank = c & 31;
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? bank : (32 + lane));
            const int rank = __shfl_sync(0xffffffffu, run, bank) + __popc(peers & lt_mask);
            const int i = __shfl_sync(0xffffffffu, start, bank) + rank;      // bank-sorted index
            int grp, col;
            if (i < phase1) { grp = i % ngroups; col = i / ngroups; }
            else { const int j = i - phase1; grp = j % nfull; col = rem + j / nfull; }
            if (valid) pairs[s + grp * 32 + col] = make_int2(c, __float_as_int(v));
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const unsigned m = __ballot_sync(0xffffffffu, valid && bank == b);
                if (lane == b) run += __popc(m);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Input contract on device (replaces the matrix work of prep_check_adata,
// liana/method/_pipe_utils/_pre.py:112-142,168, and the gene subset of liana_pipe,
// liana/method/sc/_liana_pipe.py:161-164): the caller's FULL matrix is staged once, statistics are
// taken over it, then it is filtered to kept rows / mapped to the resource's gene columns.
// ------------------------------------------------------------------------------------------
// One warp per row of the staged matrix: per-column sums over ALL rows (float64 atomics; used only for
// the "gene is empty" test and the matrix mean), number of non-finite values, number of empty rows,
// total over the kept rows, and the number of stored entries whose column is not greater than its predecessor's in
// the same row (0 = every row strictly increasing = SciPy's canonical format: sorted, no duplicates - what the host
// would otherwise establish with a pass of its own over every index, `has_canonical_format`).
__global__ void stage_stats_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const float* __restrict__ data, const uint8_t* __restrict__ row_keep, int32_t N, int32_t G,
                                   double* __restrict__ col_sums, double* __restrict__ kept_total,
                                   unsigned long long* __restrict__ n_nonfinite, unsigned long long* __restrict__ n_empty_rows,
                                   unsigned long long* __restrict__ n_unordered, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double block_total = 0.0;
    unsigned long long bad = 0, empty = 0, unordered = 0;
    for (int r = warp; r < N; r += nwarps) {
        const int64_t s = indptr[r], e = indptr[r + 1];
        double row = 0.0;
        for (int64_t o = s + lane; o < e; o += 32) {
            const int c = indices[o];
            const float v = data[o];
            if (o > s && indices[o - 1] >= c) ++unordered;      // the neighbour's load: same or previous cache line
            if (c < 0 || c >= G) { atomicOr(err, 1); continue; }
            if (!isfinite(v)) { ++bad; continue; }
            if (v != 0.0f) atomicAdd(&col_sums[c], (double)v);
            row += (double)v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) row += __shfl_xor_sync(0xffffffffu, row, d);
        if (lane == 0) {
            if (row == 0.0) ++empty;
            if (!row_keep || row_keep[r]) block_total += row;
        }
    }
    if (bad) atomicAdd(n_nonfinite, bad);
    if (unordered) atomicAdd(n_unordered, unordered);
    if (lane == 0) {
        if (empty) atomicAdd(n_empty_rows, empty);
        if (block_total != 0.0) atomicAdd(kept_total, block_total);
    }
}

// Dense row-major input (adata.X as a 2-D array) -> the staged CSR the other kernels read; `csr_matrix(X)` of the
// reference's `_choose_mtx_rep` (liana/method/_pipe_utils/_pre.py:259-262) keeps the entries that compare != 0 (NaN
// included, -0.0 not) in column order.  HBM-bound streaming passes, one warp per row, 128 B coalesced loads:
// pass 1 counts a row's stored entries (-> counts[r + 1], 64-bit total), pass 2 compacts them behind indptr[r] with a
// ballot prefix so the column order is kept.  T = float or double (rounded to float32 like `astype(np.float32)`).
template <typename T>
__global__ void dense_count_kernel(const T* __restrict__ X, int64_t ld, int32_t N, int32_t G, int64_t* __restrict__ counts,
                                   unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long mine = 0;
    for (int r = warp; r < N; r += nwarps) {
        const T* row = X + (int64_t)r * ld;
        int n = 0;
        for (int c0 = lane; c0 < G; c0 += 128) {           // 4 independent 128 B loads in flight per warp
            T t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = (c0 + 32 * k < G) ? row[c0 + 32 * k] : (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) n += (t[k] != (T)0) ? 1 : 0;   // tested in T: a double that underflows float32 stays stored
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
        if (lane == 0) { counts[r + 1] = n; mine += (unsigned long long)n; }
    }
    if (lane == 0 && mine) atomicAdd(total, mine);
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
}

template <typename T>
__global__ void dense_fill_kernel(const T* __restrict__ X, int64_t ld, int32_t N, int32_t G, const int64_t* __restrict__ indptr,
                                  int32_t* __restrict__ indices, float* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t below = (1u << lane) - 1u;
    for (int r = warp; r < N; r += nwarps) {
        const T* row = X + (int64_t)r * ld;
        int64_t base = indptr[r];
        for (int c0 = 0; c0 < G; c0 += 128) {              // 4 independent 128 B loads in flight per warp
            T t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = (c0 + 32 * k + lane < G) ? row[c0 + 32 * k + lane] : (T)0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool keep = t[k] != (T)0;
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int64_t o = base + __popc(m & below);
                    indices[o] = c0 + 32 * k + lane;
                    data[o] = (float)t[k];
                }
                base += __popc(m);
            }
        }
    }
}

// largest stored value over the kept rows of the staged matrix (CellChat's `mat_max = adata.X.max()`,
// liana/method/sc/_liana_pipe.py:143-146), as an order-preserving unsigned key so that atomicMax applies
__device__ __forceinline__ uint32_t float_order_key(float v) {
    const uint32_t b = (uint32_t)__float_as_int(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void stage_max_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ data,
                                 const uint8_t* __restrict__ row_keep, int32_t N, uint32_t* __restrict__ max_key,
                                 unsigned long long* __restrict__ n_stored) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t best = 0;
    unsigned long long stored = 0;
    for (int r = warp; r < N; r += nwarps) {
        if (row_keep && !row_keep[r]) continue;
        const int64_t s = indptr[r], e = indptr[r + 1];
        for (int64_t o = s + lane; o < e; o += 32) best = max(best, float_order_key(data[o]));
        if (lane == 0) stored += (unsigned long long)(e - s);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0) {
        if (best) atomicMax(max_key, best);
        if (stored) atomicAdd(n_stored, stored);
    }
}

// packed {column, value bits} pairs -> separate index / value arrays (lrperm_export_matrix_csr)
__global__ void unpack_pairs_kernel(const int2* __restrict__ pairs, int64_t n, int32_t* __restrict__ indices, float* __restrict__ data) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 pr = pairs[i];
        indices[i] = pr.x;
        data[i] = __int_as_float(pr.y);
    }
}

// 16-bit column indices of a narrowed stage upload -> int32
__global__ void widen_u16_kernel(const uint16_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

// column range check of raw CSR indices (the packing kernel does it too, but FAST-only flows never pack)
__global__ void validate_columns_kernel(const int32_t* __restrict__ indices, int64_t nnz, int32_t G, int* __restrict__ err) {
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = indices[i];
        bad |= c < 0 || c >= G;
    }
    if (bad) atomicOr(err, 1);
}

// Per-cluster sum x and sum x^2 over ALL stored entries of the staged matrix (every gene, before the resource subset):
// the cluster mean / standard deviation of scSeqComm's `_cluster_stats` (liana/method/sc/_liana_pipe.py:466-476).
// row_label[r] = cluster of staged row r, -1 = row not kept.  float64 and DETERMINISTIC: one warp per CTA walks its
// rows in a fixed order into private shared accumulators (no atomics), the per-CTA partials are summed in CTA order.
__global__ void __launch_bounds__(32)
staged_cluster_stats_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ data,
                            const int32_t* __restrict__ row_label, int32_t N, int32_t L,
                            double* __restrict__ partial /*[gridDim.x][2 L]*/) {
    extern __shared__ double cs_smem[];          // [2][L]
    const int lane = threadIdx.x;
    for (int c = lane; c < 2 * L; c += 32) cs_smem[c] = 0.0;
    __syncwarp();
    for (int r = blockIdx.x; r < N; r += gridDim.x) {
        const int c = row_label[r];
        if (c < 0 || c >= L) continue;
        const int64_t s = indptr[r], e = indptr[r + 1];
        double a = 0.0, b = 0.0;
        for (int64_t o = s + lane; o < e; o += 32) { const double x = (double)data[o]; a += x; b += x * x; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
        if (lane == 0) { cs_smem[c] += a; cs_smem[L + c] += b; }
        __syncwarp();
    }
    __syncwarp();
    for (int c = lane; c < 2 * L; c += 32) partial[(int64_t)blockIdx.x * 2 * L + c] = cs_smem[c];
}

__global__ void staged_cluster_stats_reduce_kernel(const double* __restrict__ partial, int32_t n_blocks, int32_t L2,
                                                   double* __restrict__ out /*[2 L]*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L2) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += partial[(int64_t)b * L2 + c];
    out[c] = s;
}

// entries of every KEPT row that survive the column map -> counts[new_row]
__global__ void mapped_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const int32_t* __restrict__ col_map, const int32_t* __restrict__ new_row /*[N], -1 = dropped*/,
                                    int32_t N, int32_t* __restrict__ counts, unsigned long long* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long mine = 0;
    for (int r = warp; r < N; r += nwarps) {
        const int nr = new_row[r];
        if (nr < 0) continue;
        const int64_t s = indptr[r], e = indptr[r + 1];
        int n = 0;
        for (int64_t o = s + lane; o < e; o += 32) n += col_map[indices[o]] >= 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
        if (lane == 0) { counts[nr] = n; mine += (unsigned long long)n; }
    }
    if (lane == 0 && mine) atomicAdd(total, mine);       // the committed matrix must stay below 2^31 entries
}

// order-preserving compaction of the kept entries into (out_indices, out_data) at out_indptr[new_row]
__global__ void mapped_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                      const float* __restrict__ data, const int32_t* __restrict__ col_map,
                                      const int32_t* __restrict__ new_row, int32_t N, const int32_t* __restrict__ out_indptr,
                                      int32_t* __restrict__ out_indices, float* __restrict__ out_data) {
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int r = warp; r < N; r += nwarps) {
        const int nr = new_row[r];
        if (nr < 0) continue;
        const int64_t s = indptr[r], e = indptr[r + 1];
        int w = out_indptr[nr];
        for (int64_t o = s; o < e; o += 32) {
            const bool in = o + lane < e;
            const int c = in ? col_map[indices[o + lane]] : -1;
            const unsigned m = __ballot_sync(0xffffffffu, c >= 0);
            if (c >= 0) {
                const int dst = w + __popc(m & lt_mask);
                out_indices[dst] = c;
                out_data[dst] = data[o + lane];
            }
            w += __popc(m);
        }
    }
}

__global__ void row_keep_to_flag_kernel(const uint8_t* __restrict__ row_keep, int32_t N, int32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flag[i] = row_keep ? (row_keep[i] != 0) : 1;
}

__global__ void new_row_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ excl, int32_t N, int32_t* __restrict__ new_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) new_row[i] = flag[i] ? excl[i] : -1;
}

template <typename T>
__global__ void narrow_indptr_kernel(const T* __restrict__ in, int64_t n, int32_t* __restrict__ out,
                                     int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        T v = in[i];
        if (v < 0 || (int64_t)v > 0x7fffffffLL) atomicOr(err, 2);
        if (i > 0 && in[i - 1] > v) atomicOr(err, 4);
        out[i] = (int32_t)v;
    }
}

// Row pointers of the STAGED (caller's full) matrix: validated and kept as int64 - the full matrix of an atlas-sized
// input exceeds 2^31 stored entries long before its resource-gene subset (the engine matrix, int32 offsets) does.
template <typename T>
__global__ void widen_indptr_kernel(const T* __restrict__ in, int64_t n, int64_t* __restrict__ out, int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const T v = in[i];
        if (v < 0) atomicOr(err, 2);
        if (i > 0 && in[i - 1] > v) atomicOr(err, 4);
        out[i] = (int64_t)v;
    }
}

// CSR -> CSC transpose of ONE cell tile (rows [row0, row1)), step 1: per stored entry the sort key (column) and the
// payload {cell, value bits} packed so that the sorted payload array IS the tile's csc_pairs range; column histogram of
// the tile privatised in shared memory (one global atomic per (CTA, column) instead of one per entry).  The entries
// of a tile's rows are contiguous in CSR order, and the tile's CSC occupies the same index range.
// `pairs` == nullptr: read the raw CSR arrays (raw_idx / raw_val) instead of the packed pairs and validate the column
// range here (flag `err`, column clamped to 0): the packed copy is only needed by the row-gather kernels and is then
// built lazily.
__global__ void csc_keys_kernel(const int32_t* __restrict__ indptr, const int2* __restrict__ pairs, int32_t row0, int32_t row1, int32_t G,
                                uint32_t* __restrict__ keys, unsigned long long* __restrict__ payload,
                                int32_t* __restrict__ col_count, int use_shared,
                                const int32_t* __restrict__ raw_idx = nullptr, const float* __restrict__ raw_val = nullptr,
                                int* __restrict__ err = nullptr) {
    extern __shared__ int32_t hist_smem[];
    int32_t* hist = use_shared ? hist_smem : col_count;      // G too large for shared memory: global atomics
    if (use_shared) for (int g = threadIdx.x; g < G; g += blockDim.x) hist[g] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = row0 + warp; r < row1; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        for (int o = s + lane; o < e; o += 32) {
            int2 pr;
            if (pairs) pr = pairs[o];
            else {
                pr.x = raw_idx[o];
                pr.y = __float_as_int(raw_val[o]);
                if (pr.x < 0 || pr.x >= G) { atomicOr(err, 1); pr.x = 0; }
            }
            keys[o] = (uint32_t)pr.x;
            payload[o] = ((unsigned long long)(uint32_t)pr.y << 32) | (uint32_t)r;
            atomicAdd(&hist[pr.x], 1);
        }
    }
    __syncthreads();
    if (use_shared)
        for (int g = threadIdx.x; g < G; g += blockDim.x)
            if (hist[g]) atomicAdd(&col_count[g], hist[g]);
}

// small host -> device table copy that does NOT go through the copy engine: `src` is pinned host memory read by the
// SMs over PCIe.  The copy engine serves H2D copies of all streams in submission order, so a few-KB table queued
// behind hundreds of MB of matrix tiles would arrive only after them (see chunks_append_tile).
__global__ void copy_words_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// the same for any byte count (4-byte aligned pointers): whole words, then the tail bytes
__global__ void copy_bytes_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int64_t n) {
    const int64_t words = n >> 2;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) d4[i] = s4[i];
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(words << 2) + threadIdx.x] = src[(words << 2) + threadIdx.x];
}

// stored-entry counts per (cluster, gene): props = count / n_c  (_common.py:41-42)
__global__ void stored_counts_kernel(const int32_t* __restrict__ indptr, const int2* __restrict__ pairs,
                                     const int32_t* __restrict__ labels, int32_t N, int32_t G,
                                     int32_t* __restrict__ counts /*[L*G]*/) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < N; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        const int c = labels[r];
        for (int o = s + lane_id(); o < e; o += 32) atomicAdd(&counts[(int64_t)c * G + pairs[o].x], 1);
    }
}

// Per-(cluster, gene) float64 moments of the observed matrix, for the non-permutation scores that
// rank_aggregate combines with the permutation p-values (liana/method/sc/_liana_pipe.py:259-273,
// 341-360): sum x (cluster means, z-scores), sum x^2 (sc.pp.scale's per-gene variance) and
// sum (base^x - 1) (the 1-vs-rest log2FC).  One warp per (segment of a cluster's cells, gene tile);
// the warp owns its shared-memory accumulators (lanes cover one row's distinct columns), segment
// partials are combined in fixed order by moments_reduce_kernel -> deterministic.
constexpr int kMomentTile = 4096;     // genes per tile: 3 x 4096 x 8 B = 96 KB of shared memory

__global__ void __launch_bounds__(32)
cluster_moments_kernel(const int32_t* __restrict__ indptr, const int2* __restrict__ pairs,
                       const int32_t* __restrict__ order, const int32_t* __restrict__ cl_start,
                       int32_t G, int32_t n_seg, double ln_base, double* __restrict__ partial /*[3][L][n_seg][G]*/,
                       int32_t L) {
    extern __shared__ double macc[];
    const int lane = threadIdx.x, seg = blockIdx.x, c = blockIdx.y, g0 = blockIdx.z * kMomentTile;
    const int gt = min(kMomentTile, G - g0);
    double* s1 = macc; double* s2 = macc + kMomentTile; double* s3 = macc + 2 * kMomentTile;
    for (int g = lane; g < gt; g += 32) { s1[g] = 0.0; s2[g] = 0.0; s3[g] = 0.0; }
    __syncwarp();
    const int kb = cl_start[c], ke = cl_start[c + 1];
    const int64_t len = ke - kb;
    const int k0 = kb + (int)(len * seg / n_seg), k1 = kb + (int)(len * (seg + 1) / n_seg);
    for (int k = k0; k < k1; ++k) {
        const int cell = order[k];
        const int rs = indptr[cell], re = indptr[cell + 1];
        for (int o = rs + lane; o < re; o += 32) {
            const int2 pr = pairs[o];
            const int g = pr.x - g0;
            if (g >= 0 && g < gt) {
                const double x = (double)__int_as_float(pr.y);
                s1[g] += x;
                s2[g] += x * x;
                s3[g] += expm1(x * ln_base);
            }
        }
        __syncwarp();
    }
    const size_t plane = (size_t)L * n_seg * G;
    double* out = partial + ((size_t)c * n_seg + seg) * G + g0;
    for (int g = lane; g < gt; g += 32) { out[g] = s1[g]; out[plane + g] = s2[g]; out[2 * plane + g] = s3[g]; }
}

__global__ void moments_reduce_kernel(const double* __restrict__ partial, int32_t G, int32_t L, int32_t n_seg,
                                      double* __restrict__ out /*[3][L][G]*/) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)3 * L * G;
    if (i >= total) return;
    const int g = (int)(i % G);
    const int64_t mc = i / G;                 // moment * L + cluster
    double s = 0.0;
    for (int seg = 0; seg < n_seg; ++seg) s += partial[(mc * n_seg + seg) * G + g];
    out[i] = s;
}

// ------------------------------------------------------------------------------------------
// Label tables (replaces the host passes of lrperm_set_labels): cluster sizes, first out-of-range cell, positions
// grouped by cluster (stable: ascending inside a cluster = the reference's labels_mask order,
// _get_mean_perms.py:47-52 + boolean row selection at :63), their inverse, byte labels.
// ------------------------------------------------------------------------------------------
__global__ void label_hist_kernel(const int32_t* __restrict__ labels, int32_t N, int32_t L, int32_t* __restrict__ counts,
                                  int32_t* __restrict__ first_bad, uint32_t* __restrict__ keys, int32_t* __restrict__ iota) {
    extern __shared__ int32_t lh_smem[];
    const bool use_shared = L <= 4096;
    if (use_shared) for (int c = threadIdx.x; c < L; c += blockDim.x) lh_smem[c] = 0;
    __syncthreads();
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const int c = labels[j];
        iota[j] = j;
        if (c < 0 || c >= L) { atomicMin(first_bad, j); keys[j] = 0; continue; }
        keys[j] = (uint32_t)c;
        atomicAdd(use_shared ? &lh_smem[c] : &counts[c], 1);
    }
    __syncthreads();
    if (use_shared) for (int c = threadIdx.x; c < L; c += blockDim.x) if (lh_smem[c]) atomicAdd(&counts[c], lh_smem[c]);
}

__global__ void label_tables_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ order, int32_t N,
                                    int32_t* __restrict__ rank, uint8_t* __restrict__ labels8) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) {
        rank[order[k]] = k;
        labels8[k] = (uint8_t)(labels[k] & 0xff);
    }
}

// ------------------------------------------------------------------------------------------
// Permutation sources
// ------------------------------------------------------------------------------------------
struct PermSource {
    const void* idx;      // INDEX mode: perm_idx of this batch, [pb][N]
    int is64;
    int philox;           // 1 = PHILOX
    uint64_t seed;
    int64_t perm_id0;     // global id of the first permutation of the batch
    int identity;         // 1 = identity permutation (observed statistics)
};

__device__ __forceinline__ int32_t perm_lookup(const PermSource& ps, int p_local, int64_t N, int32_t j) {
    if (ps.is64) return (int32_t)((const int64_t*)ps.idx)[(int64_t)p_local * N + j];
    return ((const int32_t*)ps.idx)[(int64_t)p_local * N + j];
}

__device__ __forceinline__ int2 ld_stream_int2(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------
// EXACT mode: one warp per (permutation, cluster).  The warp walks the cluster's positions in
// ascending order (the reference's order), fetches the CSR row of the cell the permutation put
// there, and adds fl32(x * inv_n) into a G-float accumulator row in shared memory with separate
// multiply and add roundings - bit-identical to scipy's (X * (1/n)).sum(0) on the permuted rows.
// Lanes are spread over the stored entries of one row (distinct columns -> no intra-row hazard).
// ------------------------------------------------------------------------------------------
// (min CTAs per SM so that ~28 warps stay resident: the register allocator must not trade occupancy -
// rows in flight - for a few registers)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (28 / WARPS) > 0 ? (28 / WARPS) : 1)
exact_means_kernel(const int32_t* __restrict__ indptr, const int2* __restrict__ pairs,
                   const int32_t* __restrict__ order, const int32_t* __restrict__ cl_start,
                   const float* __restrict__ inv_n, PermSource ps,
                   int32_t N, int32_t G, int32_t L, int32_t n_perms_batch,
                   float* __restrict__ cube, int32_t PP) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // items ordered cluster-major so the warps of one CTA have equal work
    const int64_t item = (int64_t)blockIdx.x * WARPS + warp;
    if (item >= (int64_t)n_perms_batch * L) return;
    const int c = (int)(item / n_perms_batch);
    const int p = (int)(item - (int64_t)c * n_perms_batch);
    // accumulator row of the warp: G columns (padded to 32) + 32 dummy slots, one per lane, that absorb the
    // "+ 0.0f" of lanes past the end of a row, so the accumulate step has no divergent branch
    const int Gpad = ((G + 31) & ~31) + 32;
    float* acc = smem + (size_t)warp * Gpad;
    for (int g = lane; g < Gpad; g += 32) acc[g] = 0.0f;
    __syncwarp();
    const int dummy_col = Gpad - 32 + lane;

    const float inv = inv_n[c];
    const int k_begin = cl_start[c], k_end = cl_start[c + 1];
    lrperm_bij_t bij;
    if (ps.philox) bij = lrperm_bij_init(ps.seed, (uint64_t)(ps.perm_id0 + p), (uint32_t)N);

    auto fetch_meta = [&](int k, int& rs, int& re) {
        rs = 0; re = 0;
        if (k < k_end) {
            int cell;
            // PHILOX: the bijection is defined on cluster-sorted positions (position order[k] holds cell
            // rho_p(k)), so neither `order` nor a label gather is needed; see slab_philox_kernel
            if (ps.philox) cell = (int)lrperm_bij_forward(&bij, (uint32_t)k);
            else {
                const int j = order[k];
                cell = ps.identity ? j : perm_lookup(ps, p, N, j);
            }
            LRPERM_CHECK((unsigned)cell < (unsigned long long)g_dbg_dims.N, kDbgCell, cell);
            rs = indptr[cell];
            re = indptr[cell + 1];
            LRPERM_CHECK(rs >= 0 && rs <= re && re <= g_dbg_dims.nnz, kDbgRow, rs);
        }
    };

    // Row pipeline: the first U*32 stored entries of row r+1 are loaded into registers while row r is
    // being accumulated, so every warp keeps ~1.3 KB of gathered row data in flight (HBM latency).
    constexpr int U = 6;
    auto prefetch = [&](int s, int e, int2 (&q)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int o = s + lane + 32 * u;
            q[u] = (o < e) ? ld_stream_int2(pairs + o) : make_int2(dummy_col, 0);
        }
    };
    // Accumulate one row.  The columns of a row are distinct, so its U groups of 32 entries are independent:
    // all loads first, then the adds, then the stores - one shared-memory round trip per row, not per group.
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc);
    auto lds = [](uint32_t a) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); return x; };
    auto sts = [](uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); };
    auto apply = [&](int s, int e, const int2 (&q)[U]) {
        const int n = e - s;                                   // warp-uniform
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (u * 32 < n) {
                LRPERM_CHECK((unsigned)q[u].x < (unsigned)Gpad, kDbgColumn, q[u].x);
                x[u] = lds(acc_s + (uint32_t)q[u].x * 4u);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (u * 32 < n) x[u] = __fadd_rn(x[u], __fmul_rn(__int_as_float(q[u].y), inv));
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (u * 32 < n) sts(acc_s + (uint32_t)q[u].x * 4u, x[u]);
        for (int o = s + 32 * U; o < e; o += 128) {            // rows longer than U*32 entries (uniform loop)
            int2 t[4];
            float y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = (o + lane + 32 * u < e) ? ld_stream_int2(pairs + o + lane + 32 * u) : make_int2(dummy_col, 0);
#pragma unroll
            for (int u = 0; u < 4; ++u) y[u] = lds(acc_s + (uint32_t)t[u].x * 4u);
#pragma unroll
            for (int u = 0; u < 4; ++u) sts(acc_s + (uint32_t)t[u].x * 4u, __fadd_rn(y[u], __fmul_rn(__int_as_float(t[u].y), inv)));
        }
        __syncwarp();   // the next row may touch the same columns from other lanes
    };

    int rs_n, re_n;
    fetch_meta(k_begin + lane, rs_n, re_n);
    int2 qa[U], qb[U];      // ping-pong row buffers (no register moves between rows)
    int sa = __shfl_sync(0xffffffffu, rs_n, 0), ea = __shfl_sync(0xffffffffu, re_n, 0), sb, eb;
    prefetch(sa, ea, qa);
    for (int k0 = k_begin; k0 < k_end; k0 += 32) {
        const int rs_c = rs_n, re_c = re_n;
        fetch_meta(k0 + 32 + lane, rs_n, re_n);   // extents of the next 32 rows ((0,0) past the end)
        const int nrows = min(32, k_end - k0);
        for (int r = 0; r < nrows; r += 2) {
            // row r is in qa; fetch row r+1 into qb, then row r+2 into qa
            sb = __shfl_sync(0xffffffffu, rs_c, (r + 1) & 31);
            eb = __shfl_sync(0xffffffffu, re_c, (r + 1) & 31);
            if (r + 1 >= nrows) { sb = 0; eb = 0; }          // nrows is odd only in the last block
            prefetch(sb, eb, qb);
            apply(sa, ea, qa);
            if (r + 2 < 32) {
                sa = __shfl_sync(0xffffffffu, rs_c, (r + 2) & 31);
                ea = __shfl_sync(0xffffffffu, re_c, (r + 2) & 31);
                if (r + 2 >= nrows) { sa = 0; ea = 0; }
            } else {
                sa = __shfl_sync(0xffffffffu, rs_n, 0);
                ea = __shfl_sync(0xffffffffu, re_n, 0);
            }
            prefetch(sa, ea, qa);
            apply(sb, eb, qb);
        }
    }
    // flush: cube[g][c][p]
    for (int g = lane; g < G; g += 32) cube[((int64_t)g * L + c) * PP + p] = acc[g];
}

// ------------------------------------------------------------------------------------------
// FAST mode, step 1: label slab.  slab[i*PB + p] = label of the position cell i lands on.
// ------------------------------------------------------------------------------------------
__global__ void slab_from_index_kernel(PermSource ps, const uint8_t* __restrict__ labels8, int32_t N,
                                       int32_t n_perms_batch, uint8_t* __restrict__ slab, int32_t PB) {
    // thread per (p, j): reads of perm_idx coalesce along j, byte writes scatter into the slab
    const int p = blockIdx.y;
    if (p >= n_perms_batch) return;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const int cell = perm_lookup(ps, p, N, j);
        slab[(int64_t)cell * PB + p] = labels8[j];
    }
}

__global__ void slab_philox_kernel(PermSource ps, const int32_t* __restrict__ cl_start, const uint8_t* __restrict__ bucket_cluster,
                                   int32_t n_buckets, int32_t bucket_shift, int32_t L, int32_t label_scale, int32_t N,
                                   int32_t n_perms_batch, uint8_t* __restrict__ slab, int32_t PB, int32_t cells_per_block,
                                   int32_t cell0 = 0, int32_t cell1 = 0x7fffffff /* rows [cell0, cell1) of the slab only */) {
    // thread = one permutation of the batch (key derived once); loop over a tile of cells.
    // PHILOX permutations are defined on CLUSTER-SORTED positions: pi_p(order[k]) = rho_p(k).  The label of
    // cell i is therefore the cluster whose range [cl_start[c], cl_start[c+1]) holds k = rho_p^-1(i): a
    // lookup in two small shared-memory tables instead of a random gather from the N-entry label array
    // (32 L1 wavefronts per warp, which also competed with the shared-memory-bound main kernel it overlaps).
    extern __shared__ int32_t s_tab[];
    int32_t* s_start = s_tab;                                             // [L + 1]
    uint8_t* s_bucket = reinterpret_cast<uint8_t*>(s_tab + L + 1);        // [n_buckets]
    for (int i = threadIdx.x; i <= L; i += blockDim.x) s_start[i] = cl_start[i];
    for (int i = threadIdx.x; i < n_buckets; i += blockDim.x) s_bucket[i] = bucket_cluster[i];
    __syncthreads();
    const int p = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = p < n_perms_batch;
    lrperm_bij_t bij = lrperm_bij_init(ps.seed, (uint64_t)(ps.perm_id0 + (active ? p : 0)), (uint32_t)N);
    if (!active) return;
    // blocks of cells_per_block cells, strided over gridDim.x: the grid may be as large as the number of blocks (one block
    // each) or a small PERSISTENT one.  The pipelined FAST path uses the latter: this kernel is ALU-bound and runs next to the
    // shared-memory-bound main kernel on a high-priority stream; launched with one CTA per block it floods every SM with up to
    // 56 always-ready warps against the main kernel's 8, and the main kernel loses 18 % (scripts/overlap_ab.py); a few
    // resident warps per scheduler keep the integer pipe just as busy.
    const int end = min(N, cell1);
    for (int64_t i0 = (int64_t)cell0 + (int64_t)blockIdx.x * cells_per_block; i0 < end; i0 += (int64_t)gridDim.x * cells_per_block) {
        const int i1 = (int)min((int64_t)end, i0 + cells_per_block);
        for (int i = (int)i0; i < i1; ++i) {
            const int k = (int)lrperm_bij_inverse(&bij, (uint32_t)i);
            int c = s_bucket[k >> bucket_shift];                          // cluster of the bucket's first position
            while (k >= s_start[c + 1]) ++c;                              // a bucket rarely spans a boundary
            slab[(int64_t)i * PB + p] = (uint8_t)(c * label_scale);
        }
    }
}

__global__ void philox_perms_kernel(uint64_t seed, int64_t perm_id0, int32_t N, int32_t n_perms,
                                    const int32_t* __restrict__ rank, int32_t* __restrict__ out) {
    // out[p][j] = cell placed at position j = rho_p(rank[j]), rank[j] = index of j in cluster-sorted order
    const int p = blockIdx.y;
    if (p >= n_perms) return;
    lrperm_bij_t bij = lrperm_bij_init(seed, (uint64_t)(perm_id0 + p), (uint32_t)N);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x)
        out[(int64_t)p * N + j] = (int32_t)lrperm_bij_forward(&bij, (uint32_t)rank[j]);
}

// slab label bytes for the v2 fast kernel: label * Q/2 (so that byte << 8 == label * Q * 128)
__global__ void scale_labels_kernel(const uint8_t* __restrict__ labels8, int32_t N, int32_t scale, uint8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = (uint8_t)(labels8[i] * scale);
}

// ------------------------------------------------------------------------------------------
// FAST mode, step 2: per-gene private histograms.
// One warp = one chunk of one gene's stored entries x 32*Q permutations (lane owns Q of them).
// Every lane keeps its own L x Q float bins in shared memory, laid out [c][q][lane] so that a
// warp-wide access always hits 32 distinct banks: no atomics, no conflicts, deterministic order
// (ascending cell within the gene).  The only per-entry global traffic is Q label bytes per lane
// from the L2-resident slab; the {cell, value} pair is a warp-uniform (broadcast) load.
// ------------------------------------------------------------------------------------------
template <int Q> struct LabelWord;
template <> struct LabelWord<1> { using type = uint8_t; };
template <> struct LabelWord<2> { using type = uint16_t; };
template <> struct LabelWord<4> { using type = uint32_t; };

struct FastChunk { int32_t gene, begin, end, slot; };   // slot = index of the partial-sum row block

template <int Q>
__device__ __forceinline__ void fast_update(float* mybins, uint32_t lw, float v) {
    float* b[Q];
    float x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) b[q] = mybins + (((lw >> (8 * q)) & 0xffu) * Q + q) * 32;
#pragma unroll
    for (int q = 0; q < Q; ++q) x[q] = *b[q];
#pragma unroll
    for (int q = 0; q < Q; ++q) *b[q] = x[q] + v;
}

template <int Q, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
fast_means_kernel(const int2* __restrict__ csc_pairs, const FastChunk* __restrict__ chunks, int32_t n_chunks,
                  const uint8_t* __restrict__ slab, int32_t PB, int32_t L, int32_t groups_per_batch,
                  float* __restrict__ partials /*[slot][L][PB]*/) {
    extern __shared__ float smem[];
    using LW = typename LabelWord<Q>::type;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * WARPS + warp;
    if (item >= (int64_t)n_chunks * groups_per_batch) return;
    // consecutive warps share the chunk (different permutation groups) -> pair loads hit L1
    const int chunk_id = (int)(item / groups_per_batch);
    const int group = (int)(item - (int64_t)chunk_id * groups_per_batch);
    const FastChunk ch = chunks[chunk_id];
    float* bins = smem + (size_t)warp * (L * Q * 32);
    for (int i = lane; i < L * Q * 32; i += 32) bins[i] = 0.0f;
    __syncwarp();

    const LW* __restrict__ slab_w = reinterpret_cast<const LW*>(slab) + (group * 32 + lane);
    const int64_t row_words = PB / Q;              // label words per slab row
    float* mybins = bins + lane;

    // Software pipeline over blocks of 32 stored entries: lane l holds entry k+l; cell ids are
    // broadcast by shuffle so that the 32 label-word loads of the NEXT block are all in flight
    // (L2 latency) while the current block's 32 x Q bin updates run out of shared memory.
    constexpr int B = 32;
    const int first_cell = __ldg(&csc_pairs[ch.begin]).x;
    auto load_pair = [&](int k) -> int2 {
        return (k + lane < ch.end) ? __ldg(&csc_pairs[k + lane]) : make_int2(first_cell, 0);   // tail: adds +0.0f
    };
    LW lw_cur[B], lw_nxt[B];
    int2 pr_cur = load_pair(ch.begin);
    int2 pr_nxt = load_pair(ch.begin + B);          // entry pairs run two blocks ahead of the updates
#pragma unroll
    for (int u = 0; u < B; ++u) lw_cur[u] = __ldg(slab_w + (int64_t)__shfl_sync(0xffffffffu, pr_cur.x, u) * row_words);
    for (int k = ch.begin; k < ch.end; k += B) {
        const bool more = k + B < ch.end;
        const int2 pr_nxt2 = load_pair(k + 2 * B);
        if (more) {
#pragma unroll
            for (int u = 0; u < B; ++u) lw_nxt[u] = __ldg(slab_w + (int64_t)__shfl_sync(0xffffffffu, pr_nxt.x, u) * row_words);
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const float v = __int_as_float(__shfl_sync(0xffffffffu, pr_cur.y, u));
            fast_update<Q>(mybins, (uint32_t)lw_cur[u], v);
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < B; ++u) lw_cur[u] = lw_nxt[u];
        }
        pr_cur = pr_nxt;
        pr_nxt = pr_nxt2;
    }
    __syncwarp();
    // partial sums out: partials[(slot*L + c)*PB + group*32*Q + lane*Q + q]
    float* out = partials + (int64_t)ch.slot * L * PB + group * 32 * Q + lane * Q;
    for (int c = 0; c < L; ++c) {
        float vals[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) vals[q] = mybins[(c * Q + q) * 32];
        float* dst = out + (int64_t)c * PB;
        if (Q == 4) {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(vals[0], vals[Q > 1 ? 1 : 0], vals[Q > 2 ? 2 : 0], vals[Q > 3 ? 3 : 0]));
        } else if (Q == 2) {
            __stcs(reinterpret_cast<float2*>(dst), make_float2(vals[0], vals[Q > 1 ? 1 : 0]));
        } else {
            __stcs(dst, vals[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// FAST mode, step 2 (v2): same decomposition as fast_means_kernel (one warp = one chunk of one gene's
// stored entries x 32*Q permutations, private per-lane bins, deterministic ascending-cell order),
// restructured around the shared-memory pipe, which is what bounds it:
//   * one warp per CTA, bins at the start of dynamic shared memory, so a bin address is ONE PRMT:
//     byte 0 = lane*4, byte 1 = the label byte (the slab stores label * Q/2, i.e. label*Q*128 >> 8),
//     and the q*128 plane offset is an immediate of the LDS/STS -> 4 instructions per update
//     (PRMT, LDS, FADD, STS) instead of ~9;
//   * the {cell, value} pairs of a block of B entries are staged in shared memory once and read back
//     as 128-bit broadcasts (0.5 wavefronts per entry) instead of two shuffles per entry (2 wavefronts);
//   * label words of block b+1 are in flight while block b updates (registers ping-pong, no moves).
// ------------------------------------------------------------------------------------------
template <int Q> struct LabelVec;
template <> struct LabelVec<2> { using type = uint16_t; };
template <> struct LabelVec<4> { using type = uint32_t; };
template <> struct LabelVec<8> { using type = uint2; };

template <int Q> __device__ __forceinline__ uint32_t label_word(const typename LabelVec<Q>::type& w, int q);
template <> __device__ __forceinline__ uint32_t label_word<2>(const uint16_t& w, int) { return (uint32_t)w; }
template <> __device__ __forceinline__ uint32_t label_word<4>(const uint32_t& w, int) { return w; }
template <> __device__ __forceinline__ uint32_t label_word<8>(const uint2& w, int q) { return q < 4 ? w.x : w.y; }

template <int Q>
__device__ __forceinline__ void fast2_update(unsigned char* bins, uint32_t lane4, const typename LabelVec<Q>::type& lw, float v) {
    float* b[Q];
    float x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        // {byte1 = label byte q, byte0 = lane*4}: offset of bin [label][q=0][lane]
        const uint32_t off = __byte_perm(label_word<Q>(lw, q), lane4, 0x5504u | ((uint32_t)(q & 3) << 4));
        LRPERM_CHECK((long long)off + q * 128 + 4 <= g_dbg_dims.L * Q * 128, kDbgBin, off);
        b[q] = reinterpret_cast<float*>(bins + off + q * 128);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) x[q] = *b[q];
#pragma unroll
    for (int q = 0; q < Q; ++q) *b[q] = x[q] + v;
}

// COUNT = true: every stored entry contributes 1.0f instead of its value (per-(chunk, cluster, permutation)
// entry counts of the trimean path, lrperm_trimean.cuh; exact in float32: a chunk holds < 2^24 entries).
template <int Q, int B, bool COUNT = false>
__global__ void __launch_bounds__(32)
fast2_means_kernel(const int2* __restrict__ csc_pairs, const FastChunk* __restrict__ chunks, int32_t n_chunks,
                   const uint8_t* __restrict__ slab, int32_t PB, int32_t L, int32_t groups_per_batch,
                   float* __restrict__ partials /*[slot][L][PB]*/) {
    extern __shared__ __align__(16) unsigned char smem2[];
    using LV = typename LabelVec<Q>::type;
    static_assert(B % 4 == 0 && B <= 32, "block of staged entries");
    const int lane = threadIdx.x;
    const int chunk_id = (int)(blockIdx.x / (unsigned)groups_per_batch);
    const int group = (int)(blockIdx.x - (unsigned)chunk_id * (unsigned)groups_per_batch);
    const FastChunk ch = chunks[chunk_id];
    const int bins_bytes = L * Q * 128;
    unsigned char* bins = smem2;
    int* st_cell = reinterpret_cast<int*>(smem2 + bins_bytes);                 // [2][B]
    float* st_val = reinterpret_cast<float*>(smem2 + bins_bytes + 2 * B * 4);  // [2][B]
    {
        float4* z = reinterpret_cast<float4*>(bins);
        for (int i = lane; i < L * Q * 8; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const LV* __restrict__ slab_w = reinterpret_cast<const LV*>(slab) + (group * 32 + lane);
    const int64_t row_words = PB / Q;              // label words per slab row
    const uint32_t lane4 = (uint32_t)lane * 4u;
    LRPERM_CHECK(ch.begin >= 0 && ch.begin < ch.end && ch.end <= g_dbg_dims.nnz, kDbgChunk, ch.begin);
    LRPERM_CHECK(ch.slot >= 0 && ch.slot < g_dbg_dims.n_slots && ch.gene >= 0, kDbgSlot, ch.slot);
    const int first_cell = __ldg(&csc_pairs[ch.begin]).x;

    auto load_pair = [&](int k) -> int2 {          // tail entries add +0.0f to a valid bin
        return (lane < B && k + lane < ch.end) ? __ldg(&csc_pairs[k + lane]) : make_int2(first_cell, 0);
    };
    auto load_pair_count = [&](int k) -> int2 {    // COUNT: 1.0f per real entry, +0.0f for the tail
        const bool in = lane < B && k + lane < ch.end;
        return make_int2(in ? __ldg(&csc_pairs[k + lane]).x : first_cell, in ? 0x3f800000 : 0);
    };
    auto stage = [&](int buf, int2 pr) {
        __syncwarp();                              // earlier broadcasts of this buffer are done
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
    auto update = [&](int buf, const LV (&lw)[B]) {
#pragma unroll
        for (int u = 0; u < B; u += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(st_val + buf * B + u);
            fast2_update<Q>(bins, lane4, lw[u + 0], v4.x);
            fast2_update<Q>(bins, lane4, lw[u + 1], v4.y);
            fast2_update<Q>(bins, lane4, lw[u + 2], v4.z);
            fast2_update<Q>(bins, lane4, lw[u + 3], v4.w);
        }
    };

    auto next_pair = [&](int k) -> int2 { return COUNT ? load_pair_count(k) : load_pair(k); };
    LV lwA[B], lwB[B];
    int2 pr = next_pair(ch.begin);
    stage(0, pr);
    fetch_labels(0, lwA);
    pr = next_pair(ch.begin + B);
    for (int k = ch.begin; k < ch.end; k += 2 * B) {
        // block k: buffer 0 / lwA ; block k+B: buffer 1 / lwB
        stage(1, pr);
        pr = next_pair(k + 2 * B);
        const bool more1 = k + B < ch.end;
        if (more1) fetch_labels(1, lwB);
        update(0, lwA);
        if (!more1) break;
        stage(0, pr);
        pr = next_pair(k + 3 * B);
        if (k + 2 * B < ch.end) fetch_labels(0, lwA);
        update(1, lwB);
    }
    __syncwarp();
    // partial sums out: partials[(slot*L + c)*PB + group*32*Q + lane*Q + q]
    const float* mybins = reinterpret_cast<const float*>(bins) + lane;
    float* out = partials + (int64_t)ch.slot * L * PB + group * 32 * Q + lane * Q;
    for (int c = 0; c < L; ++c) {
        float vals[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) vals[q] = mybins[(c * Q + q) * 32];
        float* dst = out + (int64_t)c * PB;
        if (Q == 2) {
            __stcs(reinterpret_cast<float2*>(dst), make_float2(vals[0], vals[1]));
        } else {
#pragma unroll
            for (int q = 0; q < Q; q += 4)
                __stcs(reinterpret_cast<float4*>(dst + q), make_float4(vals[q], vals[(q + 1) % Q], vals[(q + 2) % Q], vals[(q + 3) % Q]));
        }
    }
}

// ------------------------------------------------------------------------------------------
// FAST mode, step 2 (v3, EXPERIMENT - option "fast_variant" = 3): fast2_means_kernel with the {cell, value} entries
// brought into shared memory by the bulk-copy engine instead of the LSU.  One elected lane issues
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (1-D TMA, SASS: UBLKCP) for a ring of R blocks of
// B = 32 entries (256 B each); an mbarrier per ring slot counts the arriving bytes, the warp waits on its phase bit,
// reads the block back as 128-bit broadcasts ({cell, value} x 2 per load) and re-arms the slot for block j + R.  Against
// v2 this takes the entry LDG (2 wavefronts per block) and the STS staging (2 wavefronts + 2 __syncwarp) off the
// LSU / shared-memory path; the bin read-modify-write (256 wavefronts per block) and the label words (32) stay.
// Bulk copies need 16-byte aligned addresses: the chunk is walked from its even-aligned start, and the (at most 1 + 31)
// entries outside [begin, end) that the first / last block carries get the value +0.0f (their cells are valid cells of
// the neighbouring chunks, or the zero padding behind the last entry, so the label loads stay in bounds).
// Same summation order as v2 -> bit-identical results (scripts/fast_ab.py, tests).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr_u32(dst)), "l"(src), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}

template <int Q, int R>
__global__ void __launch_bounds__(32)
fast3_means_kernel(const int2* __restrict__ csc_pairs, const FastChunk* __restrict__ chunks, int32_t n_chunks,
                   const uint8_t* __restrict__ slab, int32_t PB, int32_t L, int32_t groups_per_batch,
                   float* __restrict__ partials /*[slot][L][PB]*/) {
    extern __shared__ __align__(16) unsigned char smem2[];
    using LV = typename LabelVec<Q>::type;
    constexpr int B = 32;
    const int lane = threadIdx.x;
    const int chunk_id = (int)(blockIdx.x / (unsigned)groups_per_batch);
    const int group = (int)(blockIdx.x - (unsigned)chunk_id * (unsigned)groups_per_batch);
    const FastChunk ch = chunks[chunk_id];
    const int bins_bytes = L * Q * 128;
    unsigned char* bins = smem2;
    int2* ring = reinterpret_cast<int2*>(smem2 + bins_bytes);                               // [R][B]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem2 + bins_bytes + R * B * (int)sizeof(int2));   // [R]
    {
        float4* z = reinterpret_cast<float4*>(bins);
        for (int i = lane; i < L * Q * 8; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int ka = ch.begin & ~1;                                   // 16-byte aligned start of the walk
    const int nblk = (ch.end - ka + B - 1) / B;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) mbar_init(&mbar[r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r < nblk) {
                mbar_expect_tx(&mbar[r], B * (uint32_t)sizeof(int2));
                bulk_copy_g2s(ring + r * B, csc_pairs + ka + r * B, B * (uint32_t)sizeof(int2), &mbar[r]);
            }
    }
    __syncwarp();
    const LV* __restrict__ slab_w = reinterpret_cast<const LV*>(slab) + (group * 32 + lane);
    const int64_t row_words = PB / Q;
    const uint32_t lane4 = (uint32_t)lane * 4u;

    auto fetch = [&](int j, LV (&lw)[B], float (&vv)[B]) {
        const int slot = j % R;
        mbar_wait(&mbar[slot], (uint32_t)((j / R) & 1));
        if (j == 0 || j == nblk - 1) {                              // entries outside [begin, end): value +0.0f
            const int idx = ka + j * B + lane;
            if (idx < ch.begin || idx >= ch.end) ring[slot * B + lane].y = 0;
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < B; u += 2) {
            const int4 c4 = *reinterpret_cast<const int4*>(ring + slot * B + u);
            LRPERM_CHECK((unsigned)c4.x < (unsigned long long)g_dbg_dims.N && (unsigned)c4.z < (unsigned long long)g_dbg_dims.N, kDbgCell, c4.x);
            lw[u + 0] = __ldg(slab_w + (int64_t)c4.x * row_words);
            lw[u + 1] = __ldg(slab_w + (int64_t)c4.z * row_words);
            vv[u + 0] = __int_as_float(c4.y);
            vv[u + 1] = __int_as_float(c4.w);
        }
        __syncwarp();                                               // every lane has read the slot: re-arm it
        if (lane == 0 && j + R < nblk) {
            mbar_expect_tx(&mbar[slot], B * (uint32_t)sizeof(int2));
            bulk_copy_g2s(ring + slot * B, csc_pairs + ka + (j + R) * B, B * (uint32_t)sizeof(int2), &mbar[slot]);
        }
    };
    auto update = [&](const LV (&lw)[B], const float (&vv)[B]) {
#pragma unroll
        for (int u = 0; u < B; ++u) fast2_update<Q>(bins, lane4, lw[u], vv[u]);
    };

    LV lwA[B], lwB[B];
    float vA[B], vB[B];
    fetch(0, lwA, vA);
    for (int j = 0; j < nblk; j += 2) {
        const bool more1 = j + 1 < nblk;
        if (more1) fetch(j + 1, lwB, vB);
        update(lwA, vA);
        if (!more1) break;
        if (j + 2 < nblk) fetch(j + 2, lwA, vA);
        update(lwB, vB);
    }
    __syncwarp();
    const float* mybins = reinterpret_cast<const float*>(bins) + lane;
    float* out = partials + (int64_t)ch.slot * L * PB + group * 32 * Q + lane * Q;
    for (int c = 0; c < L; ++c) {
        float vals[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) vals[q] = mybins[(c * Q + q) * 32];
        float* dst = out + (int64_t)c * PB;
#pragma unroll
        for (int q = 0; q < Q; q += 4)
            __stcs(reinterpret_cast<float4*>(dst + q), make_float4(vals[q], vals[(q + 1) % Q], vals[(q + 2) % Q], vals[(q + 3) % Q]));
    }
}

// FAST mode, step 3: combine a gene's chunk partials in fixed order (cell tiles ascending, chunks ascending inside a
// tile: slots are numbered tile by tile, tile_slot_ptr[t][g] .. [t][g+1] are gene g's slots in tile t), scale by
// 1/n_c, write cube.
// `needed` (or nullptr): genes some kept triple reads; the others were not accumulated (their chunk warps never ran) and
// their cube rows are left as they are - nothing reads them.
__global__ void fast_finalize_kernel(const float* __restrict__ partials, const int32_t* __restrict__ tile_slot_ptr /*[n_tiles][G+1]*/,
                                     int32_t n_tiles, const float* __restrict__ inv_n, int32_t G, int32_t L, int32_t PB,
                                     float* __restrict__ cube, int32_t PP, int32_t p_off, const uint8_t* __restrict__ needed = nullptr) {
    // thread per (g, c, p) with p fastest
    const int64_t total = (int64_t)G * L * PB;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % PB);
        const int64_t gc = i / PB;
        const int c = (int)(gc % L);
        const int g = (int)(gc / L);
        if (needed && !needed[g]) continue;
        float s = 0.0f;
        for (int t = 0; t < n_tiles; ++t) {
            const int s0 = tile_slot_ptr[(int64_t)t * (G + 1) + g], s1 = tile_slot_ptr[(int64_t)t * (G + 1) + g + 1];
            LRPERM_CHECK(s0 >= 0 && s0 <= s1 && s1 <= g_dbg_dims.n_slots, kDbgSlot, s1);
            for (int sl = s0; sl < s1; ++sl) s = __fadd_rn(s, __ldcs(&partials[((int64_t)sl * L + c) * PB + p]));
        }
        cube[((int64_t)g * L + c) * PP + p_off + p] = __fmul_rn(s, inv_n[c]);
    }
}

// ------------------------------------------------------------------------------------------
// Score + compare + count (never materialises the reference's perm_stats[2,P,T]).
// Lanes run over 32 consecutive permutations (coalesced 128 B reads of the permutation-minor
// cube), each warp sweeps the CTA's tile of triples; ballots turn 32 comparisons into one popc.
//   CellPhoneDB : (a + b)/2 >= m      in float64 (exact for float32 a, b)   _cellphonedb.py:28
//   gmean       : a*b >= g*g          in float64 (exact products)           _geometric_mean.py:23
//                 or exp((log a + log b)/2) >= g (LITERAL)
// ------------------------------------------------------------------------------------------
// blockIdx.y = block of `groups_per_block` 32-permutation groups: the CTAs of one permutation block are scheduled
// together and touch only that slice of every cube row, so the gathered slice (G x L x groups_per_block x 128 B) stays
// L2 resident instead of streaming the whole cube batch from DRAM once per 256-triple tile.
template <bool CPDB, bool GMEAN, bool LITERAL>
__global__ void __launch_bounds__(kScoreThreads)
score_count_kernel(const float* __restrict__ cube, int32_t PP, int32_t n_perms_batch, int32_t groups_per_block,
                   const int32_t* __restrict__ rowA, const int32_t* __restrict__ rowB,
                   const float* __restrict__ thr_c, const float* __restrict__ thr_g, int64_t T,
                   uint32_t* __restrict__ counts_c, uint32_t* __restrict__ counts_g) {
    __shared__ int32_t s_ra[kScoreTile], s_rb[kScoreTile];
    __shared__ float s_tc[kScoreTile], s_tg[kScoreTile];
    __shared__ uint32_t s_cc[kScoreTile], s_cg[kScoreTile];
    const int64_t t0 = (int64_t)blockIdx.x * kScoreTile;
    const int nt = (int)min((int64_t)kScoreTile, T - t0);
    for (int i = threadIdx.x; i < kScoreTile; i += blockDim.x) {
        const bool ok = i < nt;
        s_ra[i] = ok ? rowA[t0 + i] : 0;
        s_rb[i] = ok ? rowB[t0 + i] : 0;
        s_tc[i] = (CPDB && ok) ? thr_c[t0 + i] : 0.0f;
        s_tg[i] = (GMEAN && ok) ? thr_g[t0 + i] : 0.0f;
        s_cc[i] = 0; s_cg[i] = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = kScoreThreads / 32;
    const int n_groups = (n_perms_batch + 31) >> 5;
    const int g_first = blockIdx.y * groups_per_block;
    const int n_g = min(groups_per_block, n_groups - g_first);
    const int n_tb = (nt + 31) >> 5;
    // work items: (permutation group of this block, 32-triple sub-tile)
    for (int item = warp; item < n_g * n_tb; item += NW) {
        const int pg = g_first + item % n_g;
        const int tb = (item / n_g) * 32;
        const int p = pg * 32 + lane;
        const bool valid = p < n_perms_batch;
        {
            uint32_t my_c = 0, my_g = 0;
            const int tn = min(32, nt - tb);
            for (int u = 0; u < tn; ++u) {
                const int t = tb + u;
                LRPERM_CHECK((unsigned)s_ra[t] < (unsigned long long)(g_dbg_dims.G * g_dbg_dims.L) &&
                             (unsigned)s_rb[t] < (unsigned long long)(g_dbg_dims.G * g_dbg_dims.L), kDbgCubeRow, s_ra[t]);
                const float a = __ldg(&cube[(int64_t)s_ra[t] * PP + p]);
                const float b = __ldg(&cube[(int64_t)s_rb[t] * PP + p]);
                if (CPDB) {
                    const bool ge = valid && (((double)a + (double)b) * 0.5 >= (double)s_tc[t]);
                    const uint32_t m = __ballot_sync(0xffffffffu, ge);
                    if (lane == u) my_c = __popc(m);
                }
                if (GMEAN) {
                    bool ge;
                    const double g = (double)s_tg[t];
                    if (LITERAL) {
                        ge = exp((log((double)a) + log((double)b)) * 0.5) >= g;
                    } else {
                        // sqrt(ab) >= g  <=>  ab >= g^2 for a, b, g >= 0 ; negatives -> NaN in the reference -> false
                        ge = (a >= 0.0f) && (b >= 0.0f) && ((g <= 0.0) || ((double)a * (double)b >= g * g));
                    }
                    ge = ge && valid;
                    const uint32_t m = __ballot_sync(0xffffffffu, ge);
                    if (lane == u) my_g = __popc(m);
                }
            }
            if (lane < tn) {
                if (CPDB) atomicAdd(&s_cc[tb + lane], my_c);
                if (GMEAN) atomicAdd(&s_cg[tb + lane], my_g);
            }
        }
    }
    __syncthreads();
    // integer adds: several permutation blocks (and batches) accumulate into the same counters, in any order
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        if (CPDB && s_cc[i]) atomicAdd(&counts_c[t0 + i], s_cc[i]);
        if (GMEAN && s_cg[i]) atomicAdd(&counts_g[t0 + i], s_cg[i]);
    }
}

// cube[g][c][PP] (perm-minor) -> out[p][c][g] (the reference's layout)
__global__ void cube_export_kernel(const float* __restrict__ cube, int32_t G, int32_t L, int32_t PP,
                                   int32_t n_perms_batch, float* __restrict__ out) {
    __shared__ float tile[32][33];
    // grid: x over gene tiles, y over perm tiles, z over clusters
    const int g0 = blockIdx.x * 32, p0 = blockIdx.y * 32, c = blockIdx.z;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int g = g0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (g < G && p < n_perms_batch) ? cube[((int64_t)g * L + c) * PP + p] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, g = g0 + threadIdx.x;
        if (g < G && p < n_perms_batch) out[((int64_t)p * L + c) * G + g] = tile[threadIdx.x][r];
    }
}

// needed[g] = 1 for every gene column some triple reads (ligand or receptor side): FAST mode accumulates only those
// when no cube is exported
__global__ void mark_genes_kernel(const int32_t* __restrict__ lig, const int32_t* __restrict__ rec, int64_t T, int32_t G,
                                  uint8_t* __restrict__ needed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = lig[i], r = rec[i];
        if (l >= 0 && l < G) needed[l] = 1;
        if (r >= 0 && r < G) needed[r] = 1;
    }
}

__global__ void build_rows_kernel(const int32_t* __restrict__ lig, const int32_t* __restrict__ rec,
                                  const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                  int64_t T, int32_t G, int32_t L, int32_t* __restrict__ rowA,
                                  int32_t* __restrict__ rowB, int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < T) {
        const int l = lig[i], r = rec[i], s = src[i], t = tgt[i];
        if (l < 0 || l >= G || r < 0 || r >= G || s < 0 || s >= L || t < 0 || t >= L) { atomicOr(err, 8); rowA[i] = 0; rowB[i] = 0; return; }
        rowA[i] = l * L + s;
        rowB[i] = r * L + t;
    }
}

// ------------------------------------------------------------------------------------------
// expr_prop filter + min-subunit resolution of the complexes + name -> index maps
// (liana/resource/_reassemble_complexes.py:11-101, liana/method/_pipe_utils/_get_mean_perms.py:90-113): the kept
// (source, target, interaction) rows with their resolved subunit genes and statistics, built from two
// [n_inter][L] side tables instead of pandas merges over the L^2 x 5,780 exploded rows.
// ------------------------------------------------------------------------------------------
// side tables: for (interaction i, cluster c) the subunit with the FIRST minimal statistic in the complex's '_' order
// (or the first subunit when `first_subunit`: return_all_lrs de-duplicates before the min-subunit step), its
// statistic / prop, and the minimal prop over all subunits
template <typename ST>
__global__ void resolve_side_kernel(const int32_t* __restrict__ sub /*[n][m], -1 padded*/, int32_t n, int32_t m, int32_t L, int32_t G,
                                    const ST* __restrict__ stat /*[L][G]*/, const double* __restrict__ props /*[L][G]*/, int first_subunit,
                                    int32_t* __restrict__ gene /*[n][L]*/, ST* __restrict__ statv, double* __restrict__ propv,
                                    double* __restrict__ pmin) {
    const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (id >= (int64_t)n * L) return;
    const int i = (int)(id / L), c = (int)(id % L);
    int best = 0;
    ST best_v = (ST)0;
    double pm = 0.0;
    bool any = false;
    for (int k = 0; k < m; ++k) {
        const int g = sub[(int64_t)i * m + k];
        if (g < 0) continue;
        const ST v = stat[(int64_t)c * G + g];
        const double p = props[(int64_t)c * G + g];
        if (!any) { best = g; best_v = v; pm = p; any = true; }
        else {
            if (!first_subunit && v < best_v) { best = g; best_v = v; }
            pm = fmin(pm, p);
        }
    }
    gene[id] = best;
    statv[id] = any ? stat[(int64_t)c * G + best] : (ST)0;
    propv[id] = any ? props[(int64_t)c * G + best] : 0.0;
    pmin[id] = any ? pm : 0.0;
}

// flat row f = (t * L + s) * n + i: the reference concatenates per (source, target) pair with the SOURCE varying
// fastest (`np.meshgrid(labels, labels).reshape(2, L*L).T`, liana/method/sc/_liana_pipe.py:310-312), interactions inside
__global__ void resolve_flags_kernel(const double* __restrict__ pmin_l, const double* __restrict__ pmin_r, const uint8_t* __restrict__ pair_mask,
                                     int32_t n, int32_t L, double expr_prop, int return_all, int32_t* __restrict__ flag) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= (int64_t)L * L * n) return;
    const int i = (int)(f % n);
    const int64_t st = f / n;
    const int t = (int)(st / L), s = (int)(st % L);
    const bool expressed = pmin_l[(int64_t)i * L + s] >= expr_prop && pmin_r[(int64_t)i * L + t] >= expr_prop;
    const bool valid = pair_mask ? pair_mask[s * L + t] != 0 : true;
    flag[f] = (return_all ? valid : (valid && expressed)) ? 1 : 0;
}

template <typename ST>
__global__ void resolve_scatter_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int32_t n, int32_t L, double expr_prop,
                                       const int32_t* __restrict__ gene_l, const ST* __restrict__ stat_l, const double* __restrict__ prop_l,
                                       const double* __restrict__ pmin_l, const int32_t* __restrict__ gene_r, const ST* __restrict__ stat_r,
                                       const double* __restrict__ prop_r, const double* __restrict__ pmin_r,
                                       int32_t* __restrict__ o_inter, int32_t* __restrict__ o_src, int32_t* __restrict__ o_tgt,
                                       int32_t* __restrict__ o_lg, int32_t* __restrict__ o_rg, ST* __restrict__ o_ls, ST* __restrict__ o_rs,
                                       double* __restrict__ o_lp, double* __restrict__ o_rp, uint8_t* __restrict__ o_keep) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= (int64_t)L * L * n || !flag[f]) return;
    const int i = (int)(f % n);
    const int64_t st = f / n;
    const int t = (int)(st / L), s = (int)(st % L);
    const int64_t a = (int64_t)i * L + s, b = (int64_t)i * L + t;
    const int32_t r = pos[f];
    o_inter[r] = i; o_src[r] = s; o_tgt[r] = t;
    o_lg[r] = gene_l[a]; o_rg[r] = gene_r[b];
    o_ls[r] = stat_l[a]; o_rs[r] = stat_r[b];
    o_lp[r] = prop_l[a]; o_rp[r] = prop_r[b];
    o_keep[r] = (pmin_l[a] >= expr_prop && pmin_r[b] >= expr_prop) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// Consensus of method scores (liana/method/_pipe_utils/_aggregate.py:70-181): average-method ranks of a score
// column (scipy.stats.rankdata(method='average'), :91-98) and the RobustRankAggregate rho of every row (:110-181).
// ------------------------------------------------------------------------------------------
// order-preserving 64-bit key of (sign * v); -0.0 and +0.0 share a key (they tie in rankdata)
__global__ void rank_keys_kernel(const double* __restrict__ v, int64_t n, int negate, unsigned long long* __restrict__ keys,
                                 uint32_t* __restrict__ idx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = v[i];
    if (negate) x = -x;
    unsigned long long b = (x == 0.0) ? 0ull : (unsigned long long)__double_as_longlong(x);
    keys[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    idx[i] = (uint32_t)i;
}

// flag[i] = 1 where a new tie group starts in the sorted key array
__global__ void rank_flags_kernel(const unsigned long long* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// group_start[g] = sorted position of the first member of tie group g (gid = inclusive sum of the flags, 1-based)
__global__ void rank_starts_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ gid, int64_t n,
                                   int32_t* __restrict__ group_start) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && flag[i]) group_start[gid[i] - 1] = (int32_t)i;
}

// average rank of the tie group [s, e] (0-based sorted positions) = (s + e) / 2 + 1, written at the original index
__global__ void rank_scatter_kernel(const int32_t* __restrict__ gid, const int32_t* __restrict__ group_start, const uint32_t* __restrict__ idx,
                                    int64_t n, int32_t n_groups, double* __restrict__ ranks) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = gid[i] - 1;
    const int64_t s = group_start[g];
    const int64_t e = (g + 1 < n_groups ? (int64_t)group_start[g + 1] : n) - 1;
    ranks[idx[i]] = (double)(s + e + 2) * 0.5;
}

// rho of one row: ranks / column maximum, sorted ascending, p_k = I_{r_(k)}(k, n - k + 1) (the regularized incomplete
// beta function with integer shape parameters = a binomial tail: sum_{j >= k} C(n, j) x^j (1 - x)^(n - j)), row
// minimum, Bonferroni over the n methods, clipped to [0, 1]
constexpr int kRraMaxCols = 16;
__global__ void rra_kernel(const double* __restrict__ rmat /*[n_rows][n_cols]*/, const double* __restrict__ col_max, int64_t n_rows,
                           int32_t n, double* __restrict__ rho) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double r[kRraMaxCols];
    bool has_nan = false;
    for (int c = 0; c < n; ++c) { r[c] = rmat[row * n + c] / col_max[c]; has_nan |= isnan(r[c]); }
    if (has_nan) { rho[row] = r[0] + r[n - 1] + nan(""); return; }      // beta.cdf(nan) -> np.min -> nan
    for (int a = 1; a < n; ++a) {                       // insertion sort, n <= 16
        const double x = r[a];
        int b = a - 1;
        while (b >= 0 && r[b] > x) { r[b + 1] = r[b]; --b; }
        r[b + 1] = x;
    }
    double best = 1.0;
    for (int k = 1; k <= n; ++k) {
        const double x = r[k - 1], y = 1.0 - x;
        // binomial tail, terms built incrementally from j = k: C(n,k) x^k y^(n-k), then * (n-j)/(j+1) * x/y
        double p;
        if (x <= 0.0) p = 0.0;
        else if (x >= 1.0) p = 1.0;
        else {
            double coef = 1.0;
            for (int j = 1; j <= k; ++j) coef = coef * (double)(n - k + j) / (double)j;       // C(n, k)
            double term = coef * pow(x, (double)k) * pow(y, (double)(n - k));
            p = term;
            for (int j = k; j < n; ++j) { term = term * (double)(n - j) / (double)(j + 1) * (x / y); p += term; }
        }
        best = fmin(best, p);
    }
    rho[row] = fmin(fmax(best * (double)n, 0.0), 1.0);
}

}  // namespace lrperm
