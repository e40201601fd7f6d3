// lrperm_api.cu - host side of the C ABI declared in include/lrperm.h.
//
// One engine = one CUDA device + one stream.  The engine owns device copies of the matrix
// (CSR pairs; a CSC copy is built lazily for FAST mode), the label tables, the triple tables
// and the work buffers (label slab, partial sums, cube batch, counters).  `lrperm_run` loops
// over permutation batches:   labels/permutation setup -> per-cluster means -> score+count.
//
// Replaces liana/method/sc/_liana_pipe.py:402-425 + liana/method/_pipe_utils/_get_mean_perms.py
// as one fused call that never materialises `perms` (P x L x G float64) or `perm_stats`
// (2 x P x T float64).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "lrperm.h"
#include "lrperm_kernels.cuh"
#include "lrperm_trimean.cuh"

using namespace lrperm;

namespace {

std::string g_create_error;

// LRPERM_DEBUG_TIMING=1: wall-clock of the host-side stages on stderr (development aid)
struct StageTimer {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit StageTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {
        static const bool enabled = std::getenv("LRPERM_DEBUG_TIMING") != nullptr;
        on = enabled;
    }
    ~StageTimer() {
        if (on) fprintf(stderr, "[lrperm] %-18s %8.3f ms\n", name,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Pinned host buffer that only grows.  Outgrown blocks are kept until the engine is destroyed: a copy kernel or an
// asynchronous copy queued earlier may still be reading the old block.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    std::vector<void*> retired;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) retired.push_back(p);
        p = nullptr; cap = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 2, 1u << 16);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        for (void* q : retired) cudaFreeHost(q);
        p = nullptr; cap = 0; retired.clear();
        cudaGetLastError();
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Host-side helper of pageable uploads: a few persistent threads that memcpy slices of one piece into a pinned
// staging buffer in parallel (one thread moves ~10 GB/s, PCIe takes ~50 GB/s).
class ParallelCopier {
public:
    explicit ParallelCopier(int n) : n_(n) {
        for (int i = 0; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~ParallelCopier() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // narrow = false: memcpy of `bytes`.  narrow = true: `bytes` OUTPUT bytes of uint16 written from bytes / 2 int32 inputs
    // (column indices < 65536 travel over PCIe at half their size); returns the OR of the inputs' high halves, i.e.
    // non-zero when an index was negative or >= 65536 (it would not survive the narrowing).
    uint32_t copy(void* dst, const void* src, size_t bytes, bool narrow = false) {
        if (bytes < (1u << 20)) {
            if (!narrow) { memcpy(dst, src, bytes); return 0; }
            return narrow_slice(static_cast<uint16_t*>(dst), static_cast<const int32_t*>(src), bytes / 2);
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = static_cast<unsigned char*>(dst); src_ = static_cast<const unsigned char*>(src); bytes_ = bytes;
            narrow_ = narrow; bad_ = 0;
            pending_ = n_; ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        return bad_;
    }
    static uint32_t narrow_slice(uint16_t* d, const int32_t* s, size_t n) {
        uint32_t bad = 0;
        for (size_t i = 0; i < n; ++i) { const uint32_t v = (uint32_t)s[i]; bad |= v >> 16; d[i] = (uint16_t)v; }
        return bad;
    }
private:
    void loop(int i) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            const size_t slice = ((bytes_ + n_ - 1) / n_ + 63) & ~(size_t)63;
            const size_t b = std::min(bytes_, slice * i), e = std::min(bytes_, slice * (i + 1));
            unsigned char* d = dst_; const unsigned char* s = src_;
            const bool narrow = narrow_;
            lk.unlock();
            uint32_t bad = 0;
            if (e > b) {
                if (!narrow) memcpy(d + b, s + b, e - b);
                else bad = narrow_slice(reinterpret_cast<uint16_t*>(d + b), reinterpret_cast<const int32_t*>(s + 2 * b), (e - b) / 2);
            }
            lk.lock();
            bad_ |= bad;
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    unsigned char* dst_ = nullptr; const unsigned char* src_ = nullptr; size_t bytes_ = 0;
    bool narrow_ = false;
    uint32_t bad_ = 0;
    int pending_ = 0;
};

}  // namespace

struct lrperm_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t aux_stream = nullptr;    // FAST mode: label slabs / finalize / score run here, overlapped with the main kernel
    cudaStream_t copy_stream = nullptr;   // matrix uploads from pinned host memory (lrperm_set_matrix_csr) run here, unwaited
    cudaEvent_t ev_matrix = nullptr;      // recorded on copy_stream after the upload + packing
    bool matrix_pending = false;          // the engine stream has not waited for ev_matrix yet (validation deferred too)
    // tile-wise upload: lrperm_set_matrix_csr submits every cell tile to the copy stream, each followed there by its
    // packing and CSC sort.  The copy engine serves H2D copies of all streams in submission order, so while the
    // upload is pending the small tables of set_labels / set_triples / the chunk tables go through copy KERNELS
    const void* up_indices = nullptr;
    const void* up_data = nullptr;
    std::vector<int64_t> up_e;            // entry offset at every tile boundary [n_tiles + 1]
    int32_t up_next_tile = 0;
    bool check_matrix_flag = false;
    int32_t eager_csc = 1;                // option: queue the CSC build of FAST mode as soon as the matrix is set
    int32_t score_gpb = 0;                // option: 32-permutation groups per block of the score kernels (0 = auto: L2 fit)
    int32_t slab0_by_tile = 1;            // option: first batch's label slab generated / consumed cell tile by cell tile
    std::vector<cudaEvent_t> ev_slab_tile;
    std::vector<cudaEvent_t> ev_main_tile;   // [2 * tile]: start / stop of the first batch's chunk warps of one tile (timing only)
    int32_t main_tiles_timed = 0;            // tiles the first batch of the last run was launched by (0 = one launch)       // the gated first batch consumed a pending upload: read err_flag_m after the run
    DevBuf err_flag_m;                    // validation flags of the pending matrix
    // pageable uploads: two pinned staging pieces filled by a few host threads while the other one is in flight
    ParallelCopier* copier = nullptr;
    static constexpr int kMaxPin = 4;
    void* pin[kMaxPin] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pin_ev[kMaxPin] = {nullptr, nullptr, nullptr, nullptr};
    int n_pin = 2;                        // staging pieces in rotation  (LRPERM_PIN_PIECES, 2..4)
    size_t pin_piece = 16u << 20;         // bytes per piece             (LRPERM_PIN_PIECE_MB)
    int32_t overlap = 1;
    std::string err;
    int sm_count = 148;
    size_t smem_optin = 0;

    // matrix
    int64_t N = 0, G = 0, nnz = 0;
    DevBuf indptr, csr_pairs, csc_pairs, err_flag;
    bool have_matrix = false, csc_ready = false;
    bool csr_packed = true;               // false after a pinned upload: csr_pairs is built from the raw arrays on first use
    std::vector<int32_t> h_csc_ptr;   // [G+1] gene pointers of the value-sorted CSC (trimean path)
    // FAST mode: the CSC copy is built cell tile by cell tile (the entries of a tile's rows are contiguous in CSR
    // order and the tile's gene-major copy occupies the same index range), so a tile can be sorted - and its chunks
    // run - while later tiles are still being uploaded
    int32_t csc_tile_cells = 0, n_tiles = 0, tiles_enqueued = 0, tiles_tabled = 0;
    PinnedBuf up_pin;                     // pinned mirror of small pageable uploads made while a matrix upload is pending
    size_t up_pin_off = 0;                // bump offset; reset by set_labels / set_triples (which end with a stream sync)
    PinnedBuf hist_pin;                   // pinned host [n_tiles][G] column counts per tile
    PinnedBuf table_pin;                  // pinned host mirror of the chunk table + tile_slot_ptr rows (read by copy_words_kernel)
    DevBuf csc_keys, csc_keys2, csc_payload, tile_hist, tile_slot_ptr;
    DevBuf csc_tmp;                       // CUB temp storage of the CSC sorts: they may run on the copy stream while the
                                          // engine stream sorts something else (labels) with sort_tmp
    std::vector<cudaEvent_t> ev_tile;     // tile t: CSC sorted and histogram on the host
    std::vector<FastChunk> h_chunks;      // tile-major; inside a tile longest first
    std::vector<int32_t> h_tile_chunk_ptr, h_tile_slot_ptr, h_tile_entry0;

    // staged (caller's full) matrix, between lrperm_stage_matrix_csr and lrperm_commit_matrix
    bool staged = false;
    int64_t st_N = 0, st_G = 0, st_nnz = 0, st_unordered = 0;
    bool st_has_keep = false;
    DevBuf st_indptr, st_indices, st_data, st_row_keep, st_small;

    // labels
    int32_t L = 0;
    bool have_labels = false;
    DevBuf labels32, labels8, labels8s, order, rank, bucket_cluster, cl_start, inv_n;
    int32_t n_buckets = 0, bucket_shift = 0;      // coarse position -> cluster table of slab_philox_kernel
    int32_t labels8s_scale = 0;          // scale the labels8s table currently holds (0 = stale)
    std::vector<int32_t> h_cl_start;

    // triples
    int64_t T = 0;
    bool have_triples = false, have_thr_c = false, have_thr_g = false;
    DevBuf rowA, rowB, thr_c, thr_g, counts_c, counts_g;

    // FAST-mode chunk tables
    int32_t fast_perm_batch = 0, fast_chunk_nnz = 0;   // 0 = auto
    int32_t chunk_table_nnz = -1;                       // chunk length the chunk tables were built for
    int32_t fast_tile_bytes = 0;                        // label-slab bytes one cell tile may span (0 = auto)
    int32_t n_sched = 0;                                // batches of the last pipelined run
    int32_t fast_variant = 2;                           // 1 = fast_means_kernel, 2 = fast2_means_kernel
    int32_t fast_q = 0;                                 // permutations per lane (0 = auto)
    int32_t fast_first_batch = 0;                       // permutations of a short first batch (0 = none; A/B only)
    int32_t n_chunks = 0, n_slots = 0;
    DevBuf chunks;
    // Genes no triple reads are not accumulated when no cube is exported (option "skip_unreferenced"): the chunk
    // table keeps ALL chunks and their slot numbers (so the summation order of every referenced gene, and with it every
    // result, is unchanged), `chunks_f` lists the chunks of the referenced genes only, tile-major like `chunks`.
    int32_t slab_mode = 2;                // option (A/B): which slab launches use the persistent grid: 1 all, 2 all but batch 0 / tile 0, 3 batches >= 1, 0 none
    int32_t slab_ctas_per_sm = 0;         // option: resident CTAs per SM of the label-slab kernel in the pipelined FAST path (0 = full grid)
    int32_t narrow_idx = 1;               // option "narrow_indices": 16-bit column indices on the wire for large pageable stage uploads
    int32_t skip_unref = 1;
    std::vector<uint8_t> h_needed;                      // [G] 1 = some triple reads the gene (from lrperm_set_triples)
    DevBuf needed_dev;
    bool needed_valid = false;
    int64_t needed_version = 0, chunk_table_id = 0;
    int64_t chunks_f_needed = -1, chunks_f_table = -1;  // versions the filtered list was built for
    int32_t chunks_f_tiles = 0;
    DevBuf chunks_f;
    std::vector<FastChunk> h_chunks_f;
    std::vector<int32_t> h_tile_chunk_ptr_f;
    bool filter_tables = false;                         // chunks_append_tile maintains the filtered list too (set by the run)
    int64_t stat_genes_scored = 0, stat_nnz_scored = 0, stat_chunks_run = 0;   // of the last lrperm_run

    // work buffers
    DevBuf slab2, partials2;              // second halves of the double buffers of the pipelined FAST path
    DevBuf perm_dev, slab, partials, cube, sort_tmp, tmp_a, tmp_b, tmp_c, tmp_d, stage_out, small_a, small_b;

    // CellChat trimean path (lrperm_trimean.cuh)
    // lrperm_resolve_triples: side tables and the resolved rows (kept until lrperm_fetch_triples)
    DevBuf rs_side, rs_flag, rs_pos, rs_rows, rs_in;
    int64_t rs_T = -1;
    int32_t rs_f64 = 0;
    DevBuf tm_diag, tm_chunk_of_slot, tm_region_dir, tm_walk_flag, tm_walk_list;
    int64_t tm_select_warps = 0, tm_walked_warps = 0;
    int32_t tm_n_batches = 0, tm_n_super = 0;
    DevBuf vsc_pairs, tm_chunks, tm_region_ptr, tm_counts, tm_prefix, tm_planes, tm_cube, tm_tgt, tm_roles, thr_d, counts_t;
    bool vsc_ready = false, tm_tables_ready = false;
    int32_t tm_chunk_nnz = 0, tm_perm_batch = 0;        // tuning (0 = auto)
    int32_t tm_table_nnz = -1;                           // chunk size the trimean chunk table was built for
    int64_t tm_table_needed = 0;                         // needed_version it was filtered by (0 = all genes)
    int32_t tm_n_chunks = 0, tm_n_slots = 0, tm_n_regions = 0;
    float tm_ms[5] = {0, 0, 0, 0, 0};                    // count, prefix, select, finalize, select-warps-that-walked (n/a)

    // timings of the last run
    float t_setup = 0, t_means = 0, t_score = 0, t_total = 0, t_main = 0;
    int32_t n_main_launches = 0, last_pb = 0;
    int64_t launches = 0;
    std::vector<cudaEvent_t> events;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CK(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return h->fail(_e == cudaErrorMemoryAllocation ? LRPERM_ERR_OOM : LRPERM_ERR_CUDA,    \
                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define CK_LAUNCH()                                                                               \
    do {                                                                                          \
        h->launches++;                                                                            \
        CK(cudaGetLastError());                                                                   \
    } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// extents for the kernels' own bounds checks (debug build only: -DLRPERM_BOUNDS_CHECK); queued on the engine stream
int dbg_set_dims(lrperm_engine* h, int64_t n_slots, int64_t PB) {
#ifdef LRPERM_BOUNDS_CHECK
    DebugDims d{(long long)h->N, (long long)h->G, (long long)h->L, (long long)h->nnz, (long long)n_slots, (long long)PB, (long long)h->T};
    CK(cudaMemcpyToSymbolAsync(g_dbg_dims, &d, sizeof(d), 0, cudaMemcpyHostToDevice, h->stream));
#else
    (void)h; (void)n_slots; (void)PB;
#endif
    return LRPERM_OK;
}

int grid_for(int64_t n, int threads, int cap = 148 * 16) {
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Large upload from PAGEABLE host memory (a NumPy array): the driver's own staging moves ~10 GB/s; here host threads
// fill one pinned piece while the previous one is on the wire.
int upload_pageable(lrperm_engine* h, void* dst, const void* src, size_t bytes, bool narrow = false, uint32_t* bad_out = nullptr) {
    if (!h->copier) {
        const unsigned hw = std::thread::hardware_concurrency();
        // host threads that fill the staging pieces: measured on the 16-core B200 host (scripts/upload_probe.py)
        unsigned n_threads = std::max(2u, std::min(8u, hw ? hw / 2 : 2u));
        if (const char* e = std::getenv("LRPERM_COPY_THREADS")) n_threads = (unsigned)std::max(1, std::min(64, atoi(e)));
        if (const char* e = std::getenv("LRPERM_PIN_PIECES")) h->n_pin = std::max(2, std::min((int)lrperm_engine::kMaxPin, atoi(e)));
        if (const char* e = std::getenv("LRPERM_PIN_PIECE_MB")) h->pin_piece = (size_t)std::max(1, std::min(1024, atoi(e))) << 20;
        h->copier = new (std::nothrow) ParallelCopier((int)n_threads);
        if (!h->copier) return h->fail(LRPERM_ERR_OOM, "host allocation failed");
        for (int i = 0; i < h->n_pin; ++i) {
            CK(cudaHostAlloc(&h->pin[i], h->pin_piece, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&h->pin_ev[i], cudaEventDisableTiming));
        }
    }
    // narrow: `bytes` counts the uint16 OUTPUT; the int32 source is read at twice the offset
    const unsigned char* s = static_cast<const unsigned char*>(src);
    unsigned char* d = static_cast<unsigned char*>(dst);
    const int np = h->n_pin;
    uint32_t bad = 0;
    int k = 0;
    for (size_t o = 0; o < bytes; o += h->pin_piece, ++k) {
        const size_t n = std::min(h->pin_piece, bytes - o);
        const int b = k % np;
        if (k >= np) CK(cudaEventSynchronize(h->pin_ev[b]));     // the DMA that read this piece one rotation ago is done
        bad |= h->copier->copy(h->pin[b], s + (narrow ? 2 * o : o), n, narrow);
        CK(cudaMemcpyAsync(d + o, h->pin[b], n, cudaMemcpyHostToDevice, h->stream));
        CK(cudaEventRecord(h->pin_ev[b], h->stream));
    }
    // the staging pieces are reused by the next upload: leave none in flight
    for (int b = 0; b < np && b < k; ++b) CK(cudaEventSynchronize(h->pin_ev[b]));
    if (bad_out) *bad_out = bad;
    return LRPERM_OK;
}

// Host -> device copy that bypasses the copy engine: while a pinned matrix upload is pending its tiles fill the H2D
// queue (copies of all streams are served in submission order), and the label / triple tables of the calls that
// follow would arrive only after the last tile.  A kernel reads them over PCIe instead - straight from the caller's
// buffer when it is pinned, else from a pinned mirror filled here.
int upload_by_kernel(lrperm_engine* h, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return LRPERM_OK;
    const unsigned char* from = static_cast<const unsigned char*>(src);
    if (!is_pinned_host(src) || (reinterpret_cast<uintptr_t>(src) & 3)) {
        const size_t need = h->up_pin_off + ((bytes + 255) & ~(size_t)255);
        if (need > h->up_pin.cap) {
            // a retired block may still be read by queued kernels: PinnedBuf keeps it; start a fresh one
            CK(h->up_pin.ensure(std::max(need, (size_t)(32u << 20))));
            h->up_pin_off = 0;
        }
        unsigned char* pin = h->up_pin.as<unsigned char>() + h->up_pin_off;
        memcpy(pin, src, bytes);
        h->up_pin_off += (bytes + 255) & ~(size_t)255;
        from = pin;
    }
    copy_bytes_kernel<<<grid_for((int64_t)(bytes >> 2) + 1, 256, h->sm_count * 4), 256, 0, st>>>(from, static_cast<unsigned char*>(dst), (int64_t)bytes);
    CK_LAUNCH();
    return LRPERM_OK;
}

// copy `bytes` from a host-or-device source into a device buffer on the engine stream
int upload(lrperm_engine* h, DevBuf& dst, const void* src, size_t bytes, int on_device) {
    CK(dst.ensure(bytes));
    if (bytes == 0) return LRPERM_OK;
    if (!on_device && h->matrix_pending && bytes <= (256u << 20)) return upload_by_kernel(h, dst.p, src, bytes, h->stream);
    if (!on_device && bytes >= (32u << 20) && !is_pinned_host(src)) return upload_pageable(h, dst.p, src, bytes);
    CK(cudaMemcpyAsync(dst.p, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    return LRPERM_OK;
}

int download(lrperm_engine* h, void* dst, const void* src_dev, size_t bytes, int on_device) {
    if (bytes == 0 || dst == nullptr) return LRPERM_OK;
    CK(cudaMemcpyAsync(dst, src_dev, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    return LRPERM_OK;
}

int check_err_flag(lrperm_engine* h, const char* what) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (flag) {
        CK(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
        return h->fail(LRPERM_ERR_INVALID, "%s: invalid input (flag=%d: 1=column out of range, 2=indptr overflow, "
                       "4=indptr not monotone, 8=triple index out of range)", what, flag);
    }
    return LRPERM_OK;
}

// ---- CSC copy + chunk tables for FAST mode, one cell tile at a time ---------------------------
int32_t fast_tile_cells(const lrperm_engine* h) {
    // cells per tile: the slab rows of one tile (tile_cells x PB bytes) should sit in L2
    // (measured on B200, scripts/perf_tune.py: 131,072-cell tiles - 64 MiB of slab at 512 permutations, 16 MiB
    // at the 125 permutations one of 8 GPUs runs - beat 32 MiB by 1% / 6% and 16 MiB by 3% on configs[2])
    const int64_t tile_bytes = h->fast_tile_bytes ? h->fast_tile_bytes : (64LL << 20);
    // The tile (and with it the chunk table, i.e. the summation order of FAST mode) is defined for the
    // nominal batch of 512 permutations whatever batch sizes are run: FAST results do not depend on the
    // batch schedule, so they are bit-identical for any number of GPUs / permutation shards.
    return (int32_t)std::min<int64_t>(std::max<int64_t>(tile_bytes / 512, 1024), std::max<int64_t>(h->N, 1));
}

// (re)start the tile-wise CSC build for `tile_cells`: buffers, per-tile events, pinned histogram
int csc_begin(lrperm_engine* h, int32_t tile_cells) {
    if (h->csc_ready && h->csc_tile_cells == tile_cells) return LRPERM_OK;
    const int32_t G = (int32_t)h->G;
    const size_t n1 = (size_t)std::max<int64_t>(h->nnz, 1);
    h->csc_tile_cells = tile_cells;
    h->n_tiles = (int32_t)std::max<int64_t>(1, (h->N + tile_cells - 1) / tile_cells);
    h->tiles_enqueued = 0;
    h->tiles_tabled = 0;
    h->chunk_table_nnz = -1;
    CK(h->csc_keys.ensure(sizeof(uint32_t) * n1));                // column keys
    CK(h->csc_keys2.ensure(sizeof(uint32_t) * n1));               // sorted keys (discarded)
    CK(h->csc_payload.ensure(sizeof(unsigned long long) * n1));   // payload {cell, value bits}
    // + 64 zero entries behind the last one: fast3_means_kernel (bulk copies of whole 32-entry blocks) may read past the end
    CK(h->csc_pairs.ensure(sizeof(int2) * (n1 + 64)));
    CK(cudaMemsetAsync(h->csc_pairs.as<int2>() + n1, 0, sizeof(int2) * 64, h->stream));
    const size_t hist_elems = (size_t)h->n_tiles * (size_t)G;
    CK(h->tile_hist.ensure(sizeof(int32_t) * hist_elems));
    CK(h->hist_pin.ensure(sizeof(int32_t) * hist_elems));
    while (h->ev_tile.size() < (size_t)h->n_tiles) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->ev_tile.push_back(ev);
    }
    // one temp buffer for the largest tile's sort
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while ((1 << end_bit) < G && end_bit < 31) ++end_bit;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->csc_keys.as<uint32_t>(), h->csc_keys2.as<uint32_t>(),
                                       h->csc_payload.as<unsigned long long>(), reinterpret_cast<unsigned long long*>(h->csc_pairs.p),
                                       (int)std::min<int64_t>(h->nnz, 0x7fffffff), 0, end_bit, h->stream));
    CK(h->csc_tmp.ensure(tmp_bytes));
    h->csc_ready = true;
    return LRPERM_OK;
}

// queue the CSC build of tile t on `st`: keys + histogram, stable radix sort by column (cells stay ascending inside a
// gene), histogram to the pinned host buffer, event.  [e0, e1) = the tile's entry range (host-known).
int csc_enqueue_tile(lrperm_engine* h, int32_t t, int64_t e0, int64_t e1, cudaStream_t st, bool from_raw = false) {
    const int32_t G = (int32_t)h->G;
    const int32_t row0 = (int32_t)std::min<int64_t>((int64_t)t * h->csc_tile_cells, h->N);
    const int32_t row1 = (int32_t)std::min<int64_t>((int64_t)(t + 1) * h->csc_tile_cells, h->N);
    int32_t* d_hist = h->tile_hist.as<int32_t>() + (size_t)t * G;
    CK(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (size_t)G, st));
    if (e1 > e0) {
        size_t smem = sizeof(int32_t) * (size_t)G;
        const int use_shared = smem <= 96 * 1024;
        if (!use_shared) smem = 0;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(csc_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)std::min<int64_t>((int64_t)h->sm_count * 4, std::max<int64_t>(1, ((int64_t)(row1 - row0) * 32 + 511) / 512));
        if (from_raw)       // pinned upload: straight from the staged raw arrays, no packed copy yet
            csc_keys_kernel<<<grid, 512, smem, st>>>(h->indptr.as<int32_t>(), nullptr, row0, row1, G,
                                                     h->csc_keys.as<uint32_t>(), h->csc_payload.as<unsigned long long>(), d_hist, use_shared,
                                                     h->st_indices.as<int32_t>(), h->st_data.as<float>(), h->err_flag_m.as<int>());
        else
            csc_keys_kernel<<<grid, 512, smem, st>>>(h->indptr.as<int32_t>(), h->csr_pairs.as<int2>(), row0, row1, G,
                                                     h->csc_keys.as<uint32_t>(), h->csc_payload.as<unsigned long long>(), d_hist, use_shared);
        CK_LAUNCH();
        int end_bit = 1;
        while ((1 << end_bit) < G && end_bit < 31) ++end_bit;
        size_t tmp_bytes = h->csc_tmp.cap;
        CK(cub::DeviceRadixSort::SortPairs(h->csc_tmp.p, tmp_bytes, h->csc_keys.as<uint32_t>() + e0, h->csc_keys2.as<uint32_t>() + e0,
                                           h->csc_payload.as<unsigned long long>() + e0,
                                           reinterpret_cast<unsigned long long*>(h->csc_pairs.p) + e0, (int)(e1 - e0), 0, end_bit, st));
        h->launches += 4;
    }
    CK(cudaMemcpyAsync(h->hist_pin.as<int32_t>() + (size_t)t * G, d_hist, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev_tile[(size_t)t], st));
    return LRPERM_OK;
}

// submit the transfer + packing + CSC sort of the next `count` tiles of a pending pinned upload (count < 0: all that
// are left) to the copy stream; after the last one ev_matrix marks "the whole matrix is on the device"
int submit_tiles(lrperm_engine* h, int32_t count) {
    cudaStream_t sC = h->copy_stream;
    int rc;
    const int32_t end = count < 0 ? h->n_tiles : std::min(h->n_tiles, h->up_next_tile + count);
    for (int32_t t = h->up_next_tile; t < end; ++t) {
        const int64_t e0 = h->up_e[(size_t)t], e1 = h->up_e[(size_t)t + 1];
        const int32_t row0 = (int32_t)std::min<int64_t>((int64_t)t * h->csc_tile_cells, h->N);
        const int32_t row1 = (int32_t)std::min<int64_t>((int64_t)(t + 1) * h->csc_tile_cells, h->N);
        if (e1 > e0) {
            CK(cudaMemcpyAsync(h->st_indices.as<int32_t>() + e0, static_cast<const int32_t*>(h->up_indices) + e0, sizeof(int32_t) * (size_t)(e1 - e0), cudaMemcpyHostToDevice, sC));
            CK(cudaMemcpyAsync(h->st_data.as<float>() + e0, static_cast<const float*>(h->up_data) + e0, sizeof(float) * (size_t)(e1 - e0), cudaMemcpyHostToDevice, sC));
            // no packing here: FAST mode builds its CSC copy from the raw arrays; the bank-dealt csr_pairs of the
            // row-gather kernels (EXACT order, observed statistics, moments, trimean) is built on first use (wait_matrix)
        }
        (void)row0; (void)row1;
        if ((rc = csc_enqueue_tile(h, t, e0, e1, sC, true))) return rc;
        h->tiles_enqueued = t + 1;
    }
    h->up_next_tile = end;
    if (end == h->n_tiles) CK(cudaEventRecord(h->ev_matrix, sC));
    return LRPERM_OK;
}

// A matrix uploaded from pinned host memory may still be in flight on the copy stream (lrperm_set_matrix_csr returned
// without waiting): submit what is left of it, make the engine stream wait, surface the deferred validation result.
// Everything that reads the whole matrix on the device calls this first; what does not need it (labels, triples,
// the label slabs of the first batches) proceeds while the upload runs.
// need_csr: the caller reads csr_pairs (row-gather kernels).  After a pinned upload that copy does not exist yet (FAST
// mode never needs it): it is packed here, once, from the raw arrays the upload left in the staging buffers.
int wait_matrix(lrperm_engine* h, bool need_csr = true) {
    if (h->matrix_pending) {
        int rc;
        if ((rc = submit_tiles(h, -1))) return rc;
        h->matrix_pending = false;
        CK(cudaStreamWaitEvent(h->stream, h->ev_matrix, 0));
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, h->err_flag_m.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (flag) {
            CK(cudaMemsetAsync(h->err_flag_m.p, 0, sizeof(int), h->stream));
            h->have_matrix = false;
            return h->fail(LRPERM_ERR_INVALID, "lrperm_set_matrix_csr (deferred check): invalid input (flag=%d: 1=column out of range, "
                           "2=indptr overflow, 4=indptr not monotone)", flag);
        }
    }
    if (need_csr && !h->csr_packed) {
        if (h->nnz > 0) {
            pack_csr_kernel<<<grid_for(h->N * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
                h->indptr.as<int32_t>(), h->st_indices.as<int32_t>(), h->st_data.as<float>(), 0, (int32_t)h->N, (int32_t)h->G,
                h->csr_pairs.as<int2>(), h->err_flag.as<int>());
            CK_LAUNCH();
        }
        h->csr_packed = true;
    }
    return LRPERM_OK;
}

// chunk table of tile t (host, after its histogram arrived): every gene's entries inside the tile, cut every
// chunk_nnz; inside the tile the longest chunks first (short ones fill the tail); slots numbered tile by tile
int chunks_append_tile(lrperm_engine* h, int32_t t, int32_t chunk_nnz, cudaStream_t st) {
    const int32_t G = (int32_t)h->G;
    CK(cudaEventSynchronize(h->ev_tile[(size_t)t]));
    if (t == 0) {
        h->h_chunks.clear();
        h->h_tile_chunk_ptr.assign(1, 0);
        h->h_tile_slot_ptr.assign((size_t)h->n_tiles * ((size_t)G + 1), 0);
        h->h_tile_entry0.assign(1, 0);
        h->n_slots = 0;
        ++h->chunk_table_id;
        h->h_chunks_f.clear();
        h->h_tile_chunk_ptr_f.assign(1, 0);
        h->chunks_f_tiles = 0;
        // device tables sized for any tiling of this matrix
        const size_t max_chunks = (size_t)(h->nnz / std::max(chunk_nnz, 1)) + (size_t)G * (size_t)h->n_tiles + 1;
        CK(h->chunks.ensure(sizeof(FastChunk) * max_chunks));
        CK(h->tile_slot_ptr.ensure(sizeof(int32_t) * h->h_tile_slot_ptr.size()));
        const size_t words = 2 * max_chunks * (sizeof(FastChunk) / 4) + h->h_tile_slot_ptr.size();
        CK(h->table_pin.ensure(sizeof(uint32_t) * words));      // an outgrown mirror is retired, not freed: nothing in flight loses it
        CK(h->chunks_f.ensure(sizeof(FastChunk) * max_chunks));
    }
    const int32_t* hist = h->hist_pin.as<int32_t>() + (size_t)t * G;
    int32_t pos = h->h_tile_entry0[(size_t)t];
    int32_t slot = h->n_slots;
    const size_t first = h->h_chunks.size();
    int32_t* tsp = h->h_tile_slot_ptr.data() + (size_t)t * ((size_t)G + 1);
    for (int32_t g = 0; g < G; ++g) {
        tsp[g] = slot;
        const int32_t e = pos + hist[g];
        for (int32_t s0 = pos; s0 < e; s0 += chunk_nnz) {
            FastChunk c;
            c.gene = g; c.begin = s0; c.end = std::min(e, s0 + chunk_nnz); c.slot = slot++;
            h->h_chunks.push_back(c);
        }
        pos = e;
    }
    tsp[G] = slot;
    std::stable_sort(h->h_chunks.begin() + (std::ptrdiff_t)first, h->h_chunks.end(),
                     [](const FastChunk& a, const FastChunk& b) { return (a.end - a.begin) > (b.end - b.begin); });
    h->h_tile_entry0.push_back(pos);
    h->h_tile_chunk_ptr.push_back((int32_t)h->h_chunks.size());
    h->n_slots = slot;
    h->n_chunks = (int32_t)h->h_chunks.size();
    // to the device through the pinned mirror and a copy KERNEL (not the copy engine: see copy_words_kernel); the mirror
    // has a distinct region per tile, so nothing in flight is overwritten
    const size_t n_new = h->h_chunks.size() - first;
    const size_t cw = sizeof(FastChunk) / 4;
    const size_t max_chunks = (size_t)(h->nnz / std::max(chunk_nnz, 1)) + (size_t)G * (size_t)h->n_tiles + 1;
    if (n_new) {
        uint32_t* pin = h->table_pin.as<uint32_t>() + first * cw;
        memcpy(pin, h->h_chunks.data() + first, sizeof(FastChunk) * n_new);
        copy_words_kernel<<<grid_for((int64_t)(n_new * cw), 256, 256), 256, 0, st>>>(pin, reinterpret_cast<uint32_t*>(h->chunks.as<FastChunk>() + first), (int64_t)(n_new * cw));
        CK_LAUNCH();
    }
    {
        uint32_t* pin = h->table_pin.as<uint32_t>() + max_chunks * cw + (size_t)t * ((size_t)G + 1);
        memcpy(pin, tsp, sizeof(int32_t) * ((size_t)G + 1));
        copy_words_kernel<<<grid_for((int64_t)G + 1, 256, 64), 256, 0, st>>>(pin, reinterpret_cast<uint32_t*>(h->tile_slot_ptr.as<int32_t>() + (size_t)t * ((size_t)G + 1)), (int64_t)G + 1);
        CK_LAUNCH();
    }
    if (h->filter_tables && h->chunks_f_tiles == t) {
        // the tile's chunks of referenced genes, same order; third region of the pinned mirror
        const size_t first_f = h->h_chunks_f.size();
        for (size_t i = first; i < h->h_chunks.size(); ++i)
            if (h->h_needed[(size_t)h->h_chunks[i].gene]) h->h_chunks_f.push_back(h->h_chunks[i]);
        h->h_tile_chunk_ptr_f.push_back((int32_t)h->h_chunks_f.size());
        const size_t n_f = h->h_chunks_f.size() - first_f;
        if (n_f) {
            uint32_t* pin = h->table_pin.as<uint32_t>() + max_chunks * cw + h->h_tile_slot_ptr.size() + first_f * cw;
            memcpy(pin, h->h_chunks_f.data() + first_f, sizeof(FastChunk) * n_f);
            copy_words_kernel<<<grid_for((int64_t)(n_f * cw), 256, 256), 256, 0, st>>>(pin, reinterpret_cast<uint32_t*>(h->chunks_f.as<FastChunk>() + first_f), (int64_t)(n_f * cw));
            CK_LAUNCH();
        }
        h->chunks_f_tiles = t + 1;
        h->chunks_f_needed = h->needed_version;
        h->chunks_f_table = h->chunk_table_id;
    }
    h->tiles_tabled = t + 1;
    return LRPERM_OK;
}

// filtered chunk list (referenced genes only) of a chunk table that is complete on the host
int ensure_filtered_chunks(lrperm_engine* h) {
    if (h->chunks_f_table == h->chunk_table_id && h->chunks_f_needed == h->needed_version && h->chunks_f_tiles == h->n_tiles) return LRPERM_OK;
    h->h_chunks_f.clear();
    h->h_tile_chunk_ptr_f.assign(1, 0);
    for (int32_t t = 0; t < h->n_tiles; ++t) {
        for (int32_t i = h->h_tile_chunk_ptr[(size_t)t]; i < h->h_tile_chunk_ptr[(size_t)t + 1]; ++i)
            if (h->h_needed[(size_t)h->h_chunks[(size_t)i].gene]) h->h_chunks_f.push_back(h->h_chunks[(size_t)i]);
        h->h_tile_chunk_ptr_f.push_back((int32_t)h->h_chunks_f.size());
    }
    CK(h->chunks_f.ensure(sizeof(FastChunk) * std::max<size_t>(h->h_chunks_f.size(), 1)));
    if (!h->h_chunks_f.empty())     // pageable source: the call returns once the data has been staged
        CK(cudaMemcpyAsync(h->chunks_f.p, h->h_chunks_f.data(), sizeof(FastChunk) * h->h_chunks_f.size(), cudaMemcpyHostToDevice, h->stream));
    h->chunks_f_table = h->chunk_table_id;
    h->chunks_f_needed = h->needed_version;
    h->chunks_f_tiles = h->n_tiles;
    return LRPERM_OK;
}

// work statistics of a FAST run: genes / stored entries / chunk warps actually accumulated
void fast_run_stats(lrperm_engine* h, bool filt) {
    const int32_t G = (int32_t)h->G;
    int64_t genes = 0, nnz = 0;
    const int32_t* hist = h->hist_pin.as<int32_t>();
    for (int32_t g = 0; g < G; ++g) {
        if (filt && !h->h_needed[(size_t)g]) continue;
        ++genes;
        if (hist) for (int32_t t = 0; t < h->n_tiles; ++t) nnz += hist[(size_t)t * G + g];
    }
    h->stat_genes_scored = genes;
    h->stat_nnz_scored = nnz;
    h->stat_chunks_run = filt ? (int64_t)h->h_chunks_f.size() : (int64_t)h->h_chunks.size();
}

// everything at once (matrix resident): CSC of all tiles, then all chunk tables
// queue the CSC build of every tile of the resident matrix on the engine stream (no wait for the result)
int csc_enqueue_all(lrperm_engine* h, int32_t tile_cells) {
    if (h->csc_ready && h->csc_tile_cells == tile_cells) return LRPERM_OK;
    StageTimer timer("build_csc");
    int rc;
    const bool raw = !h->csr_packed;                        // no packed copy yet: sort from the raw arrays in the staging buffers
    if ((rc = wait_matrix(h, false))) return rc;
    if ((rc = csc_begin(h, tile_cells))) return rc;
    // entry ranges of the tiles: indptr at the tile boundaries
    std::vector<int32_t> e((size_t)h->n_tiles + 1, 0);
    for (int32_t t = 0; t <= h->n_tiles; ++t) {
        const int64_t row = std::min<int64_t>((int64_t)t * tile_cells, h->N);
        CK(cudaMemcpyAsync(&e[(size_t)t], h->indptr.as<int32_t>() + row, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    for (int32_t t = 0; t < h->n_tiles; ++t)
        if ((rc = csc_enqueue_tile(h, t, e[(size_t)t], e[(size_t)t + 1], h->stream, raw))) return rc;
    h->tiles_enqueued = h->n_tiles;
    return LRPERM_OK;
}

int ensure_fast_tables(lrperm_engine* h, int32_t chunk_nnz, int32_t tile_cells) {
    int rc;
    if ((rc = csc_enqueue_all(h, tile_cells))) return rc;
    if (h->tiles_enqueued == h->n_tiles && (h->chunk_table_nnz != chunk_nnz || h->tiles_tabled < h->n_tiles)) {
        StageTimer timer("build_chunks");
        const int32_t from = h->chunk_table_nnz == chunk_nnz ? h->tiles_tabled : 0;
        for (int32_t t = from; t < h->n_tiles; ++t)
            if ((rc = chunks_append_tile(h, t, chunk_nnz, h->stream))) return rc;
        h->chunk_table_nnz = chunk_nnz;
    }
    return LRPERM_OK;
}

template <int W>
int launch_exact_w(lrperm_engine* h, const PermSource& ps, int32_t pb, float* cube, int32_t PP, size_t smem) {
    const int32_t G = (int32_t)h->G, L = h->L;
    CK(cudaFuncSetAttribute(exact_means_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t items = (int64_t)pb * L;
    const int64_t grid = (items + W - 1) / W;
    if (grid > 0) {
        exact_means_kernel<W><<<(unsigned)grid, W * 32, smem, h->stream>>>(
            h->indptr.as<int32_t>(), h->csr_pairs.as<int2>(), h->order.as<int32_t>(), h->cl_start.as<int32_t>(),
            h->inv_n.as<float>(), ps, (int32_t)h->N, G, L, pb, cube, PP);
        CK_LAUNCH();
    }
    return LRPERM_OK;
}

int launch_exact(lrperm_engine* h, const PermSource& ps, int32_t pb, float* cube, int32_t PP) {
    const int32_t G = (int32_t)h->G;
    const int Gpad = ((G + 31) & ~31) + 32;      // + one dummy slot per lane (see exact_means_kernel)
    const size_t per_warp = (size_t)Gpad * sizeof(float);
    // warps per CTA: maximise resident warps per SM under the shared-memory budget
    // (1 KB per CTA is reserved by the driver) and the 2048-thread limit
    const size_t sm_budget = 227 * 1024;
    int best_w = 0, best_res = 0;
    for (int w : {16, 12, 10, 9, 8, 7, 6, 4, 2, 1}) {
        const size_t cta = per_warp * w + 1024;
        if (per_warp * w > h->smem_optin) continue;
        int ctas = (int)std::min<size_t>(sm_budget / cta, 32);
        ctas = std::min(ctas, 2048 / (w * 32));
        const int res = std::min(ctas * w, 28);       // 72 registers/thread (launch bounds) cap residency at 28 warps
        if (res > best_res) { best_res = res; best_w = w; }
    }
    if (best_w == 0)
        return h->fail(LRPERM_ERR_UNSUPPORTED, "EXACT mode needs %zu B shared memory per warp (G=%d); limit %zu", per_warp, G, h->smem_optin);
    const size_t smem = per_warp * best_w;
    switch (best_w) {
        case 16: return launch_exact_w<16>(h, ps, pb, cube, PP, smem);
        case 12: return launch_exact_w<12>(h, ps, pb, cube, PP, smem);
        case 10: return launch_exact_w<10>(h, ps, pb, cube, PP, smem);
        case 9: return launch_exact_w<9>(h, ps, pb, cube, PP, smem);
        case 8: return launch_exact_w<8>(h, ps, pb, cube, PP, smem);
        case 7: return launch_exact_w<7>(h, ps, pb, cube, PP, smem);
        case 6: return launch_exact_w<6>(h, ps, pb, cube, PP, smem);
        case 4: return launch_exact_w<4>(h, ps, pb, cube, PP, smem);
        case 2: return launch_exact_w<2>(h, ps, pb, cube, PP, smem);
        default: return launch_exact_w<1>(h, ps, pb, cube, PP, smem);
    }
}

template <int Q>
int launch_fast_q(lrperm_engine* h, int32_t PB, const uint8_t* slab, float* partials, const FastChunk* table, int32_t c0, int32_t cn) {
    const int32_t L = h->L;
    const size_t smem = (size_t)kFastWarps * L * Q * 32 * sizeof(float);
    CK(cudaFuncSetAttribute(fast_means_kernel<Q, kFastWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int32_t groups = PB / (32 * Q);
    const int64_t items = (int64_t)cn * groups;
    const int64_t grid = (items + kFastWarps - 1) / kFastWarps;
    if (grid > 0) {
        fast_means_kernel<Q, kFastWarps><<<(unsigned)grid, kFastWarps * 32, smem, h->stream>>>(
            h->csc_pairs.as<int2>(), table + c0, cn, slab, PB, L, groups, partials);
        CK_LAUNCH();
    }
    return LRPERM_OK;
}

template <int Q, int B>
int launch_fast2_q(lrperm_engine* h, int32_t PB, const uint8_t* slab, float* partials, const FastChunk* table, int32_t c0, int32_t cn) {
    const int32_t L = h->L;
    const size_t smem = (size_t)L * Q * 128 + 16 * B;
    CK(cudaFuncSetAttribute(fast2_means_kernel<Q, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int32_t groups = PB / (32 * Q);
    const int64_t grid = (int64_t)cn * groups;
    if (grid > 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "FAST mode: too many chunks (%lld CTAs)", (long long)grid);
    if (grid > 0) {
        fast2_means_kernel<Q, B><<<(unsigned)grid, 32, smem, h->stream>>>(
            h->csc_pairs.as<int2>(), table + c0, cn, slab, PB, L, groups, partials);
        CK_LAUNCH();
    }
    return LRPERM_OK;
}

template <int Q, int R>
int launch_fast3_q(lrperm_engine* h, int32_t PB, const uint8_t* slab, float* partials, const FastChunk* table, int32_t c0, int32_t cn) {
    const int32_t L = h->L;
    const size_t smem = (size_t)L * Q * 128 + (size_t)R * 32 * sizeof(int2) + (size_t)R * 8;
    CK(cudaFuncSetAttribute(fast3_means_kernel<Q, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int32_t groups = PB / (32 * Q);
    const int64_t grid = (int64_t)cn * groups;
    if (grid > 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "FAST mode: too many chunks (%lld CTAs)", (long long)grid);
    if (grid > 0) {
        fast3_means_kernel<Q, R><<<(unsigned)grid, 32, smem, h->stream>>>(
            h->csc_pairs.as<int2>(), table + c0, cn, slab, PB, L, groups, partials);
        CK_LAUNCH();
    }
    return LRPERM_OK;
}

// v2 kernel: permutations per lane.  The slab byte is label*Q/2 (must fit 8 bits) and the bins take
// L*Q*128 B of shared memory per warp; more, smaller CTAs hide the LDS->FADD->STS chain better.
int pick_q2(const lrperm_engine* h) {
    const size_t budget = h->smem_optin;
    if (h->fast_q) {
        const int q = h->fast_q;
        if ((q == 2 || q == 4 || q == 8) && (h->L - 1) * q / 2 <= 255 && (size_t)h->L * q * 128 + 512 <= budget) return q;
        return 0;
    }
    for (int q : {4, 2}) {
        if ((h->L - 1) * q / 2 > 255) continue;
        if ((size_t)h->L * q * 128 + 512 <= budget) return q;
    }
    return 0;
}

int pick_q(const lrperm_engine* h) {
    // keep >= 1 CTA of kFastWarps warps resident: bins = warps * L * Q * 128 B
    const size_t budget = h->smem_optin;
    for (int q : {4, 2, 1}) {
        const size_t smem = (size_t)kFastWarps * h->L * q * 32 * sizeof(float);
        if (smem <= budget) return q;
    }
    return 0;
}

int32_t score_groups_per_block(const lrperm_engine* h, int32_t n_groups, size_t elem) {
    if (h->score_gpb > 0) return std::min(h->score_gpb, std::max(n_groups, 1));
    const size_t per_group = (size_t)h->G * (size_t)h->L * 32 * elem;
    const size_t fit = (48u << 20) / std::max<size_t>(per_group, 1);
    return (int32_t)std::max<size_t>(1, std::min<size_t>(fit, (size_t)std::max(n_groups, 1)));
}

template <bool LIT>
int launch_score_t(lrperm_engine* h, int scores, const float* cube, int32_t PP, int32_t pb) {
    const int64_t T = h->T;
    if (T == 0) return LRPERM_OK;
    const bool c = scores & LRPERM_SCORE_CPDB, g = scores & LRPERM_SCORE_GMEAN;
    // permutation groups per block: the gathered cube slice (G x L rows x 128 B per group) should fit L2 (~48 MiB)
    const int32_t n_groups = (pb + 31) / 32;
    const int32_t gpb = score_groups_per_block(h, n_groups, sizeof(float));
    dim3 grid((unsigned)((T + kScoreTile - 1) / kScoreTile), (unsigned)((n_groups + gpb - 1) / gpb));
#define LAUNCH_SCORE(C, Gm)                                                                                   \
    score_count_kernel<C, Gm, LIT><<<grid, kScoreThreads, 0, h->stream>>>(                                    \
        cube, PP, pb, gpb, h->rowA.as<int32_t>(), h->rowB.as<int32_t>(), h->thr_c.as<float>(), h->thr_g.as<float>(), T, \
        h->counts_c.as<uint32_t>(), h->counts_g.as<uint32_t>())
    if (c && g) LAUNCH_SCORE(true, true);
    else if (c) LAUNCH_SCORE(true, false);
    else if (g) LAUNCH_SCORE(false, true);
    else return LRPERM_OK;
#undef LAUNCH_SCORE
    CK_LAUNCH();
    return LRPERM_OK;
}


// chunks [c0, c0 + cn) of the full table, or of the filtered one (referenced genes only) when `filt`; cn < 0: all
int launch_fast_any(lrperm_engine* h, int Q, int32_t PB, const uint8_t* slab, float* partials, bool filt, int32_t c0 = 0, int32_t cn = -1) {
    const FastChunk* table = filt ? h->chunks_f.as<FastChunk>() : h->chunks.as<FastChunk>();
    if (cn < 0) cn = (filt ? (int32_t)h->h_chunks_f.size() : h->n_chunks) - c0;
    if (h->fast_variant == 3 && Q == 4) return launch_fast3_q<4, 4>(h, PB, slab, partials, table, c0, cn);
    if (h->fast_variant >= 2) {
        if (Q == 8) return launch_fast2_q<8, 16>(h, PB, slab, partials, table, c0, cn);
        if (Q == 4) return launch_fast2_q<4, 32>(h, PB, slab, partials, table, c0, cn);
        return launch_fast2_q<2, 32>(h, PB, slab, partials, table, c0, cn);
    }
    if (Q == 4) return launch_fast_q<4>(h, PB, slab, partials, table, c0, cn);
    if (Q == 2) return launch_fast_q<2>(h, PB, slab, partials, table, c0, cn);
    return launch_fast_q<1>(h, PB, slab, partials, table, c0, cn);
}

// FAST mode accumulates only the genes some triple reads when nothing else consumes the cube
bool fast_filter_on(const lrperm_engine* h, int scores, const float* cube_out) {
    return h->skip_unref && scores != 0 && cube_out == nullptr && h->have_triples && h->needed_valid &&
           (int64_t)h->h_needed.size() == h->G;
}

struct RunArgs {
    int64_t n_perms, perm_offset;
    uint64_t seed;
    int perm_source;
    const void* perm_idx;
    int perm_idx_is_64, perm_idx_on_device, scores, gmean_mode;
    float* cube_out;
    int cube_on_device;
};

// Batch schedule of the pipelined FAST path, in permutations: batches of PB, the last one rounded up to `gran`
// (the permutations one warp group covers) only.  A schedule with one-group first / last batches (to shorten
// the slab prologue and the finalize + score epilogue that nothing overlaps) was measured and dropped: small
// batches leave the main kernel with too few CTAs per wave (configs[1]: 4.8 ms instead of 3.8 ms).
std::vector<int32_t> batch_schedule(int64_t n_perms, int32_t PB, int32_t gran, int32_t first_batch) {
    std::vector<int32_t> out;
    int64_t left = n_perms;
    if (first_batch > 0 && first_batch < PB && left > first_batch) { out.push_back(first_batch); left -= first_batch; }
    for (; left > 0; left -= PB)
        out.push_back((int32_t)(left >= PB ? PB : (left + gran - 1) / gran * gran));
    return out;
}

// FAST mode, pipelined over two streams.  Stream A (the engine stream) runs only the main kernel;
// stream B runs everything around it: the label slab of batch b+1 while main(b) runs, and
// finalize(b) + score(b) while main(b+1) runs.  Slab and partial-sum buffers are double-buffered.
//   B: slab(0) | slab(1)         | wait m0, finalize(0), score(0), slab(2) | wait m1, finalize(1), ...
//   A: wait s0, main(0), rec m0  | wait s1, main(1), rec m1                | wait s2, wait f0, main(2) ...
int run_fast_pipelined(lrperm_engine* h, const RunArgs& a, int Q, int32_t PB, const uint8_t* slab_labels,
                       int32_t chunk_nnz, int32_t tile_cells) {
    const int64_t N = h->N;
    const int32_t label_scale = h->fast_variant >= 2 ? Q / 2 : 1;
    const int32_t G = (int32_t)h->G, L = h->L;
    const int32_t gran = std::max(128, 32 * Q);
    // option "fast_first_batch": a short first batch would shorten the slab prologue that nothing overlaps.  Measured
    // (scripts/first_batch_ab.py, n_perms=1000): 128 / 256 permutations first -> configs[1] 3.79 -> 4.22 / 4.14 ms,
    // configs[3] shape 14.44 -> 14.60 / 14.44, configs[2] 32.71 -> 33.06 / 32.89: never a gain, default off.
    const int32_t first_batch = std::max(h->fast_first_batch, 0);
    const std::vector<int32_t> sched = batch_schedule(a.n_perms, PB, gran, first_batch / gran * gran);
    const int64_t n_batches = (int64_t)sched.size();
    std::vector<int64_t> first(sched.size() + 1, 0);
    int32_t pb_max = gran;
    for (size_t i = 0; i < sched.size(); ++i) { first[i + 1] = first[i] + sched[i]; pb_max = std::max(pb_max, sched[i]); }
    h->n_sched = (int32_t)n_batches;
    int rc;
    cudaStream_t sA = h->stream, sB = h->aux_stream;
    const size_t slab_max = (size_t)N * pb_max;
    const bool dbl = n_batches > 1;
    CK(h->slab.ensure(slab_max));
    CK(h->cube.ensure(sizeof(float) * (size_t)G * L * pb_max));
    if (dbl) CK(h->slab2.ensure(slab_max));
    uint8_t* slabs[2] = {h->slab.as<uint8_t>(), dbl ? h->slab2.as<uint8_t>() : h->slab.as<uint8_t>()};

    // events: [0] start, [1] end, then per batch: s0 s1 (slab), m0 m1 (main), f0 f1 (finalize), c1 (score end)
    const size_t n_events = 2 + (size_t)n_batches * 7 + 1;
    while (h->events.size() < n_events) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        h->events.push_back(ev);
    }
    auto EV = [&](int64_t b, int k) { return h->events[2 + (size_t)b * 7 + k]; };
    cudaEvent_t ev_fork = h->events[2 + (size_t)n_batches * 7];
    CK(cudaEventRecord(h->events[0], sA));       // run start (after the counter memsets on A)
    CK(cudaEventRecord(ev_fork, sA));
    CK(cudaStreamWaitEvent(sB, ev_fork, 0));     // B starts after everything already queued on A

    const size_t idx_elem = a.perm_idx_is_64 ? 8 : 4;
    // Batch 0 has nothing to hide its label slab behind.  With Philox labels the slab rows of one cell tile can be
    // generated on their own, so the first batch's slab is made tile by tile and the tile's chunk warps start as soon
    // as ITS rows exist (the chunk schedule is tile-major); later batches' slabs are hidden behind the previous batch.
    const int32_t tile_rows = std::max(tile_cells, 1);
    const int32_t n_tiles_run = (int32_t)std::max<int64_t>(1, (N + tile_rows - 1) / tile_rows);
    // (measured, scripts/small_shard_ab.py: with <= 256 permutations - one rank's share on 8 GPUs - the tile-wise launches
    // cost more in partial waves than the hidden slab rows save: 3.94 vs 3.78 ms at 125 permutations; a run that is gated
    // by a pending upload is tile-wise anyway)
    const bool slab0_tiled = a.perm_source == LRPERM_PERMS_PHILOX && n_tiles_run > 1 && h->slab0_by_tile &&
                             (a.n_perms > 256 || h->matrix_pending);
    while (h->ev_slab_tile.size() < (size_t)n_tiles_run) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->ev_slab_tile.push_back(ev);
    }
    while (h->ev_main_tile.size() < 2 * (size_t)n_tiles_run) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        h->ev_main_tile.push_back(ev);
    }
    h->main_tiles_timed = 0;
    auto enqueue_slab = [&](int64_t b) -> int {
        const int64_t p0 = first[(size_t)b];
        const int32_t PBb = sched[(size_t)b];
        const int32_t pb = (int32_t)std::min<int64_t>(PBb, a.n_perms - p0);
        uint8_t* slab = slabs[b & 1];
        PermSource ps{};
        ps.is64 = a.perm_idx_is_64;
        ps.seed = a.seed;
        ps.perm_id0 = a.perm_offset + p0;
        ps.philox = a.perm_source == LRPERM_PERMS_PHILOX;
        CK(cudaEventRecord(EV(b, 0), sB));
        if (pb < PBb) CK(cudaMemsetAsync(slab, 0, (size_t)N * PBb, sB));
        if (ps.philox) {
            const int cells_per_block = 128;
            const size_t tab_smem = sizeof(int32_t) * (size_t)(L + 1) + (size_t)h->n_buckets;
            const int32_t nt = (b == 0 && slab0_tiled) ? n_tiles_run : 1;
            for (int32_t t = 0; t < nt; ++t) {
                const int64_t c0 = nt == 1 ? 0 : (int64_t)t * tile_rows, c1 = nt == 1 ? N : std::min<int64_t>(N, c0 + tile_rows);
                // persistent grid: `slab_ctas_per_sm` CTAs of 4 warps per SM in total (see slab_philox_kernel); 0 = one CTA per block
                const unsigned n_blocks = (unsigned)((c1 - c0 + cells_per_block - 1) / cells_per_block), n_grp = (unsigned)((pb + 127) / 128);
                // (a run of ONE batch has nothing to run next to and its slab is on the critical path: full grid, like the very first
                // tile of a longer run.  Measured, scripts/overlap_ab.py, configs[2], 1000 permutations: 25.9 ms full grid everywhere,
                // 24.0 persistent everywhere, 23.5 persistent everywhere but the first tile; 125 permutations 3.80 vs 4.28 ms)
                unsigned gx = n_blocks;
                const bool first_alone = (b == 0 && t == 0);              // nothing runs next to the very first slab rows
                const bool persist = h->slab_mode == 1 ? true : h->slab_mode == 2 ? !first_alone : h->slab_mode == 3 ? b > 0 : false;
                // CTAs per SM: a quarter of what the main kernel keeps resident (its one-warp CTAs are limited by their bins: 8 per
                // SM at 50 clusters, 19 at 20) - measured best at configs[2] (2: 25.9 -> 23.5 ms), the configs[3] shape (3: 11.9 ->
                // 11.4 ms) and neutral at configs[1] (5: 3.03 ms), where the main kernel's many warps are not starved anyway
                int32_t sc = h->slab_ctas_per_sm;
                if (sc == 0) {
                    const size_t per_cta = (size_t)L * Q * 128 + 512 + 1024;
                    const int32_t main_ctas = (int32_t)std::min<size_t>(32, (228u << 10) / per_cta);
                    sc = main_ctas >= 16 ? -1 : std::max(2, (main_ctas + 3) / 4);      // many small main CTAs: nothing to protect
                }
                // (a single batch gains nothing, tile-wise or not: 13.49 vs 13.37 ms at 500 permutations)
                if (sc > 0 && n_batches > 1 && persist) gx = std::min(n_blocks, std::max(1u, (unsigned)(sc * h->sm_count) / n_grp));
                dim3 grd(gx, n_grp);
                slab_philox_kernel<<<grd, 128, tab_smem, sB>>>(ps, h->cl_start.as<int32_t>(), h->bucket_cluster.as<uint8_t>(), h->n_buckets,
                                                               h->bucket_shift, L, label_scale, (int32_t)N, pb, slab, PBb, cells_per_block,
                                                               (int32_t)c0, (int32_t)c1);
                CK_LAUNCH();
                if (nt > 1) CK(cudaEventRecord(h->ev_slab_tile[(size_t)t], sB));
            }
        } else {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.perm_idx) + idx_elem * (size_t)p0 * (size_t)N;
            if (a.perm_idx_on_device) ps.idx = src;
            else {
                // the staging buffer is reused by every batch: slab(b) is queued on B after slab(b-1), in order
                CK(h->perm_dev.ensure(idx_elem * (size_t)pb_max * (size_t)N));
                CK(cudaMemcpyAsync(h->perm_dev.p, src, idx_elem * (size_t)pb * (size_t)N, cudaMemcpyHostToDevice, sB));
                ps.idx = h->perm_dev.p;
            }
            dim3 grd(grid_for(N, 256, 1024), (unsigned)pb);
            slab_from_index_kernel<<<grd, 256, 0, sB>>>(ps, slab_labels, (int32_t)N, pb, slab, PBb);
            CK_LAUNCH();
        }
        CK(cudaEventRecord(EV(b, 1), sB));
        return LRPERM_OK;
    };

    // The label slabs of the first two batches need the labels only: they are queued BEFORE the engine stream waits
    // for a matrix upload still in flight (lrperm_set_matrix_csr from pinned memory) and before the host blocks in
    // the CSC / chunk-table builds, so they run while the matrix crosses PCIe.
    if ((rc = enqueue_slab(0))) return rc;
    if (n_batches > 1 && (rc = enqueue_slab(1))) return rc;
    // A pinned upload still in flight, tiled the way this run wants it: the first batch runs TILE BY TILE behind the
    // transfer (chunks are scheduled tile-major anyway): the host waits for tile t's CSC sort, builds its chunk
    // table, launches the tile's chunk warps; tiles t+1.. keep crossing PCIe and being sorted on the copy stream.
    const bool gated = h->matrix_pending && h->csc_tile_cells == tile_cells;
    const bool filt = fast_filter_on(h, a.scores, a.cube_out);
    h->filter_tables = filt;
    size_t slots_cap;
    if (gated) {
        if ((rc = submit_tiles(h, -1))) return rc;           // behind the label / triple uploads in the copy engine's queue
        h->chunk_table_nnz = -1;
        h->tiles_tabled = 0;
        slots_cap = (size_t)(h->nnz / std::max(chunk_nnz, 1)) + (size_t)G * (size_t)h->n_tiles + 1;
    } else {
        if ((rc = wait_matrix(h, false))) return rc;
        if ((rc = ensure_fast_tables(h, chunk_nnz, tile_cells))) return rc;
        if (filt && (rc = ensure_filtered_chunks(h))) return rc;
        slots_cap = (size_t)std::max(h->n_slots, 1);
    }
    const uint8_t* d_needed = filt ? h->needed_dev.as<uint8_t>() : nullptr;
    if ((rc = dbg_set_dims(h, (int64_t)slots_cap, pb_max))) return rc;
    const size_t part_max = sizeof(float) * slots_cap * L * pb_max;
    CK(h->partials.ensure(part_max));
    if (dbl) CK(h->partials2.ensure(part_max));
    float* parts[2] = {h->partials.as<float>(), dbl ? h->partials2.as<float>() : h->partials.as<float>()};
    for (int64_t b = 0; b < n_batches; ++b) {
        const int64_t p0 = first[(size_t)b];
        const int32_t PBb = sched[(size_t)b];
        const int32_t pb = (int32_t)std::min<int64_t>(PBb, a.n_perms - p0);
        if (b >= 1 && b + 1 < n_batches && (rc = enqueue_slab(b + 1))) return rc;
        // ---- A: main kernel ----
        const bool by_tile = b == 0 && slab0_tiled && h->n_tiles == n_tiles_run;
        if (!by_tile) CK(cudaStreamWaitEvent(sA, EV(b, 1), 0));
        if (b >= 2) CK(cudaStreamWaitEvent(sA, EV(b - 2, 5), 0));     // partials[b&1] consumed by finalize(b-2)
        CK(cudaEventRecord(EV(b, 2), sA));
        if (b == 0 && (gated || by_tile)) {
            for (int32_t t = 0; t < h->n_tiles; ++t) {
                if (gated && (rc = chunks_append_tile(h, t, chunk_nnz, sA))) return rc;      // host waits for tile t's CSC
                if (by_tile) CK(cudaStreamWaitEvent(sA, h->ev_slab_tile[(size_t)t], 0));   // the tile's slab rows exist
                const std::vector<int32_t>& tcp = filt ? h->h_tile_chunk_ptr_f : h->h_tile_chunk_ptr;
                const int32_t c0 = tcp[(size_t)t], c1 = tcp[(size_t)t + 1];
                // kernel time of the tile alone (the waits above - slab rows, CSC of an arriving tile - are not kernel time)
                CK(cudaEventRecord(h->ev_main_tile[2 * (size_t)t], sA));
                if (c1 > c0 && (rc = launch_fast_any(h, Q, PBb, slabs[0], parts[0], filt, c0, c1 - c0))) return rc;
                CK(cudaEventRecord(h->ev_main_tile[2 * (size_t)t + 1], sA));
            }
            h->main_tiles_timed = h->n_tiles;
            if (gated) {
                h->chunk_table_nnz = chunk_nnz;
                // the whole matrix is resident now; its column-range validation is read back with the results
                h->matrix_pending = false;
                CK(cudaStreamWaitEvent(sA, h->ev_matrix, 0));
                h->check_matrix_flag = true;
            }
        } else if ((rc = launch_fast_any(h, Q, PBb, slabs[b & 1], parts[b & 1], filt))) return rc;
        CK(cudaEventRecord(EV(b, 3), sA));
        // ---- B: finalize + score (+ optional cube export) ----
        CK(cudaStreamWaitEvent(sB, EV(b, 3), 0));
        CK(cudaEventRecord(EV(b, 4), sB));
        fast_finalize_kernel<<<grid_for((int64_t)G * L * PBb, 256, h->sm_count * 8), 256, 0, sB>>>(
            parts[b & 1], h->tile_slot_ptr.as<int32_t>(), h->n_tiles, h->inv_n.as<float>(), G, L, PBb, h->cube.as<float>(), PBb, 0, d_needed);
        CK_LAUNCH();
        CK(cudaEventRecord(EV(b, 5), sB));
        if (a.scores) {
            cudaStream_t saved = h->stream;
            h->stream = sB;
            if (a.gmean_mode == LRPERM_GMEAN_LITERAL) rc = launch_score_t<true>(h, a.scores, h->cube.as<float>(), PBb, pb);
            else rc = launch_score_t<false>(h, a.scores, h->cube.as<float>(), PBb, pb);
            h->stream = saved;
            if (rc) return rc;
        }
        if (a.cube_out) {
            float* dst = a.cube_out + (size_t)p0 * L * G;
            float* dev_dst = dst;
            if (!a.cube_on_device) { CK(h->stage_out.ensure(sizeof(float) * (size_t)pb_max * L * G)); dev_dst = h->stage_out.as<float>(); }
            dim3 blk(32, 8), grd((G + 31) / 32, (pb + 31) / 32, L);
            cube_export_kernel<<<grd, blk, 0, sB>>>(h->cube.as<float>(), G, L, PBb, pb, dev_dst);
            CK_LAUNCH();
            if (!a.cube_on_device) {
                CK(cudaMemcpyAsync(dst, dev_dst, sizeof(float) * (size_t)pb * L * G, cudaMemcpyDeviceToHost, sB));
                CK(cudaStreamSynchronize(sB));   // staging buffer is reused by the next batch
            }
        }
        CK(cudaEventRecord(EV(b, 6), sB));
    }
    // join: A continues (result copies) after the last score on B
    CK(cudaStreamWaitEvent(sA, EV(n_batches - 1, 6), 0));
    CK(cudaEventRecord(h->events[1], sA));
    h->filter_tables = false;
    fast_run_stats(h, filt);
    return LRPERM_OK;
}

int collect_pipelined_timings(lrperm_engine* h, int64_t n_batches) {
    auto EV = [&](int64_t b, int k) { return h->events[2 + (size_t)b * 7 + k]; };
    h->t_setup = h->t_means = h->t_score = h->t_main = 0.f;
    h->n_main_launches = (int32_t)n_batches;
    for (int64_t b = 0; b < n_batches; ++b) {
        float ms;
        CK(cudaEventElapsedTime(&ms, EV(b, 0), EV(b, 1))); h->t_setup += ms;
        CK(cudaEventElapsedTime(&ms, EV(b, 2), EV(b, 3))); h->t_means += ms;
        if (b == 0 && h->main_tiles_timed > 0) {        // launched tile by tile: sum of the launches, without the gating waits
            for (int32_t t = 0; t < h->main_tiles_timed; ++t) {
                CK(cudaEventElapsedTime(&ms, h->ev_main_tile[2 * (size_t)t], h->ev_main_tile[2 * (size_t)t + 1]));
                h->t_main += ms;
            }
        } else h->t_main += ms;
        CK(cudaEventElapsedTime(&ms, EV(b, 4), EV(b, 5))); h->t_means += ms;
        CK(cudaEventElapsedTime(&ms, EV(b, 5), EV(b, 6))); h->t_score += ms;
    }
    CK(cudaEventElapsedTime(&h->t_total, h->events[0], h->events[1]));
    return LRPERM_OK;
}


// ---- CellChat trimean path: value-sorted CSC, chunk / rank tables -------------------------------------
int build_vsc(lrperm_engine* h) {
    if (h->vsc_ready) return LRPERM_OK;
    StageTimer timer("build_vsc");
    const int64_t nnz = h->nnz;
    const int32_t G = (int32_t)h->G;
    const size_t n1 = (size_t)std::max<int64_t>(nnz, 1);
    CK(h->tmp_a.ensure(sizeof(unsigned long long) * n1));    // keys {column, value key}
    CK(h->tmp_c.ensure(sizeof(unsigned long long) * n1));    // sorted keys (discarded)
    CK(h->tmp_b.ensure(sizeof(unsigned long long) * n1));    // payload {value bits, cell}
    DevBuf& col_count = h->small_a;
    CK(col_count.ensure(sizeof(int32_t) * (size_t)(G + 1)));
    CK(cudaMemsetAsync(col_count.p, 0, sizeof(int32_t) * (size_t)(G + 1), h->stream));
    CK(h->vsc_pairs.ensure(sizeof(int2) * n1));
    h->h_csc_ptr.assign((size_t)G + 1, 0);
    if (nnz > 0) {
        vsc_keys_kernel<<<h->sm_count * 4, 512, 0, h->stream>>>(h->indptr.as<int32_t>(), h->csr_pairs.as<int2>(), (int32_t)h->N, G,
                                                               h->tmp_a.as<unsigned long long>(), h->tmp_b.as<unsigned long long>(), col_count.as<int32_t>());
        CK_LAUNCH();
        int end_bit = 1;
        while ((1 << end_bit) < G && end_bit < 31) ++end_bit;
        end_bit += 32;
        size_t tmp_bytes = 0;
        unsigned long long* sorted = reinterpret_cast<unsigned long long*>(h->vsc_pairs.p);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->tmp_a.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(),
                                           h->tmp_b.as<unsigned long long>(), sorted, (int)nnz, 0, end_bit, h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        CK(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, h->tmp_a.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(),
                                           h->tmp_b.as<unsigned long long>(), sorted, (int)nnz, 0, end_bit, h->stream));
        h->launches += 8;
        std::vector<int32_t> cnt((size_t)G + 1, 0);
        CK(cudaMemcpyAsync(cnt.data(), col_count.p, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int32_t g = 0; g < G; ++g) h->h_csc_ptr[(size_t)g + 1] = h->h_csc_ptr[(size_t)g] + cnt[(size_t)g];
    } else {
        CK(cudaStreamSynchronize(h->stream));
    }
    h->vsc_ready = true;
    h->tm_table_nnz = -1;
    return LRPERM_OK;
}

// chunk table of the trimean path: every gene's positive region and negative region (the explicit zeros between
// them need no walk), cut every chunk_nnz entries; slots are consecutive inside a region; longest chunks first
int build_tm_chunks(lrperm_engine* h, int32_t chunk_nnz, bool filt) {
    // filt: only the genes some triple reads get chunks (their regions are independent of each other, so nothing else
    // changes); the table is then tied to the triple set it was built for
    const int64_t want_needed = filt ? h->needed_version : 0;
    if (h->tm_table_nnz == chunk_nnz && h->tm_table_needed == want_needed) return LRPERM_OK;
    StageTimer timer("build_tm_chunks");
    const int32_t G = (int32_t)h->G;
    std::vector<int32_t> pos_end((size_t)G), neg_begin((size_t)G);
    {
        DevBuf &d_ptr = h->small_a, &d_b = h->small_b;
        CK(d_ptr.ensure(sizeof(int32_t) * ((size_t)G + 1)));
        CK(d_b.ensure(sizeof(int32_t) * 2 * (size_t)G));
        CK(cudaMemcpyAsync(d_ptr.p, h->h_csc_ptr.data(), sizeof(int32_t) * ((size_t)G + 1), cudaMemcpyHostToDevice, h->stream));
        tm_bounds_kernel<<<grid_for(G, 128, 1 << 30), 128, 0, h->stream>>>(h->vsc_pairs.as<int2>(), d_ptr.as<int32_t>(), G,
                                                                         d_b.as<int32_t>(), d_b.as<int32_t>() + G);
        CK_LAUNCH();
        CK(cudaMemcpyAsync(pos_end.data(), d_b.p, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(neg_begin.data(), d_b.as<int32_t>() + G, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    std::vector<FastChunk> chunks;
    std::vector<int32_t> region_ptr(1, 0);
    std::vector<uint8_t> region_dir;
    int32_t slot = 0;
    for (int32_t g = 0; g < G; ++g) {
        if (filt && !h->h_needed[(size_t)g]) continue;
        const int32_t b[2] = {h->h_csc_ptr[(size_t)g], neg_begin[(size_t)g]};
        const int32_t e[2] = {pos_end[(size_t)g], h->h_csc_ptr[(size_t)g + 1]};
        for (int dir = 0; dir < 2; ++dir) {
            if (e[dir] <= b[dir]) continue;
            region_dir.push_back((uint8_t)dir);
            for (int32_t s0 = b[dir]; s0 < e[dir]; s0 += chunk_nnz) {
                FastChunk c;
                c.gene = g | (dir << kTmDirBit); c.begin = s0; c.end = std::min(e[dir], s0 + chunk_nnz); c.slot = slot++;
                chunks.push_back(c);
            }
            region_ptr.push_back(slot);
        }
    }
    std::stable_sort(chunks.begin(), chunks.end(), [](const FastChunk& a, const FastChunk& b) { return (a.end - a.begin) > (b.end - b.begin); });
    h->tm_n_chunks = (int32_t)chunks.size();
    h->tm_n_slots = slot;
    h->tm_n_regions = (int32_t)region_ptr.size() - 1;
    std::vector<int32_t> chunk_of_slot((size_t)std::max(slot, 1), 0);
    for (size_t i = 0; i < chunks.size(); ++i) chunk_of_slot[(size_t)chunks[i].slot] = (int32_t)i;
    int rc;
    if ((rc = upload(h, h->tm_chunk_of_slot, chunk_of_slot.data(), sizeof(int32_t) * chunk_of_slot.size(), 0))) return rc;
    if (region_dir.empty()) region_dir.push_back(0);
    if ((rc = upload(h, h->tm_region_dir, region_dir.data(), region_dir.size(), 0))) return rc;
    CK(h->tm_chunks.ensure(sizeof(FastChunk) * std::max<size_t>(chunks.size(), 1)));
    CK(h->tm_region_ptr.ensure(sizeof(int32_t) * region_ptr.size()));
    if (!chunks.empty())
        CK(cudaMemcpyAsync(h->tm_chunks.p, chunks.data(), sizeof(FastChunk) * chunks.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->tm_region_ptr.p, region_ptr.data(), sizeof(int32_t) * region_ptr.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm_table_nnz = chunk_nnz;
    h->tm_table_needed = want_needed;
    return LRPERM_OK;
}

// Which order statistics a trimean of n values needs (numpy: np.quantile method 'linear' = Hyndman & Fan 7,
// virtual index (n - 1) q, neighbours floor / floor + 1; np.median = mean of the middle one / two), as ranks to
// count down to while walking the positives from the top (dir 0) or the negatives from the bottom (dir 1).
int build_tm_tables(lrperm_engine* h) {
    if (h->tm_tables_ready) return LRPERM_OK;
    const int32_t L = h->L;
    std::vector<int32_t> tgt((size_t)2 * L * 8, 0);
    std::vector<TmRoles> roles((size_t)L);
    for (int32_t c = 0; c < L; ++c) {
        const int64_t n = h->h_cl_start[(size_t)c + 1] - h->h_cl_start[(size_t)c];
        int64_t want[6];
        const double v25 = (double)(n - 1) * 0.25, v75 = (double)(n - 1) * 0.75;
        auto neighbours = [&](double virt, int64_t& lo, int64_t& hi) {
            lo = (int64_t)std::floor(virt); hi = lo + 1;
            if (virt >= (double)(n - 1)) { lo = n - 1; hi = n - 1; }
        };
        neighbours(v25, want[0], want[1]);
        want[2] = (n & 1) ? n / 2 : n / 2 - 1;
        want[3] = n / 2;
        neighbours(v75, want[4], want[5]);
        std::vector<int64_t> uniq(want, want + 6);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        const int m = (int)uniq.size();
        for (int j = 0; j < 6; ++j) roles[(size_t)c].plane[j] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), want[j]) - uniq.begin());
        roles[(size_t)c].g25 = (float)(v25 - std::floor(v25));
        roles[(size_t)c].g75 = (float)(v75 - std::floor(v75));
        for (int k = 0; k < m; ++k) {
            tgt[((size_t)0 * L + c) * 8 + k] = (int32_t)(n - uniq[(size_t)(m - 1 - k)]);     // rank from the top, ascending
            tgt[((size_t)1 * L + c) * 8 + k] = (int32_t)(uniq[(size_t)k] + 1);               // rank from the bottom, ascending
        }
        tgt[((size_t)0 * L + c) * 8 + 6] = m;
        tgt[((size_t)1 * L + c) * 8 + 6] = m;
    }
    int rc;
    if ((rc = upload(h, h->tm_tgt, tgt.data(), sizeof(int32_t) * tgt.size(), 0))) return rc;
    if ((rc = upload(h, h->tm_roles, roles.data(), sizeof(TmRoles) * roles.size(), 0))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->tm_tables_ready = true;
    return LRPERM_OK;
}

template <int Q>
int launch_tm_count(lrperm_engine* h, int32_t PB, const uint8_t* slab, float* counts, uint32_t* prefix, int32_t n_valid, int32_t batch_in_super,
                    cudaEvent_t* ev /*[3]*/) {
    constexpr int B = 32;
    const int32_t L = h->L;
    const int32_t groups = PB / (32 * Q);
    const int64_t grid = (int64_t)h->tm_n_chunks * groups;
    if (grid > 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "trimean: too many chunks (%lld CTAs)", (long long)grid);
    const size_t smem_count = (size_t)L * Q * 128 + 16 * B;
    CK(cudaFuncSetAttribute(fast2_means_kernel<Q, B, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
    CK(cudaEventRecord(ev[0], h->stream));
    if (grid > 0) {
        fast2_means_kernel<Q, B, true><<<(unsigned)grid, 32, smem_count, h->stream>>>(
            h->vsc_pairs.as<int2>(), h->tm_chunks.as<FastChunk>(), h->tm_n_chunks, slab, PB, L, groups, counts);
        CK_LAUNCH();
    }
    CK(cudaEventRecord(ev[1], h->stream));
    if (h->tm_n_regions > 0) {
        uint32_t* wl = h->tm_walk_list.as<uint32_t>();           // [0] item count, [1] cursor, [2..] items
        tm_prefix_kernel<<<grid_for((int64_t)h->tm_n_regions * L * PB, 256, h->sm_count * 16), 256, 0, h->stream>>>(
            counts, h->tm_region_ptr.as<int32_t>(), h->tm_region_dir.as<uint8_t>(), h->tm_tgt.as<int32_t>(), h->tm_n_regions, L, PB,
            32 * Q, n_valid, (uint32_t)batch_in_super * (uint32_t)h->tm_n_slots * (uint32_t)groups, prefix,
            h->tm_walk_flag.as<uint32_t>(), wl + 2, wl);
        CK_LAUNCH();
    }
    CK(cudaEventRecord(ev[2], h->stream));
    return LRPERM_OK;
}

template <int Q>
int launch_tm_select(lrperm_engine* h, int32_t PB, int32_t nb, int32_t n_perms_super, size_t slot_elems, size_t plane_elems) {
    constexpr int B = 32;
    const int32_t L = h->L;
    const int32_t groups = PB / (32 * Q);
    const size_t smem_sel = (size_t)L * Q * 128 + 16 * B + sizeof(int32_t) * 8 * (size_t)L;
    CK(cudaFuncSetAttribute(tm_select_kernel<Q, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
    if (h->tm_n_chunks > 0) {
        // persistent: as many one-warp CTAs as stay resident (shared memory bound), they pull items off the list
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tm_select_kernel<Q, B>, 32, smem_sel));
        const unsigned grid = (unsigned)std::max(1, per_sm) * (unsigned)h->sm_count;
        uint32_t* wl = h->tm_walk_list.as<uint32_t>();
        tm_select_kernel<Q, B><<<grid, 32, smem_sel, h->stream>>>(
            h->vsc_pairs.as<int2>(), h->tm_chunks.as<FastChunk>(), h->tm_chunk_of_slot.as<int32_t>(), h->slab.as<uint8_t>(), PB, L, groups,
            h->tm_n_slots, h->tm_counts.as<float>(), h->tm_prefix.as<uint32_t>(), h->tm_tgt.as<int32_t>(), h->tm_planes.as<float>(),
            (int64_t)nb * (int64_t)plane_elems, (int64_t)h->N * PB, (int64_t)slot_elems, (int64_t)plane_elems, n_perms_super,
            wl + 2, wl, wl + 1, h->tm_diag.as<unsigned long long>());
        CK_LAUNCH();
        h->tm_select_warps += (int64_t)h->tm_n_chunks * groups * nb;
    }
    return LRPERM_OK;
}

struct TrimeanArgs {
    int64_t n_perms, perm_offset;
    uint64_t seed;
    int perm_source;            // LRPERM_PERMS_INDEX / LRPERM_PERMS_PHILOX, or -1 = identity (observed statistic)
    const void* perm_idx;
    int perm_idx_is_64, perm_idx_on_device;
    float r32;
    double r64, kh;
    bool score;
    double* cube_out;           // [n_perms][L][G] or NULL
    int cube_on_device;
};

// Super-batches of nb batches of PB permutations.  Per batch: slab -> count -> prefix (the slab of ONE batch is what
// should sit in L2 during the count pass).  Per super-batch: ONE select launch over all its batches, one finalize,
// then per batch score (-> export).
int run_trimean_batches(lrperm_engine* h, const TrimeanArgs& a) {
    const int64_t N = h->N;
    const int32_t G = (int32_t)h->G, L = h->L;
    const int64_t T = h->T;
    int rc;
    if (L > 255) return h->fail(LRPERM_ERR_UNSUPPORTED, "trimean: at most 255 clusters (got %d)", L);
    if (N >= (1LL << 28)) return h->fail(LRPERM_ERR_UNSUPPORTED, "trimean: n_cells must be < 2^28");
    int Q = 0;
    for (int q : {4, 2}) {
        if ((L - 1) * q / 2 > 255) continue;
        if ((size_t)L * q * 128 + 512 + 32 * (size_t)L <= h->smem_optin) { Q = q; break; }
    }
    if (Q == 0) return h->fail(LRPERM_ERR_UNSUPPORTED, "trimean: %d clusters do not fit the shared-memory bins", L);
    if ((rc = build_vsc(h))) return rc;
    // entries per chunk: the select pass re-walks whole chunks, so they are shorter than FAST mode's
    const int32_t chunk_nnz = h->tm_chunk_nnz ? h->tm_chunk_nnz : 8192;
    const bool filt = a.score && a.cube_out == nullptr && h->skip_unref && h->have_triples && h->needed_valid && (int64_t)h->h_needed.size() == h->G;
    if ((rc = build_tm_chunks(h, chunk_nnz, filt))) return rc;
    if ((rc = build_tm_tables(h))) return rc;
    if ((rc = dbg_set_dims(h, std::max(h->tm_n_slots, 1), 0))) return rc;
    // permutations per batch: the walk visits cells in VALUE order, so the whole label slab of a batch (N x PB
    // bytes) should sit in L2: <= 64 MiB
    int32_t PB = h->tm_perm_batch;
    if (!PB) PB = (int32_t)std::min<int64_t>(512, std::max<int64_t>(128, ((64LL << 20) / std::max<int64_t>(N, 1)) / 128 * 128));
    if (a.n_perms < PB) PB = (int32_t)std::max<int64_t>(128, ((a.n_perms + 127) / 128) * 128);
    h->last_pb = PB;
    const size_t slot_elems = (size_t)std::max(h->tm_n_slots, 1) * L * PB;
    const size_t plane_elems = (size_t)G * L * PB;
    const size_t slab_elems = (size_t)std::max<int64_t>(N, 1) * PB;
    // batches per super-batch: everything of a super-batch stays resident (<= ~12 GiB)
    const size_t per_batch = slab_elems + 8 * slot_elems + (4 * kTmPlanes + 8) * plane_elems;
    const int64_t n_batches = (a.n_perms + PB - 1) / PB;
    const int32_t NB = (int32_t)std::max<int64_t>(1, std::min<int64_t>(n_batches, (int64_t)((12ULL << 30) / std::max<size_t>(per_batch, 1))));
    CK(h->slab.ensure(slab_elems * NB));
    CK(h->tm_counts.ensure(sizeof(float) * slot_elems * NB));
    CK(h->tm_prefix.ensure(sizeof(uint32_t) * slot_elems * NB));
    CK(h->tm_planes.ensure(sizeof(float) * kTmPlanes * plane_elems * NB));
    CK(h->tm_cube.ensure(sizeof(double) * plane_elems * NB));
    const size_t n_items_max = (size_t)std::max(h->tm_n_slots, 1) * (size_t)(PB / (32 * Q)) * (size_t)NB;
    if (n_items_max >= 0xffffffffULL) return h->fail(LRPERM_ERR_UNSUPPORTED, "trimean: too many (chunk, group) items per super-batch");
    CK(h->tm_walk_flag.ensure(sizeof(uint32_t) * n_items_max));
    CK(h->tm_walk_list.ensure(sizeof(uint32_t) * (n_items_max + 2)));
    // label bytes pre-scaled by Q/2 (see fast2_means_kernel)
    const int32_t scale = Q / 2;
    if (h->labels8s_scale != scale) {
        CK(h->labels8s.ensure((size_t)std::max<int64_t>(N, 1)));
        if (N > 0) {
            scale_labels_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, h->stream>>>(h->labels8.as<uint8_t>(), (int32_t)N, scale, h->labels8s.as<uint8_t>());
            CK_LAUNCH();
        }
        h->labels8s_scale = scale;
    }
    const int64_t n_super = (n_batches + NB - 1) / NB;
    // events: [0] start; per batch 4 (slab start, count start, count end, prefix end) + 2 (score start, score end)
    // + 1 (slab end, on the aux stream); per super-batch 3 (select start, select end, finalize end); last = end
    const size_t n_events = 2 + (size_t)n_batches * 7 + (size_t)n_super * 3;
    // the label slabs are generated on the aux stream, ahead of and concurrently with the counting of earlier batches
    // (ALU-bound Feistel rounds next to a shared-memory-bound kernel); every batch of a super-batch has its own slab
    cudaStream_t sM = h->stream, sS = h->overlap ? h->aux_stream : h->stream;
    while (h->events.size() < n_events) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        h->events.push_back(ev);
    }
    CK(h->tm_diag.ensure(8));
    CK(cudaMemsetAsync(h->tm_diag.p, 0, 8, h->stream));
    h->tm_select_warps = 0;
    h->tm_n_batches = (int32_t)n_batches;
    h->tm_n_super = (int32_t)n_super;
    CK(cudaEventRecord(h->events[0], h->stream));
    if (sS != sM) CK(cudaStreamWaitEvent(sS, h->events[0], 0));       // the aux stream starts after everything queued so far
    const size_t idx_elem = a.perm_idx_is_64 ? 8 : 4;
    for (int64_t sb = 0; sb < n_super; ++sb) {
        const int64_t b0 = sb * NB;
        // the slabs of this super-batch overwrite those the previous select pass reads
        if (sb > 0 && sS != sM) CK(cudaStreamWaitEvent(sS, h->events[1 + (size_t)n_batches * 7 + (size_t)(sb - 1) * 3 + 1], 0));
        const int32_t nb = (int32_t)std::min<int64_t>(NB, n_batches - b0);
        const int32_t n_perms_super = (int32_t)std::min<int64_t>((int64_t)nb * PB, a.n_perms - b0 * PB);
        CK(cudaMemsetAsync(h->tm_planes.p, 0, sizeof(float) * kTmPlanes * plane_elems * nb, h->stream));
        CK(cudaMemsetAsync(h->tm_walk_flag.p, 0, sizeof(uint32_t) * n_items_max, h->stream));
        CK(cudaMemsetAsync(h->tm_walk_list.p, 0, 2 * sizeof(uint32_t), h->stream));
        for (int32_t bi = 0; bi < nb; ++bi) {
            const int64_t b = b0 + bi;
            cudaEvent_t* ev = &h->events[1 + (size_t)b * 7];
            const int64_t p0 = b * PB;
            const int32_t pb = (int32_t)std::min<int64_t>(PB, a.n_perms - p0);
            uint8_t* slab = h->slab.as<uint8_t>() + slab_elems * bi;
            CK(cudaEventRecord(ev[0], sS));
            // ---- label slab (aux stream) ----
            if (pb < PB) CK(cudaMemsetAsync(slab, 0, slab_elems, sS));
            if (a.perm_source < 0) {
                slab_identity_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, sS>>>(h->labels8s.as<uint8_t>(), (int32_t)N, slab, PB);
            } else {
                PermSource ps{};
                ps.is64 = a.perm_idx_is_64;
                ps.seed = a.seed;
                ps.perm_id0 = a.perm_offset + p0;
                ps.philox = a.perm_source == LRPERM_PERMS_PHILOX;
                if (ps.philox) {
                    const int cells_per_block = 128;
                    // later batches' slabs run next to the count kernel of an earlier batch: persistent small grid (see slab_philox_kernel)
                    const unsigned n_blocks = (unsigned)((N + cells_per_block - 1) / cells_per_block), n_grp = (unsigned)((pb + 127) / 128);
                    unsigned gx = n_blocks;
                    {
                        int32_t sc = h->slab_ctas_per_sm;
                        if (sc == 0) {
                            const int32_t main_ctas = (int32_t)std::min<size_t>(32, (228u << 10) / ((size_t)L * Q * 128 + 512 + 1024));
                            sc = main_ctas >= 16 ? -1 : std::max(2, (main_ctas + 3) / 4);
                        }
                        if (sc > 0 && h->slab_mode > 0 && b > 0 && sS != sM) gx = std::min(n_blocks, std::max(1u, (unsigned)(sc * h->sm_count) / n_grp));
                    }
                    dim3 grd(gx, n_grp);
                    const size_t tab_smem = sizeof(int32_t) * (size_t)(L + 1) + (size_t)h->n_buckets;
                    slab_philox_kernel<<<grd, 128, tab_smem, sS>>>(ps, h->cl_start.as<int32_t>(), h->bucket_cluster.as<uint8_t>(), h->n_buckets,
                                                                   h->bucket_shift, L, scale, (int32_t)N, pb, slab, PB, cells_per_block);
                } else {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.perm_idx) + idx_elem * (size_t)p0 * (size_t)N;
                    if (a.perm_idx_on_device) ps.idx = src;
                    else {
                        // one staging buffer, reused batch after batch: uploads and slab kernels are ordered on the aux stream
                        CK(h->perm_dev.ensure(idx_elem * (size_t)PB * (size_t)N));
                        CK(cudaMemcpyAsync(h->perm_dev.p, src, idx_elem * (size_t)pb * (size_t)N, cudaMemcpyHostToDevice, sS));
                        ps.idx = h->perm_dev.p;
                    }
                    dim3 grd(grid_for(N, 256, 1024), (unsigned)pb);
                    slab_from_index_kernel<<<grd, 256, 0, sS>>>(ps, h->labels8s.as<uint8_t>(), (int32_t)N, pb, slab, PB);
                }
            }
            CK_LAUNCH();
            CK(cudaEventRecord(ev[6], sS));
            if (sS != sM) CK(cudaStreamWaitEvent(sM, ev[6], 0));
            // ---- count, prefix ----
            float* counts = h->tm_counts.as<float>() + slot_elems * bi;
            uint32_t* prefix = h->tm_prefix.as<uint32_t>() + slot_elems * bi;
            if (Q == 4) rc = launch_tm_count<4>(h, PB, slab, counts, prefix, pb, bi, ev + 1);
            else rc = launch_tm_count<2>(h, PB, slab, counts, prefix, pb, bi, ev + 1);
            if (rc) return rc;
        }
        cudaEvent_t* sev = &h->events[1 + (size_t)n_batches * 7 + (size_t)sb * 3];
        // ---- select: one launch over the batches of the super-batch ----
        CK(cudaEventRecord(sev[0], h->stream));
        if (Q == 4) rc = launch_tm_select<4>(h, PB, nb, n_perms_super, slot_elems, plane_elems);
        else rc = launch_tm_select<2>(h, PB, nb, n_perms_super, slot_elems, plane_elems);
        if (rc) return rc;
        CK(cudaEventRecord(sev[1], h->stream));
        // ---- numpy's interpolation arithmetic -> float64 cube [batch][G][L][PB] ----
        {
            const int64_t total = (int64_t)plane_elems * nb;
            const int grid = grid_for(total, 256, h->sm_count * 16);
            if (a.perm_source < 0)
                tm_finalize_kernel<true><<<grid, 256, 0, h->stream>>>(h->tm_planes.as<float>(), total, h->tm_roles.as<TmRoles>(),
                                                                      G * nb, L, PB, a.r32, a.r64, h->tm_cube.as<double>(), PB);
            else
                tm_finalize_kernel<false><<<grid, 256, 0, h->stream>>>(h->tm_planes.as<float>(), total, h->tm_roles.as<TmRoles>(),
                                                                       G * nb, L, PB, a.r32, a.r64, h->tm_cube.as<double>(), PB);
            CK_LAUNCH();
        }
        CK(cudaEventRecord(sev[2], h->stream));
        for (int32_t bi = 0; bi < nb; ++bi) {
            const int64_t b = b0 + bi;
            cudaEvent_t* ev = &h->events[1 + (size_t)b * 7];
            const int64_t p0 = b * PB;
            const int32_t pb = (int32_t)std::min<int64_t>(PB, a.n_perms - p0);
            const double* cube = h->tm_cube.as<double>() + plane_elems * bi;
            CK(cudaEventRecord(ev[4], h->stream));
            if (a.score && T > 0) {
                const int32_t n_groups = (pb + 31) / 32;
                const int32_t gpb = score_groups_per_block(h, n_groups, sizeof(double));
                dim3 sgrid((unsigned)((T + kScoreTile - 1) / kScoreTile), (unsigned)((n_groups + gpb - 1) / gpb));
                score_count_cellchat_kernel<<<sgrid, kScoreThreads, 0, h->stream>>>(
                    cube, PB, pb, gpb, h->rowA.as<int32_t>(), h->rowB.as<int32_t>(), h->thr_d.as<double>(), a.kh, T, h->counts_t.as<uint32_t>());
                CK_LAUNCH();
            }
            CK(cudaEventRecord(ev[5], h->stream));
            if (a.cube_out) {
                double* dst = a.cube_out + (size_t)p0 * L * G;
                double* dev_dst = dst;
                if (!a.cube_on_device) { CK(h->stage_out.ensure(sizeof(double) * (size_t)pb * L * G)); dev_dst = h->stage_out.as<double>(); }
                dim3 blk(32, 8), grd((G + 31) / 32, (pb + 31) / 32, L);
                cube_export_f64_kernel<<<grd, blk, 0, h->stream>>>(cube, G, L, PB, pb, dev_dst);
                CK_LAUNCH();
                if (!a.cube_on_device) {
                    CK(cudaMemcpyAsync(dst, dev_dst, sizeof(double) * (size_t)pb * L * G, cudaMemcpyDeviceToHost, h->stream));
                    CK(cudaStreamSynchronize(h->stream));
                }
            }
        }
    }
    CK(cudaEventRecord(h->events[n_events - 1], h->stream));
    unsigned long long walked = 0;
    CK(cudaMemcpyAsync(&walked, h->tm_diag.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm_walked_warps = (int64_t)walked;
    return LRPERM_OK;
}

int collect_trimean_timings(lrperm_engine* h) {
    const int64_t n_batches = h->tm_n_batches, n_super = h->tm_n_super;
    h->t_setup = h->t_means = h->t_score = h->t_main = 0.f;
    for (float& v : h->tm_ms) v = 0.f;
    h->n_main_launches = (int32_t)n_batches;
    float ms;
    for (int64_t b = 0; b < n_batches; ++b) {
        cudaEvent_t* ev = &h->events[1 + (size_t)b * 7];
        CK(cudaEventElapsedTime(&ms, ev[0], ev[6])); h->t_setup += ms;
        CK(cudaEventElapsedTime(&ms, ev[1], ev[2])); h->tm_ms[0] += ms; h->t_main += ms;
        CK(cudaEventElapsedTime(&ms, ev[2], ev[3])); h->tm_ms[1] += ms;
        CK(cudaEventElapsedTime(&ms, ev[4], ev[5])); h->t_score += ms;
    }
    for (int64_t sb = 0; sb < n_super; ++sb) {
        cudaEvent_t* sev = &h->events[1 + (size_t)n_batches * 7 + (size_t)sb * 3];
        CK(cudaEventElapsedTime(&ms, sev[0], sev[1])); h->tm_ms[2] += ms;
        CK(cudaEventElapsedTime(&ms, sev[1], sev[2])); h->tm_ms[3] += ms;
    }
    h->t_means = h->tm_ms[0] + h->tm_ms[1] + h->tm_ms[2] + h->tm_ms[3];
    CK(cudaEventElapsedTime(&h->t_total, h->events[0], h->events[1 + (size_t)n_batches * 7 + (size_t)n_super * 3]));
    return LRPERM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int lrperm_abi_version(void) { return LRPERM_ABI_VERSION; }

const char* lrperm_last_error(const lrperm_engine* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lrperm_create(int device, lrperm_engine** out) {
    if (!out) { g_create_error = "lrperm_create: out is NULL"; return LRPERM_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("lrperm_create: no CUDA device available (") + cudaGetErrorString(e) + ")";
        return LRPERM_ERR_CUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "lrperm_create: device index out of range"; return LRPERM_ERR_INVALID; }
    lrperm_engine* h = new (std::nothrow) lrperm_engine();
    if (!h) { g_create_error = "lrperm_create: host allocation failed"; return LRPERM_ERR_OOM; }
    h->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // highest priority: the CTA dispatcher otherwise drains the main kernel's whole grid before it
        // starts a kernel of another stream, and the overlap would only happen at the tail
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_matrix, cudaEventDisableTiming);
    if (e == cudaSuccess) e = h->err_flag.ensure(sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->err_flag.p, 0, sizeof(int));
    if (e == cudaSuccess) e = h->err_flag_m.ensure(sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->err_flag_m.p, 0, sizeof(int));
    if (e != cudaSuccess) {
        g_create_error = std::string("lrperm_create: ") + cudaGetErrorString(e);
        delete h;
        return LRPERM_ERR_CUDA;
    }
    h->stream = h->own_stream;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    *out = h;
    return LRPERM_OK;
}

void lrperm_destroy(lrperm_engine* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    h->err_flag_m.release();
    for (DevBuf* b : {&h->indptr, &h->csr_pairs, &h->csc_pairs, &h->err_flag, &h->st_indptr, &h->st_indices, &h->st_data, &h->st_row_keep, &h->st_small,
                      &h->labels32, &h->labels8, &h->labels8s, &h->order, &h->rank, &h->bucket_cluster,
                      &h->cl_start, &h->inv_n, &h->rowA, &h->rowB, &h->thr_c, &h->thr_g, &h->counts_c, &h->counts_g,
                      &h->chunks, &h->csc_keys, &h->csc_keys2, &h->csc_payload, &h->tile_hist, &h->tile_slot_ptr, &h->csc_tmp, &h->slab2, &h->partials2, &h->perm_dev, &h->slab, &h->partials, &h->cube, &h->sort_tmp,
                      &h->tmp_a, &h->tmp_b, &h->tmp_c, &h->tmp_d, &h->stage_out, &h->small_a, &h->small_b,
                      &h->vsc_pairs, &h->tm_chunks, &h->tm_region_ptr, &h->tm_counts, &h->tm_prefix, &h->tm_planes, &h->tm_cube,
                      &h->tm_tgt, &h->tm_roles, &h->thr_d, &h->counts_t, &h->tm_diag, &h->tm_chunk_of_slot, &h->tm_region_dir,
                      &h->tm_walk_flag, &h->tm_walk_list, &h->rs_side, &h->rs_flag, &h->rs_pos, &h->rs_rows, &h->rs_in, &h->needed_dev, &h->chunks_f})
        b->release();
    for (cudaEvent_t ev : h->events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->ev_tile) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->ev_slab_tile) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->ev_main_tile) cudaEventDestroy(ev);
    h->hist_pin.release();
    h->table_pin.release();
    h->up_pin.release();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_matrix) cudaEventDestroy(h->ev_matrix);
    delete h->copier;
    for (int i = 0; i < lrperm_engine::kMaxPin; ++i) { if (h->pin[i]) cudaFreeHost(h->pin[i]); if (h->pin_ev[i]) cudaEventDestroy(h->pin_ev[i]); }
    delete h;
}

int lrperm_set_stream(lrperm_engine* h, void* cuda_stream) {
    if (!h) return LRPERM_ERR_INVALID;
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return LRPERM_OK;
}

int lrperm_set_tuning(lrperm_engine* h, int32_t fast_perm_batch, int32_t fast_chunk_nnz, int32_t fast_tile_mib) {
    if (!h) return LRPERM_ERR_INVALID;
    if (fast_perm_batch < 0 || fast_chunk_nnz < 0 || fast_tile_mib < 0 || fast_tile_mib > 1024 || (fast_perm_batch % 128) != 0)
        return h->fail(LRPERM_ERR_INVALID, "lrperm_set_tuning: fast_perm_batch must be a non-negative multiple of 128");
    h->fast_perm_batch = fast_perm_batch;
    h->fast_chunk_nnz = fast_chunk_nnz;
    h->fast_tile_bytes = fast_tile_mib << 20;
    return LRPERM_OK;
}

int lrperm_set_option(lrperm_engine* h, const char* name, int64_t value) {
    if (!h || !name) return LRPERM_ERR_INVALID;
    const std::string key(name);
    if (key == "fast_variant") {
        if (value != 1 && value != 2 && value != 3) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: fast_variant must be 1, 2 or 3");
        h->fast_variant = (int32_t)value;
    } else if (key == "overlap") {
        if (value != 0 && value != 1) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: overlap must be 0 or 1");
        h->overlap = (int32_t)value;
    } else if (key == "trimean_chunk_nnz") {
        if (value < 0 || value > (1 << 22) || (value % 1024) != 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: trimean_chunk_nnz must be a multiple of 1024 in [0, 2^22]");
        h->tm_chunk_nnz = (int32_t)value;
    } else if (key == "trimean_perm_batch") {
        if (value < 0 || value > 4096 || (value % 128) != 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: trimean_perm_batch must be a multiple of 128 in [0, 4096]");
        h->tm_perm_batch = (int32_t)value;
    } else if (key == "eager_csc") {
        if (value != 0 && value != 1) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: eager_csc must be 0 or 1");
        h->eager_csc = (int32_t)value;
    } else if (key == "score_groups_per_block") {
        if (value < 0 || value > 1024) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: score_groups_per_block must be in [0, 1024]");
        h->score_gpb = (int32_t)value;
    } else if (key == "slab0_by_tile") {
        if (value != 0 && value != 1) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: slab0_by_tile must be 0 or 1");
        h->slab0_by_tile = (int32_t)value;
    } else if (key == "fast_first_batch") {
        if (value < 0 || value > 4096) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: fast_first_batch must be in [0, 4096]");
        h->fast_first_batch = (int32_t)value;
    } else if (key == "slab_mode") {
        if (value < 0 || value > 3) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: slab_mode must be in [0, 3]");
        h->slab_mode = (int32_t)value;
    } else if (key == "slab_ctas_per_sm") {
        if (value < 0 || value > 32) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: slab_ctas_per_sm must be in [0, 32] (0 = auto)");
        h->slab_ctas_per_sm = (int32_t)value;
    } else if (key == "narrow_indices") {
        if (value != 0 && value != 1) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: narrow_indices must be 0 or 1");
        h->narrow_idx = (int32_t)value;
    } else if (key == "skip_unreferenced") {
        if (value != 0 && value != 1) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: skip_unreferenced must be 0 or 1");
        h->skip_unref = (int32_t)value;
    } else if (key == "fast_q") {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: fast_q must be 0, 1, 2, 4 or 8");
        h->fast_q = (int32_t)value;
    } else {
        return h->fail(LRPERM_ERR_INVALID, "lrperm_set_option: unknown option '%s'", name);
    }
    return LRPERM_OK;
}

int lrperm_set_matrix_csr(lrperm_engine* h, int64_t n_cells, int64_t n_genes, const void* indptr, int indptr_is_64,
                          const int32_t* indices, const float* data, int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("set_matrix_csr");
    DeviceGuard guard(h->device);
    h->have_matrix = false; h->csc_ready = false; h->vsc_ready = false; h->have_triples = false; h->have_labels = false;
    if (n_cells < 0 || n_genes <= 0 || !indptr) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_matrix_csr: bad shape or NULL indptr");
    if (n_cells >= 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_set_matrix_csr: n_cells must be < 2^31");
    if (n_genes > 65535) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_set_matrix_csr: n_genes must be <= 65535 (subset to resource genes first)");
    // total stored entries = indptr[N]
    int64_t nnz = 0;
    const size_t isz = indptr_is_64 ? 8 : 4;
    {
        unsigned char last[8] = {0};
        const unsigned char* src = reinterpret_cast<const unsigned char*>(indptr) + isz * (size_t)n_cells;
        if (on_device) { CK(cudaMemcpyAsync(last, src, isz, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream)); }
        else memcpy(last, src, isz);
        if (indptr_is_64) { int64_t v; memcpy(&v, last, 8); nnz = v; } else { int32_t v; memcpy(&v, last, 4); nnz = v; }
    }
    if (nnz < 0 || nnz >= 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_set_matrix_csr: nnz=%lld must be in [0, 2^31)", (long long)nnz);
    if (nnz > 0 && (!indices || !data)) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_matrix_csr: NULL indices/data");
    // an earlier pending upload must not race with this one
    if (h->matrix_pending) { CK(cudaStreamSynchronize(h->copy_stream)); h->matrix_pending = false; CK(cudaMemsetAsync(h->err_flag_m.p, 0, sizeof(int), h->stream)); }
    h->N = n_cells; h->G = n_genes; h->nnz = nnz;

    if (!on_device && nnz > 0 && is_pinned_host(indptr) && is_pinned_host(indices) && is_pinned_host(data)) {
        // Pinned host buffers: nothing waits here.  The row pointers and every cell tile go to the copy stream now; every
        // tile is packed and its CSC copy sorted on that stream right behind its transfer, so lrperm_run can start the
        // first batch on tile 0 while tiles 1.. are still crossing PCIe.  The buffers must stay valid and unchanged until that call returns
        // (as with cudaMemcpyAsync).  Column-range validation is reported by the consuming call.
        cudaStream_t sC = h->copy_stream;
        h->staged = false;
        {   // row pointers are validated here, on the host: device code below indexes with them
            int64_t prev = 0;
            bool ok = true;
            if (indptr_is_64) { const int64_t* ip = static_cast<const int64_t*>(indptr); ok = ip[0] == 0; for (int64_t r = 0; r <= n_cells && ok; ++r) { ok = ip[r] >= prev; prev = ip[r]; } }
            else { const int32_t* ip = static_cast<const int32_t*>(indptr); ok = ip[0] == 0; for (int64_t r = 0; r <= n_cells && ok; ++r) { ok = ip[r] >= prev; prev = ip[r]; } }
            if (!ok) { h->have_matrix = false; return h->fail(LRPERM_ERR_INVALID, "lrperm_set_matrix_csr: invalid input (indptr must start at 0 and be non-decreasing)"); }
        }
        CK(h->indptr.ensure(sizeof(int32_t) * (size_t)(n_cells + 1)));
        CK(h->st_indptr.ensure(isz * (size_t)(n_cells + 1)));
        CK(h->st_indices.ensure(sizeof(int32_t) * (size_t)nnz));
        CK(h->st_data.ensure(sizeof(float) * (size_t)nnz));
        CK(h->csr_pairs.ensure(sizeof(int2) * (size_t)nnz));
        CK(cudaStreamSynchronize(h->stream));          // earlier work on the engine stream may still read the old matrix
        int rc;
        const int32_t tile_cells = fast_tile_cells(h);
        h->csc_ready = false;
        if ((rc = csc_begin(h, tile_cells))) return rc;
        h->up_e.assign((size_t)h->n_tiles + 1, 0);
        for (int32_t t = 0; t <= h->n_tiles; ++t) {
            const int64_t row = std::min<int64_t>((int64_t)t * tile_cells, n_cells);
            h->up_e[(size_t)t] = indptr_is_64 ? static_cast<const int64_t*>(indptr)[row] : (int64_t)static_cast<const int32_t*>(indptr)[row];
        }
        h->up_indices = indices; h->up_data = data; h->up_next_tile = 0;
        h->csr_packed = false;
        CK(cudaMemcpyAsync(h->st_indptr.p, indptr, isz * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, sC));
        if (indptr_is_64)
            narrow_indptr_kernel<int64_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, sC>>>(h->st_indptr.as<int64_t>(), n_cells + 1, h->indptr.as<int32_t>(), h->err_flag_m.as<int>());
        else
            narrow_indptr_kernel<int32_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, sC>>>(h->st_indptr.as<int32_t>(), n_cells + 1, h->indptr.as<int32_t>(), h->err_flag_m.as<int>());
        CK_LAUNCH();
        h->matrix_pending = true;
        h->have_matrix = true;
        // all tiles now: the calls that follow do not use the copy engine while this upload is pending (upload_by_kernel)
        return submit_tiles(h, -1);
    }

    CK(h->indptr.ensure(sizeof(int32_t) * (size_t)(n_cells + 1)));
    if (indptr_is_64) {
        int rc = upload(h, h->tmp_a, indptr, 8 * (size_t)(n_cells + 1), on_device);
        if (rc) return rc;
        narrow_indptr_kernel<int64_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, h->stream>>>(
            h->tmp_a.as<int64_t>(), n_cells + 1, h->indptr.as<int32_t>(), h->err_flag.as<int>());
        CK_LAUNCH();
    } else {
        int rc = upload(h, h->tmp_a, indptr, 4 * (size_t)(n_cells + 1), on_device);
        if (rc) return rc;
        narrow_indptr_kernel<int32_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, h->stream>>>(
            h->tmp_a.as<int32_t>(), n_cells + 1, h->indptr.as<int32_t>(), h->err_flag.as<int>());
        CK_LAUNCH();
    }
    CK(h->csr_pairs.ensure(sizeof(int2) * (size_t)std::max<int64_t>(nnz, 1)));
    if (nnz > 0 && on_device) {
        // Device input (e.g. the sharded multi-GPU upload): keep a raw copy in the staging buffers (the caller's
        // tensors may go away), validate the columns, and leave the packed copy to the first row-gather consumer -
        // FAST mode sorts its CSC copy straight from the raw arrays.
        h->staged = false;
        CK(h->st_indices.ensure(sizeof(int32_t) * (size_t)nnz));
        CK(h->st_data.ensure(sizeof(float) * (size_t)nnz));
        CK(cudaMemcpyAsync(h->st_indices.p, indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->st_data.p, data, sizeof(float) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
        validate_columns_kernel<<<grid_for(nnz, 256, h->sm_count * 16), 256, 0, h->stream>>>(h->st_indices.as<int32_t>(), nnz, (int32_t)n_genes, h->err_flag.as<int>());
        CK_LAUNCH();
        int rc = check_err_flag(h, "lrperm_set_matrix_csr");
        if (rc) return rc;
        h->have_matrix = true;
        h->csr_packed = false;
        if (h->eager_csc && (rc = csc_enqueue_all(h, fast_tile_cells(h)))) return rc;
        return LRPERM_OK;
    }
    if (nnz > 0) {
        const int32_t* d_idx = indices;
        const float* d_val = data;
        if (!on_device) {
            int rc = upload(h, h->tmp_b, indices, sizeof(int32_t) * (size_t)nnz, 0);
            if (rc) return rc;
            rc = upload(h, h->tmp_c, data, sizeof(float) * (size_t)nnz, 0);
            if (rc) return rc;
            d_idx = h->tmp_b.as<int32_t>();
            d_val = h->tmp_c.as<float>();
        }
        pack_csr_kernel<<<grid_for(n_cells * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->indptr.as<int32_t>(), d_idx, d_val, 0, (int32_t)n_cells, (int32_t)n_genes, h->csr_pairs.as<int2>(), h->err_flag.as<int>());
        CK_LAUNCH();
    }
    int rc = check_err_flag(h, "lrperm_set_matrix_csr");
    if (rc) return rc;
    h->have_matrix = true;
    h->csr_packed = true;
    // FAST mode is what normally follows: start the CSC copy now, it sorts while the caller prepares labels / triples
    if (h->eager_csc && nnz > 0 && (rc = csc_enqueue_all(h, fast_tile_cells(h)))) return rc;
    return LRPERM_OK;
}

namespace {
// a matrix whose raw arrays still sit in the staging buffers (pinned upload, not packed yet) is finished first
int stage_begin(lrperm_engine* h) {
    if (h->have_matrix && (h->matrix_pending || !h->csr_packed)) {
        const int rcw = wait_matrix(h, true);
        if (rcw) { h->have_matrix = false; h->matrix_pending = false; h->csr_packed = true; CK(cudaStreamSynchronize(h->copy_stream)); }
        CK(cudaStreamSynchronize(h->stream));
    }
    h->staged = false;
    return LRPERM_OK;
}

// statistics of the staged CSR arrays (st_indptr / st_indices / st_data hold n_cells rows, nnz entries)
int stage_finish(lrperm_engine* h, const char* what, int64_t n_cells, int64_t n_genes, int64_t nnz, const uint8_t* row_keep,
                 int on_device, double* col_sums, double* kept_total, int64_t* n_nonfinite, int64_t* n_empty_rows) {
    int rc;
    h->st_has_keep = row_keep != nullptr;
    if (row_keep && (rc = upload(h, h->st_row_keep, row_keep, (size_t)n_cells, on_device))) return rc;
    // [G doubles col sums][kept_total][n_nonfinite][n_empty_rows][n_unordered]
    const size_t small_bytes = sizeof(double) * ((size_t)n_genes + 4);
    CK(h->st_small.ensure(small_bytes));
    CK(cudaMemsetAsync(h->st_small.p, 0, small_bytes, h->stream));
    double* d_small = h->st_small.as<double>();
    if (n_cells > 0) {
        stage_stats_kernel<<<grid_for(n_cells * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->st_indptr.as<int64_t>(), h->st_indices.as<int32_t>(), h->st_data.as<float>(),
            row_keep ? h->st_row_keep.as<uint8_t>() : nullptr, (int32_t)n_cells, (int32_t)n_genes, d_small,
            d_small + n_genes, reinterpret_cast<unsigned long long*>(d_small + n_genes + 1),
            reinterpret_cast<unsigned long long*>(d_small + n_genes + 2),
            reinterpret_cast<unsigned long long*>(d_small + n_genes + 3), h->err_flag.as<int>());
        CK_LAUNCH();
    }
    std::vector<double> host((size_t)n_genes + 4);
    CK(cudaMemcpyAsync(host.data(), d_small, small_bytes, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = check_err_flag(h, what))) return rc;     // synchronises the stream
    if (col_sums) memcpy(col_sums, host.data(), sizeof(double) * (size_t)n_genes);
    if (kept_total) *kept_total = host[(size_t)n_genes];
    unsigned long long u;
    memcpy(&u, &host[(size_t)n_genes + 1], 8); if (n_nonfinite) *n_nonfinite = (int64_t)u;
    memcpy(&u, &host[(size_t)n_genes + 2], 8); if (n_empty_rows) *n_empty_rows = (int64_t)u;
    memcpy(&u, &host[(size_t)n_genes + 3], 8); h->st_unordered = (int64_t)u;
    h->st_N = n_cells; h->st_G = n_genes; h->st_nnz = nnz;
    h->staged = true;
    return LRPERM_OK;
}
}  // namespace

int lrperm_stage_matrix_csr(lrperm_engine* h, int64_t n_cells, int64_t n_genes, const void* indptr, int indptr_is_64,
                            const int32_t* indices, const float* data, int on_device, const uint8_t* row_keep,
                            double* col_sums, double* kept_total, int64_t* n_nonfinite, int64_t* n_empty_rows) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("stage_matrix_csr");
    DeviceGuard guard(h->device);
    int rc;
    if ((rc = stage_begin(h))) return rc;
    if (n_cells < 0 || n_genes <= 0 || !indptr) return h->fail(LRPERM_ERR_INVALID, "lrperm_stage_matrix_csr: bad shape or NULL indptr");
    if (n_cells >= 0x7fffffffLL || n_genes >= 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_stage_matrix_csr: n_cells and n_genes must be < 2^31");
    int64_t nnz = 0;
    const size_t isz = indptr_is_64 ? 8 : 4;
    {
        unsigned char last[8] = {0};
        const unsigned char* src = reinterpret_cast<const unsigned char*>(indptr) + isz * (size_t)n_cells;
        if (on_device) { CK(cudaMemcpyAsync(last, src, isz, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream)); }
        else memcpy(last, src, isz);
        if (indptr_is_64) { int64_t v; memcpy(&v, last, 8); nnz = v; } else { int32_t v; memcpy(&v, last, 4); nnz = v; }
    }
    // the STAGED matrix may hold 2^31 entries or more (int64 row pointers on the device); the committed one may not
    if (nnz < 0 || nnz >= (1LL << 40)) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_stage_matrix_csr: nnz=%lld must be in [0, 2^40)", (long long)nnz);
    if (nnz > 0 && (!indices || !data)) return h->fail(LRPERM_ERR_INVALID, "lrperm_stage_matrix_csr: NULL indices/data");
    CK(h->st_indptr.ensure(sizeof(int64_t) * (size_t)(n_cells + 1)));
    if ((rc = upload(h, h->tmp_a, indptr, isz * (size_t)(n_cells + 1), on_device))) return rc;
    if (indptr_is_64)
        widen_indptr_kernel<int64_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, h->stream>>>(h->tmp_a.as<int64_t>(), n_cells + 1, h->st_indptr.as<int64_t>(), h->err_flag.as<int>());
    else
        widen_indptr_kernel<int32_t><<<grid_for(n_cells + 1, 256, 1 << 30), 256, 0, h->stream>>>(h->tmp_a.as<int32_t>(), n_cells + 1, h->st_indptr.as<int64_t>(), h->err_flag.as<int>());
    CK_LAUNCH();
    // Column indices of a large pageable upload travel as 16 bits when they fit (n_genes <= 65536): the host threads that fill the
    // pinned staging pieces narrow them on the way (the transfer is bound by host memory traffic: read + staging write + DMA read),
    // a kernel widens them again on the device.  An index that would not survive (negative or >= 65536) is an invalid column.
    const bool narrow = !on_device && h->narrow_idx && n_genes <= 65536 && nnz >= (4LL << 20) && !is_pinned_host(indices);
    if (narrow) {
        CK(h->st_indices.ensure(sizeof(int32_t) * (size_t)nnz));
        CK(h->tmp_b.ensure(sizeof(uint16_t) * (size_t)nnz));
        uint32_t bad = 0;
        if ((rc = upload_pageable(h, h->tmp_b.p, indices, sizeof(uint16_t) * (size_t)nnz, true, &bad))) return rc;
        if (bad) return h->fail(LRPERM_ERR_INVALID, "lrperm_stage_matrix_csr: invalid input (flag=1: 1=column out of range, 2=indptr overflow, "
                                "4=indptr not monotone, 8=triple index out of range)");
        widen_u16_kernel<<<grid_for(nnz, 256, h->sm_count * 16), 256, 0, h->stream>>>(h->tmp_b.as<uint16_t>(), nnz, h->st_indices.as<int32_t>());
        CK_LAUNCH();
    } else if ((rc = upload(h, h->st_indices, indices, sizeof(int32_t) * (size_t)nnz, on_device))) return rc;
    if ((rc = upload(h, h->st_data, data, sizeof(float) * (size_t)nnz, on_device))) return rc;
    return stage_finish(h, "lrperm_stage_matrix_csr", n_cells, n_genes, nnz, row_keep, on_device, col_sums, kept_total,
                        n_nonfinite, n_empty_rows);
}

int lrperm_stage_matrix_dense(lrperm_engine* h, int64_t n_cells, int64_t n_genes, const void* values, int values_are_f64,
                              int64_t row_stride, int on_device, const uint8_t* row_keep,
                              double* col_sums, double* kept_total, int64_t* n_nonfinite, int64_t* n_empty_rows) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("stage_matrix_dense");
    DeviceGuard guard(h->device);
    int rc;
    if ((rc = stage_begin(h))) return rc;
    if (n_cells < 0 || n_genes <= 0 || (n_cells > 0 && !values) || row_stride < n_genes)
        return h->fail(LRPERM_ERR_INVALID, "lrperm_stage_matrix_dense: bad shape, NULL values or row_stride < n_genes");
    if (n_cells >= 0x7fffffffLL || n_genes >= 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_stage_matrix_dense: n_cells and n_genes must be < 2^31");
    const size_t esz = values_are_f64 ? 8 : 4;
    // the dense rows sit on the device only while they are converted (host input: uploaded whole, released below)
    DevBuf dense;
    const void* X = values;
    if (!on_device && n_cells > 0) {
        const size_t bytes = esz * ((size_t)(n_cells - 1) * (size_t)row_stride + (size_t)n_genes);
        if ((rc = upload(h, dense, values, bytes, 0))) { dense.release(); return rc; }
        X = dense.p;
    }
    struct Release { DevBuf& b; cudaStream_t s; ~Release() { if (b.p) { cudaStreamSynchronize(s); b.release(); } } } release{dense, h->stream};
    CK(h->st_indptr.ensure(sizeof(int64_t) * (size_t)(n_cells + 1)));
    CK(h->tmp_a.ensure(sizeof(int64_t) * (size_t)(n_cells + 1)));
    CK(h->small_b.ensure(16));
    CK(cudaMemsetAsync(h->small_b.p, 0, 16, h->stream));
    CK(cudaMemsetAsync(h->tmp_a.p, 0, sizeof(int64_t), h->stream));
    const int grid = grid_for(std::max<int64_t>(n_cells, 1) * 32, 256, h->sm_count * 32);
    if (n_cells > 0) {
        if (values_are_f64)
            dense_count_kernel<double><<<grid, 256, 0, h->stream>>>(static_cast<const double*>(X), row_stride, (int32_t)n_cells, (int32_t)n_genes, h->tmp_a.as<int64_t>(), h->small_b.as<unsigned long long>());
        else
            dense_count_kernel<float><<<grid, 256, 0, h->stream>>>(static_cast<const float*>(X), row_stride, (int32_t)n_cells, (int32_t)n_genes, h->tmp_a.as<int64_t>(), h->small_b.as<unsigned long long>());
        CK_LAUNCH();
    }
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, h->small_b.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int64_t nnz = (int64_t)total;
    {
        // inclusive scan of counts[1..N] with counts[0] = 0 == exclusive row pointers (int64: no 2^31 limit here)
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, h->tmp_a.as<int64_t>(), h->st_indptr.as<int64_t>(), (int)(n_cells + 1), h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::InclusiveSum(h->sort_tmp.p, tmp_bytes, h->tmp_a.as<int64_t>(), h->st_indptr.as<int64_t>(), (int)(n_cells + 1), h->stream));
    }
    CK(h->st_indices.ensure(sizeof(int32_t) * (size_t)nnz));
    CK(h->st_data.ensure(sizeof(float) * (size_t)nnz));
    if (n_cells > 0 && nnz > 0) {
        if (values_are_f64)
            dense_fill_kernel<double><<<grid, 256, 0, h->stream>>>(static_cast<const double*>(X), row_stride, (int32_t)n_cells, (int32_t)n_genes, h->st_indptr.as<int64_t>(), h->st_indices.as<int32_t>(), h->st_data.as<float>());
        else
            dense_fill_kernel<float><<<grid, 256, 0, h->stream>>>(static_cast<const float*>(X), row_stride, (int32_t)n_cells, (int32_t)n_genes, h->st_indptr.as<int64_t>(), h->st_indices.as<int32_t>(), h->st_data.as<float>());
        CK_LAUNCH();
    }
    return stage_finish(h, "lrperm_stage_matrix_dense", n_cells, n_genes, nnz, row_keep, on_device, col_sums, kept_total,
                        n_nonfinite, n_empty_rows);
}

int lrperm_staged_info(lrperm_engine* h, int64_t* n_cells, int64_t* n_genes, int64_t* nnz, int64_t* n_unordered) {
    if (!h) return LRPERM_ERR_INVALID;
    if (!h->staged) return h->fail(LRPERM_ERR_INVALID, "lrperm_staged_info: call lrperm_stage_matrix_csr first");
    if (n_cells) *n_cells = h->st_N;
    if (n_genes) *n_genes = h->st_G;
    if (nnz) *nnz = h->st_nnz;
    if (n_unordered) *n_unordered = h->st_unordered;
    return LRPERM_OK;
}

int lrperm_staged_max(lrperm_engine* h, float* max_stored, int64_t* n_stored) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->staged) return h->fail(LRPERM_ERR_INVALID, "lrperm_staged_max: call lrperm_stage_matrix_csr first");
    // [uint32 max key][pad][uint64 stored entries]
    CK(h->small_b.ensure(16));
    CK(cudaMemsetAsync(h->small_b.p, 0, 16, h->stream));
    if (h->st_N > 0 && h->st_nnz > 0) {
        stage_max_kernel<<<grid_for(h->st_N * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->st_indptr.as<int64_t>(), h->st_data.as<float>(), h->st_has_keep ? h->st_row_keep.as<uint8_t>() : nullptr,
            (int32_t)h->st_N, h->small_b.as<uint32_t>(), reinterpret_cast<unsigned long long*>(h->small_b.as<unsigned char>() + 8));
        CK_LAUNCH();
    }
    unsigned char host[16];
    CK(cudaMemcpyAsync(host, h->small_b.p, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    uint32_t key; unsigned long long cnt;
    memcpy(&key, host, 4); memcpy(&cnt, host + 8, 8);
    if (max_stored) {
        if (key == 0) *max_stored = -INFINITY;           // no stored entry in the kept rows
        else { const uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key; memcpy(max_stored, &b, 4); }
    }
    if (n_stored) *n_stored = (int64_t)cnt;
    return LRPERM_OK;
}

int lrperm_staged_cluster_stats(lrperm_engine* h, const int32_t* row_label, int32_t n_clusters, double* sum_x, double* sum_sq) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->staged) return h->fail(LRPERM_ERR_INVALID, "lrperm_staged_cluster_stats: call lrperm_stage_matrix_csr first");
    if (!row_label || n_clusters <= 0 || n_clusters > 2048 || !sum_x || !sum_sq)
        return h->fail(LRPERM_ERR_INVALID, "lrperm_staged_cluster_stats: bad arguments (at most 2048 clusters)");
    int rc;
    const size_t L = (size_t)n_clusters;
    if ((rc = upload(h, h->tmp_d, row_label, sizeof(int32_t) * (size_t)std::max<int64_t>(h->st_N, 1), 0))) return rc;
    CK(h->small_b.ensure(sizeof(double) * 2 * L));
    CK(cudaMemsetAsync(h->small_b.p, 0, sizeof(double) * 2 * L, h->stream));
    if (h->st_N > 0 && h->st_nnz > 0) {
        const int n_blocks = (int)std::min<int64_t>(h->st_N, (int64_t)h->sm_count * 16);
        CK(h->tmp_a.ensure(sizeof(double) * 2 * L * (size_t)n_blocks));
        staged_cluster_stats_kernel<<<n_blocks, 32, sizeof(double) * 2 * L, h->stream>>>(
            h->st_indptr.as<int64_t>(), h->st_data.as<float>(), h->tmp_d.as<int32_t>(), (int32_t)h->st_N, n_clusters, h->tmp_a.as<double>());
        CK_LAUNCH();
        staged_cluster_stats_reduce_kernel<<<grid_for((int64_t)(2 * L), 128, 1 << 30), 128, 0, h->stream>>>(
            h->tmp_a.as<double>(), n_blocks, (int32_t)(2 * L), h->small_b.as<double>());
        CK_LAUNCH();
    }
    if ((rc = download(h, sum_x, h->small_b.p, sizeof(double) * L, 0))) return rc;
    if ((rc = download(h, sum_sq, h->small_b.as<double>() + L, sizeof(double) * L, 0))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_commit_matrix(lrperm_engine* h, const int32_t* col_map, int64_t n_genes_out) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("commit_matrix");
    DeviceGuard guard(h->device);
    if (!h->staged) return h->fail(LRPERM_ERR_INVALID, "lrperm_commit_matrix: call lrperm_stage_matrix_csr first");
    if (!col_map || n_genes_out <= 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_commit_matrix: NULL col_map or n_genes_out <= 0");
    if (n_genes_out > 65535) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_commit_matrix: n_genes_out must be <= 65535");
    const int64_t N = h->st_N, Gin = h->st_G;
    {   // every output column has at most one source column (distinct columns per row is an engine invariant)
        std::vector<uint8_t> seen((size_t)n_genes_out, 0);
        for (int64_t g = 0; g < Gin; ++g) {
            const int32_t m = col_map[g];
            if (m < -1 || m >= n_genes_out) return h->fail(LRPERM_ERR_INVALID, "lrperm_commit_matrix: col_map[%lld]=%d outside [-1,%lld)", (long long)g, m, (long long)n_genes_out);
            if (m >= 0) {
                if (seen[(size_t)m]) return h->fail(LRPERM_ERR_INVALID, "lrperm_commit_matrix: output column %d has two source columns", m);
                seen[(size_t)m] = 1;
            }
        }
    }
    h->have_matrix = false; h->csc_ready = false; h->vsc_ready = false; h->have_triples = false; h->have_labels = false;
    int rc;
    DevBuf &d_map = h->small_a, &d_flag = h->tmp_a, &d_excl = h->tmp_d, &d_newrow = h->stage_out;
    if ((rc = upload(h, d_map, col_map, sizeof(int32_t) * (size_t)Gin, 0))) return rc;
    CK(d_flag.ensure(sizeof(int32_t) * (size_t)(N + 1)));
    CK(d_excl.ensure(sizeof(int32_t) * (size_t)(N + 1)));
    CK(d_newrow.ensure(sizeof(int32_t) * (size_t)(N + 1)));
    int64_t N_out = N;
    if (N > 0) {
        row_keep_to_flag_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, h->stream>>>(h->st_has_keep ? h->st_row_keep.as<uint8_t>() : nullptr, (int32_t)N, d_flag.as<int32_t>());
        CK_LAUNCH();
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag.as<int32_t>(), d_excl.as<int32_t>(), (int)N, h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(h->sort_tmp.p, tmp_bytes, d_flag.as<int32_t>(), d_excl.as<int32_t>(), (int)N, h->stream));
        new_row_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, h->stream>>>(d_flag.as<int32_t>(), d_excl.as<int32_t>(), (int32_t)N, d_newrow.as<int32_t>());
        CK_LAUNCH();
        int32_t last_excl = 0, last_flag = 0;
        CK(cudaMemcpyAsync(&last_excl, d_excl.as<int32_t>() + (N - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&last_flag, d_flag.as<int32_t>() + (N - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        N_out = (int64_t)last_excl + last_flag;
    }
    // per-row counts of surviving entries -> exclusive scan -> indptr of the engine matrix
    DevBuf& d_counts = h->tmp_b;
    CK(d_counts.ensure(sizeof(int32_t) * (size_t)(N_out + 1)));
    CK(cudaMemsetAsync(d_counts.p, 0, sizeof(int32_t) * (size_t)(N_out + 1), h->stream));
    CK(h->indptr.ensure(sizeof(int32_t) * (size_t)(N_out + 1)));
    int64_t nnz_out = 0;
    if (N > 0) {
        CK(h->small_b.ensure(16));
        CK(cudaMemsetAsync(h->small_b.p, 0, 8, h->stream));
        mapped_count_kernel<<<grid_for(N * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->st_indptr.as<int64_t>(), h->st_indices.as<int32_t>(), d_map.as<int32_t>(), d_newrow.as<int32_t>(), (int32_t)N, d_counts.as<int32_t>(),
            h->small_b.as<unsigned long long>());
        CK_LAUNCH();
        unsigned long long kept = 0;
        CK(cudaMemcpyAsync(&kept, h->small_b.p, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (kept >= 0x7fffffffULL)
            return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_commit_matrix: %llu entries survive the row / column filter, must be < 2^31", kept);
    }
    {
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts.as<int32_t>(), h->indptr.as<int32_t>(), (int)(N_out + 1), h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(h->sort_tmp.p, tmp_bytes, d_counts.as<int32_t>(), h->indptr.as<int32_t>(), (int)(N_out + 1), h->stream));
        int32_t total = 0;
        CK(cudaMemcpyAsync(&total, h->indptr.as<int32_t>() + N_out, 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        nnz_out = total;
    }
    CK(h->tmp_c.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(nnz_out, 1)));
    CK(h->tmp_d.ensure(sizeof(float) * (size_t)std::max<int64_t>(nnz_out, 1)));     // d_excl is dead from here on
    CK(h->csr_pairs.ensure(sizeof(int2) * (size_t)std::max<int64_t>(nnz_out, 1)));
    if (N > 0 && nnz_out > 0) {
        mapped_scatter_kernel<<<grid_for(N * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->st_indptr.as<int64_t>(), h->st_indices.as<int32_t>(), h->st_data.as<float>(), d_map.as<int32_t>(), d_newrow.as<int32_t>(),
            (int32_t)N, h->indptr.as<int32_t>(), h->tmp_c.as<int32_t>(), h->tmp_d.as<float>());
        CK_LAUNCH();
        pack_csr_kernel<<<grid_for(N_out * 32, 256, h->sm_count * 32), 256, 0, h->stream>>>(
            h->indptr.as<int32_t>(), h->tmp_c.as<int32_t>(), h->tmp_d.as<float>(), 0, (int32_t)N_out, (int32_t)n_genes_out, h->csr_pairs.as<int2>(), h->err_flag.as<int>());
        CK_LAUNCH();
    }
    if ((rc = check_err_flag(h, "lrperm_commit_matrix"))) return rc;
    h->N = N_out; h->G = n_genes_out; h->nnz = nnz_out;
    h->have_matrix = true;
    h->csr_packed = true;
    if (h->eager_csc && nnz_out > 0 && (rc = csc_enqueue_all(h, fast_tile_cells(h)))) return rc;
    return LRPERM_OK;
}

int lrperm_set_labels(lrperm_engine* h, const int32_t* labels, int32_t n_clusters, int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("set_labels");
    DeviceGuard guard(h->device);
    h->have_labels = false;
    h->up_pin_off = 0;
    if (!h->have_matrix) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_labels: call lrperm_set_matrix_csr first");
    if (!labels || n_clusters <= 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_labels: NULL labels or n_clusters <= 0");
    const int64_t N = h->N;
    const size_t N1 = (size_t)std::max<int64_t>(N, 1);
    int rc;
    // Everything that scales with N runs on the device: cluster sizes (histogram), positions grouped by cluster,
    // ascending within a cluster (stable radix sort of the labels: the reference's labels_mask order,
    // _get_mean_perms.py:47-52 + boolean row selection at :63), their inverse, byte labels.  The host only sees the
    // L cluster sizes.
    if ((rc = upload(h, h->labels32, labels, sizeof(int32_t) * (size_t)N, on_device))) return rc;
    CK(h->labels8.ensure(N1));
    CK(h->order.ensure(sizeof(int32_t) * N1));
    CK(h->rank.ensure(sizeof(int32_t) * N1));
    DevBuf &d_small = h->small_b, &d_keys = h->tmp_a, &d_keys2 = h->tmp_c, &d_iota = h->tmp_d;
    CK(d_small.ensure(sizeof(int32_t) * ((size_t)n_clusters + 1)));          // [L] counts, [L] first out-of-range cell
    CK(d_keys.ensure(sizeof(uint32_t) * N1));
    CK(d_keys2.ensure(sizeof(uint32_t) * N1));
    CK(d_iota.ensure(sizeof(int32_t) * N1));
    CK(cudaMemsetAsync(d_small.p, 0, sizeof(int32_t) * (size_t)n_clusters, h->stream));
    CK(cudaMemsetAsync(d_small.as<int32_t>() + n_clusters, 0x7f, sizeof(int32_t), h->stream));
    std::vector<int32_t> cnt((size_t)n_clusters + 1, 0);
    if (N > 0) {
        const size_t smem = n_clusters <= 4096 ? sizeof(int32_t) * (size_t)n_clusters : 0;
        label_hist_kernel<<<grid_for(N, 256, h->sm_count * 8), 256, smem, h->stream>>>(
            h->labels32.as<int32_t>(), (int32_t)N, n_clusters, d_small.as<int32_t>(), d_small.as<int32_t>() + n_clusters,
            d_keys.as<uint32_t>(), d_iota.as<int32_t>());
        CK_LAUNCH();
        int end_bit = 1;
        while ((1 << end_bit) < n_clusters && end_bit < 31) ++end_bit;
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_iota.as<int32_t>(),
                                           h->order.as<int32_t>(), (int)N, 0, end_bit, h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        tmp_bytes = h->sort_tmp.cap;
        CK(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, d_keys.as<uint32_t>(), d_keys2.as<uint32_t>(), d_iota.as<int32_t>(),
                                           h->order.as<int32_t>(), (int)N, 0, end_bit, h->stream));
        label_tables_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, h->stream>>>(h->labels32.as<int32_t>(), h->order.as<int32_t>(), (int32_t)N,
                                                                            h->rank.as<int32_t>(), h->labels8.as<uint8_t>());
        CK_LAUNCH();
    }
    CK(cudaMemcpyAsync(cnt.data(), d_small.p, sizeof(int32_t) * ((size_t)n_clusters + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (N > 0 && cnt[(size_t)n_clusters] < N) {
        const int64_t j = cnt[(size_t)n_clusters];
        int32_t c = 0;
        CK(cudaMemcpy(&c, h->labels32.as<int32_t>() + j, sizeof(int32_t), cudaMemcpyDeviceToHost));
        return h->fail(LRPERM_ERR_INVALID, "lrperm_set_labels: label %d at cell %lld outside [0,%d)", c, (long long)j, n_clusters);
    }
    std::vector<int32_t> start((size_t)n_clusters + 1, 0);
    for (int32_t c = 0; c < n_clusters; ++c) {
        if (cnt[(size_t)c] == 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_labels: cluster %d has no cells (drop unused categories first)", c);
        start[(size_t)c + 1] = start[(size_t)c] + cnt[(size_t)c];
    }
    std::vector<float> inv((size_t)n_clusters);
    for (int32_t c = 0; c < n_clusters; ++c) inv[(size_t)c] = (float)(1.0 / (double)(start[(size_t)c + 1] - start[(size_t)c]));
    // PHILOX permutations live on cluster-sorted positions: a coarse table (<= 4096 buckets of 2^shift sorted
    // positions) gives the cluster of a sorted position
    int32_t shift = 0;
    while (((N - 1) >> shift) >= 4096) ++shift;
    const int32_t n_buckets = (int32_t)(((N > 0 ? N - 1 : 0) >> shift) + 1);
    std::vector<uint8_t> bucket((size_t)n_buckets);
    {
        int32_t c = 0;
        for (int32_t b = 0; b < n_buckets; ++b) {
            const int64_t k0 = (int64_t)b << shift;
            while (c + 1 < n_clusters && k0 >= start[(size_t)c + 1]) ++c;
            bucket[(size_t)b] = (uint8_t)(c & 0xff);
        }
    }
    h->L = n_clusters;
    h->tm_tables_ready = false;
    h->labels8s_scale = 0;
    h->h_cl_start = start;
    if ((rc = upload(h, h->bucket_cluster, bucket.data(), (size_t)n_buckets, 0))) return rc;
    h->n_buckets = n_buckets;
    h->bucket_shift = shift;
    if ((rc = upload(h, h->cl_start, start.data(), sizeof(int32_t) * ((size_t)n_clusters + 1), 0))) return rc;
    if ((rc = upload(h, h->inv_n, inv.data(), sizeof(float) * (size_t)n_clusters, 0))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->have_labels = true;
    h->have_triples = false;
    return LRPERM_OK;
}

int lrperm_set_triples(lrperm_engine* h, int64_t n_triples, const int32_t* ligand_idx, const int32_t* receptor_idx,
                       const int32_t* source_idx, const int32_t* target_idx, const float* obs_cpdb, const float* obs_gmean,
                       int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("set_triples");
    DeviceGuard guard(h->device);
    h->have_triples = false;
    h->up_pin_off = 0;
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_triples: set the matrix and labels first");
    if (n_triples < 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_set_triples: n_triples < 0");
    if (n_triples > 0 && (!ligand_idx || !receptor_idx || !source_idx || !target_idx))
        return h->fail(LRPERM_ERR_INVALID, "lrperm_set_triples: NULL index array");
    const size_t T = (size_t)n_triples;
    int rc;
    CK(h->rowA.ensure(sizeof(int32_t) * std::max<size_t>(T, 1)));
    CK(h->rowB.ensure(sizeof(int32_t) * std::max<size_t>(T, 1)));
    CK(h->counts_c.ensure(sizeof(uint32_t) * std::max<size_t>(T, 1)));
    CK(h->counts_g.ensure(sizeof(uint32_t) * std::max<size_t>(T, 1)));
    CK(h->thr_c.ensure(sizeof(float) * std::max<size_t>(T, 1)));
    CK(h->thr_g.ensure(sizeof(float) * std::max<size_t>(T, 1)));
    if (T > 0) {
        const int32_t *dl = ligand_idx, *dr = receptor_idx, *dsrc = source_idx, *dt = target_idx;
        if (!on_device) {
            CK(h->tmp_a.ensure(sizeof(int32_t) * T * 4));
            int32_t* base = h->tmp_a.as<int32_t>();
            const int32_t* srcs[4] = {ligand_idx, receptor_idx, source_idx, target_idx};
            for (int k = 0; k < 4; ++k) {
                if (h->matrix_pending) { if ((rc = upload_by_kernel(h, base + (size_t)k * T, srcs[k], 4 * T, h->stream))) return rc; }
                else CK(cudaMemcpyAsync(base + (size_t)k * T, srcs[k], 4 * T, cudaMemcpyHostToDevice, h->stream));
            }
            dl = base; dr = base + T; dsrc = base + 2 * T; dt = base + 3 * T;
        }
        build_rows_kernel<<<grid_for((int64_t)T, 256, 1 << 30), 256, 0, h->stream>>>(
            dl, dr, dsrc, dt, (int64_t)T, (int32_t)h->G, h->L, h->rowA.as<int32_t>(), h->rowB.as<int32_t>(), h->err_flag.as<int>());
        CK_LAUNCH();
        // genes some triple reads (FAST mode skips the others when no cube is exported)
        CK(h->needed_dev.ensure((size_t)h->G));
        CK(cudaMemsetAsync(h->needed_dev.p, 0, (size_t)h->G, h->stream));
        mark_genes_kernel<<<grid_for((int64_t)T, 256, h->sm_count * 8), 256, 0, h->stream>>>(dl, dr, (int64_t)T, (int32_t)h->G, h->needed_dev.as<uint8_t>());
        CK_LAUNCH();
        h->h_needed.assign((size_t)h->G, 0);
        CK(cudaMemcpyAsync(h->h_needed.data(), h->needed_dev.p, (size_t)h->G, cudaMemcpyDeviceToHost, h->stream));
        if (obs_cpdb && (rc = upload(h, h->thr_c, obs_cpdb, sizeof(float) * T, on_device))) return rc;
        if (obs_gmean && (rc = upload(h, h->thr_g, obs_gmean, sizeof(float) * T, on_device))) return rc;
    }
    h->needed_valid = false;
    rc = check_err_flag(h, "lrperm_set_triples");      // synchronises the stream: h_needed has arrived
    if (rc) return rc;
    if (T == 0) { CK(h->needed_dev.ensure((size_t)h->G)); CK(cudaMemsetAsync(h->needed_dev.p, 0, (size_t)h->G, h->stream)); h->h_needed.assign((size_t)h->G, 0); }
    h->needed_valid = true;
    ++h->needed_version;
    h->T = n_triples;
    h->have_thr_c = obs_cpdb != nullptr;
    h->have_thr_g = obs_gmean != nullptr;
    h->have_triples = true;
    return LRPERM_OK;
}

int lrperm_observed(lrperm_engine* h, float* means, int32_t* stored_counts, int32_t* cluster_sizes, int out_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_observed: set the matrix and labels first");
    { const int rcw = wait_matrix(h); if (rcw) return rcw; }
    const int32_t G = (int32_t)h->G, L = h->L;
    const size_t LG = (size_t)L * G;
    int rc;
    if (means) {
        // identity permutation through the exact-order kernel: cube[g][c][0]
        CK(h->cube.ensure(sizeof(float) * LG));
        CK(h->stage_out.ensure(sizeof(float) * LG));
        PermSource ps{}; ps.identity = 1;
        if ((rc = dbg_set_dims(h, 1, 0))) return rc;
        if ((rc = launch_exact(h, ps, 1, h->cube.as<float>(), 1))) return rc;
        dim3 blk(32, 8), grd((G + 31) / 32, 1, L);
        cube_export_kernel<<<grd, blk, 0, h->stream>>>(h->cube.as<float>(), G, L, 1, 1, h->stage_out.as<float>());
        CK_LAUNCH();
        if ((rc = download(h, means, h->stage_out.p, sizeof(float) * LG, out_on_device))) return rc;
    }
    if (stored_counts) {
        CK(h->tmp_a.ensure(sizeof(int32_t) * LG));
        CK(cudaMemsetAsync(h->tmp_a.p, 0, sizeof(int32_t) * LG, h->stream));
        if (h->nnz > 0) {
            stored_counts_kernel<<<grid_for(h->N * 32, 256), 256, 0, h->stream>>>(
                h->indptr.as<int32_t>(), h->csr_pairs.as<int2>(), h->labels32.as<int32_t>(), (int32_t)h->N, G, h->tmp_a.as<int32_t>());
            CK_LAUNCH();
        }
        if ((rc = download(h, stored_counts, h->tmp_a.p, sizeof(int32_t) * LG, out_on_device))) return rc;
    }
    if (cluster_sizes) {
        std::vector<int32_t> sz((size_t)L);
        for (int32_t c = 0; c < L; ++c) sz[(size_t)c] = h->h_cl_start[(size_t)c + 1] - h->h_cl_start[(size_t)c];
        if (out_on_device) CK(cudaMemcpyAsync(cluster_sizes, sz.data(), sizeof(int32_t) * (size_t)L, cudaMemcpyHostToDevice, h->stream));
        else memcpy(cluster_sizes, sz.data(), sizeof(int32_t) * (size_t)L);
    }
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_cluster_moments(lrperm_engine* h, double logbase, double* sum_x, double* sum_sq, double* sum_expm1, int out_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_cluster_moments: set the matrix and labels first");
    { const int rcw = wait_matrix(h); if (rcw) return rcw; }
    if (!(logbase > 0.0)) return h->fail(LRPERM_ERR_INVALID, "lrperm_cluster_moments: logbase must be > 0");
    const int32_t G = (int32_t)h->G, L = h->L;
    const size_t LG = (size_t)L * G;
    const int32_t n_seg = 16;
    CK(h->tmp_a.ensure(sizeof(double) * 3 * LG * n_seg));
    CK(h->tmp_b.ensure(sizeof(double) * 3 * LG));
    const size_t smem = sizeof(double) * 3 * kMomentTile;
    CK(cudaFuncSetAttribute(cluster_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grd((unsigned)n_seg, (unsigned)L, (unsigned)((G + kMomentTile - 1) / kMomentTile));
    cluster_moments_kernel<<<grd, 32, smem, h->stream>>>(h->indptr.as<int32_t>(), h->csr_pairs.as<int2>(), h->order.as<int32_t>(),
                                                         h->cl_start.as<int32_t>(), G, n_seg, std::log(logbase), h->tmp_a.as<double>(), L);
    CK_LAUNCH();
    moments_reduce_kernel<<<grid_for((int64_t)3 * LG, 256, 1 << 30), 256, 0, h->stream>>>(h->tmp_a.as<double>(), G, L, n_seg, h->tmp_b.as<double>());
    CK_LAUNCH();
    int rc;
    const double* base = h->tmp_b.as<double>();
    if ((rc = download(h, sum_x, base, sizeof(double) * LG, out_on_device))) return rc;
    if ((rc = download(h, sum_sq, base + LG, sizeof(double) * LG, out_on_device))) return rc;
    if ((rc = download(h, sum_expm1, base + 2 * LG, sizeof(double) * LG, out_on_device))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_philox_perms(lrperm_engine* h, int64_t n_cells, int64_t n_perms, int64_t perm_offset, uint64_t seed,
                        int32_t* out, int out_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (n_cells <= 0 || n_cells >= 0x7fffffffLL || n_perms < 0 || !out) return h->fail(LRPERM_ERR_INVALID, "lrperm_philox_perms: bad arguments");
    if (!h->have_labels || n_cells != h->N)
        return h->fail(LRPERM_ERR_INVALID, "lrperm_philox_perms: set the matrix and labels first (the permutations are defined on "
                       "cluster-sorted positions) and pass n_cells == the engine's cell count");
    if (n_perms == 0) return LRPERM_OK;
    if (n_perms > 65535) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_philox_perms: at most 65535 permutations per call");
    const size_t bytes = sizeof(int32_t) * (size_t)n_cells * (size_t)n_perms;
    int32_t* dst = out;
    if (!out_on_device) { CK(h->stage_out.ensure(bytes)); dst = h->stage_out.as<int32_t>(); }
    dim3 grd(grid_for(n_cells, 256, 1024), (unsigned)n_perms);
    philox_perms_kernel<<<grd, 256, 0, h->stream>>>(seed, perm_offset, (int32_t)n_cells, (int32_t)n_perms, h->rank.as<int32_t>(), dst);
    CK_LAUNCH();
    if (!out_on_device) CK(cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_last_timings(lrperm_engine* h, float* out8) {
    if (!h || !out8) return LRPERM_ERR_INVALID;
    out8[0] = h->t_setup; out8[1] = h->t_means; out8[2] = h->t_score; out8[3] = h->t_total; out8[4] = (float)h->launches;
    out8[5] = h->t_main; out8[6] = (float)h->n_main_launches; out8[7] = (float)h->last_pb;
    return LRPERM_OK;
}

int lrperm_run(lrperm_engine* h, int64_t n_perms, int64_t perm_offset, uint64_t seed, int perm_source,
               const void* perm_idx, int perm_idx_is_64, int perm_idx_on_device, int order_mode, int scores, int gmean_mode,
               uint32_t* counts_cpdb, uint32_t* counts_gmean, int counts_on_device, float* cube_out, int cube_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("run");
    DeviceGuard guard(h->device);
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: set the matrix and labels first");
    if (n_perms < 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: n_perms < 0");
    if (perm_source != LRPERM_PERMS_INDEX && perm_source != LRPERM_PERMS_PHILOX) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: unknown perm_source %d", perm_source);
    if (order_mode != LRPERM_ORDER_EXACT && order_mode != LRPERM_ORDER_FAST) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: unknown order_mode %d", order_mode);
    if (perm_source == LRPERM_PERMS_INDEX && n_perms > 0 && !perm_idx) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: perm_idx is NULL");
    if (scores & ~(LRPERM_SCORE_CPDB | LRPERM_SCORE_GMEAN)) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: unknown score bits %d", scores);
    if (scores && !h->have_triples) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: scores requested but no triples set");
    if ((scores & LRPERM_SCORE_CPDB) && (!h->have_thr_c || !counts_cpdb)) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: CPDB requested without obs_cpdb / counts_cpdb");
    if ((scores & LRPERM_SCORE_GMEAN) && (!h->have_thr_g || !counts_gmean)) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: GMEAN requested without obs_gmean / counts_gmean");
    if (gmean_mode != LRPERM_GMEAN_PRODUCT && gmean_mode != LRPERM_GMEAN_LITERAL) return h->fail(LRPERM_ERR_INVALID, "lrperm_run: unknown gmean_mode %d", gmean_mode);

    const int64_t N = h->N;
    const int32_t G = (int32_t)h->G, L = h->L;
    const int64_t T = h->T;
    h->launches = 0;
    int rc;

    // ---- batch geometry ----
    int32_t PB;   // permutations per batch (cube row stride PP == PB)
    int Q = 0;
    int32_t fast_chunk = 0, fast_tile = 0;
    if (order_mode == LRPERM_ORDER_FAST) {
        if (L > 255) return h->fail(LRPERM_ERR_UNSUPPORTED, "FAST mode supports at most 255 clusters (got %d); use EXACT", L);
        Q = h->fast_variant >= 2 ? pick_q2(h) : pick_q(h);
        if (Q == 0) return h->fail(LRPERM_ERR_UNSUPPORTED, "FAST mode: %d clusters do not fit the shared-memory bins", L);
        PB = h->fast_perm_batch ? h->fast_perm_batch : 512;
        const int32_t gran = std::max(128, 32 * Q);        // a warp covers 32*Q permutations
        // (cutting a run that would be ONE batch - one rank's share on 2-4 GPUs - in two, so that the second half's slab hides behind the
        // first half's main kernel, was measured and NOT adopted: 500 permutations 13.4 -> 12.4-12.9 ms in most runs but 14.3-14.7 ms in
        // others, depending on what ran before - scripts/overlap_ab.py; `lrperm_set_tuning(256, ...)` asks for it explicitly)
        if (n_perms < PB) PB = (int32_t)std::max<int64_t>(gran, ((n_perms + gran - 1) / gran) * gran);
        PB = ((PB + gran - 1) / gran) * gran;
        const int32_t tile_cells = fast_tile_cells(h);
        // entries per chunk: long chunks amortise the bin zeroing / partial-sum write-out, but the grid should
        // still hold ~16 waves of (chunk x permutation group) warps; 4096 for configs[1], 16384 for configs[2]
        int32_t chunk_nnz = h->fast_chunk_nnz;
        if (!chunk_nnz) {
            // (for the nominal 4 permutations per lane whatever Q runs: the chunk table defines the summation order, and
            // FAST results must not depend on tuning knobs or on the number of GPUs)
            const int64_t groups = 512 / (32 * 4);
            const int64_t target_items = (int64_t)h->sm_count * 8 * 16;
            int64_t c = h->nnz * groups / std::max<int64_t>(target_items, 1);
            c = std::min<int64_t>(std::max<int64_t>(c, 4096), 16384);
            chunk_nnz = (int32_t)((c + 1023) / 1024 * 1024);
        }
        fast_chunk = chunk_nnz;
        fast_tile = tile_cells;
    } else {
        if ((rc = wait_matrix(h))) return rc;
        if ((rc = dbg_set_dims(h, 1, 0))) return rc;
        // enough (perm, cluster) warps for several waves, cube batch <= ~256 MB
        PB = 512;
        while (PB > 128 && (int64_t)PB * L * G * 4 > (256LL << 20)) PB >>= 1;
    }
    if (order_mode != LRPERM_ORDER_FAST && n_perms < PB) PB = (int32_t)std::max<int64_t>(128, ((n_perms + 127) / 128) * 128);
    const int32_t PP = PB;
    h->last_pb = PB;
    CK(h->cube.ensure(sizeof(float) * (size_t)G * L * PP));
    const int64_t n_batches = (n_perms + PB - 1) / PB;
    const bool pipelined = order_mode == LRPERM_ORDER_FAST && h->overlap && n_batches > 0;
    const bool filt_serial = order_mode == LRPERM_ORDER_FAST && !pipelined && fast_filter_on(h, scores, cube_out);
    if (order_mode == LRPERM_ORDER_FAST && !pipelined) {
        if ((rc = wait_matrix(h, false))) return rc;
        if ((rc = ensure_fast_tables(h, fast_chunk, fast_tile))) return rc;
        if (filt_serial && (rc = ensure_filtered_chunks(h))) return rc;
        fast_run_stats(h, filt_serial);
        if ((rc = dbg_set_dims(h, std::max(h->n_slots, 1), PB))) return rc;
        CK(h->slab.ensure((size_t)N * PB));
        CK(h->partials.ensure(sizeof(float) * (size_t)std::max(h->n_slots, 1) * L * PB));
    }
    if (scores & LRPERM_SCORE_CPDB) CK(cudaMemsetAsync(h->counts_c.p, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(T, 1), h->stream));
    if (scores & LRPERM_SCORE_GMEAN) CK(cudaMemsetAsync(h->counts_g.p, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(T, 1), h->stream));

    const uint8_t* slab_labels = h->labels8.as<uint8_t>();
    const int32_t label_scale = (order_mode == LRPERM_ORDER_FAST && h->fast_variant >= 2) ? Q / 2 : 1;
    if (order_mode == LRPERM_ORDER_FAST && h->fast_variant >= 2) {
        const int32_t scale = Q / 2;
        if (h->labels8s_scale != scale) {
            CK(h->labels8s.ensure((size_t)std::max<int64_t>(N, 1)));
            if (N > 0) {
                scale_labels_kernel<<<grid_for(N, 256, 1 << 30), 256, 0, h->stream>>>(h->labels8.as<uint8_t>(), (int32_t)N, scale, h->labels8s.as<uint8_t>());
                CK_LAUNCH();
            }
            h->labels8s_scale = scale;
        }
        slab_labels = h->labels8s.as<uint8_t>();
    }
    if (pipelined) {
        RunArgs ra{n_perms, perm_offset, seed, perm_source, perm_idx, perm_idx_is_64, perm_idx_on_device, scores, gmean_mode,
                   cube_out, cube_on_device};
        if ((rc = run_fast_pipelined(h, ra, Q, PB, slab_labels, fast_chunk, fast_tile))) { cudaStreamSynchronize(h->aux_stream); cudaStreamSynchronize(h->stream); return rc; }
        if (scores & LRPERM_SCORE_CPDB) if ((rc = download(h, counts_cpdb, h->counts_c.p, sizeof(uint32_t) * (size_t)T, counts_on_device))) return rc;
        if (scores & LRPERM_SCORE_GMEAN) if ((rc = download(h, counts_gmean, h->counts_g.p, sizeof(uint32_t) * (size_t)T, counts_on_device))) return rc;
        CK(cudaStreamSynchronize(h->stream));
        if (h->check_matrix_flag) {
            h->check_matrix_flag = false;
            int flag = 0;
            CK(cudaMemcpy(&flag, h->err_flag_m.p, sizeof(int), cudaMemcpyDeviceToHost));
            if (flag) {
                CK(cudaMemset(h->err_flag_m.p, 0, sizeof(int)));
                h->have_matrix = false;
                return h->fail(LRPERM_ERR_INVALID, "lrperm_set_matrix_csr (deferred check): invalid input (flag=%d: 1=column out of range)", flag);
            }
        }
        return collect_pipelined_timings(h, h->n_sched);
    }
    const size_t n_events = (size_t)n_batches * 5 + 2;
    while (h->events.size() < n_events) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        h->events.push_back(ev);
    }
    size_t evi = 0;
    CK(cudaEventRecord(h->events[evi++], h->stream));   // run start

    const size_t idx_elem = perm_idx_is_64 ? 8 : 4;
    for (int64_t b = 0; b < n_batches; ++b) {
        const int64_t p0 = b * PB;
        const int32_t pb = (int32_t)std::min<int64_t>(PB, n_perms - p0);
        CK(cudaEventRecord(h->events[evi++], h->stream));
        // ---- stage 1: permutation source for this batch ----
        PermSource ps{};
        ps.is64 = perm_idx_is_64;
        ps.seed = seed;
        ps.perm_id0 = perm_offset + p0;
        ps.philox = perm_source == LRPERM_PERMS_PHILOX;
        if (!ps.philox) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(perm_idx) + idx_elem * (size_t)p0 * (size_t)N;
            if (perm_idx_on_device) ps.idx = src;
            else {
                if ((rc = upload(h, h->perm_dev, src, idx_elem * (size_t)pb * (size_t)N, 0))) return rc;
                ps.idx = h->perm_dev.p;
            }
        }
        if (order_mode == LRPERM_ORDER_FAST) {
            if (pb < PB) CK(cudaMemsetAsync(h->slab.p, 0, (size_t)N * PB, h->stream));
            if (ps.philox) {
                const int cells_per_block = 128;
                dim3 grd((unsigned)((N + cells_per_block - 1) / cells_per_block), (unsigned)((pb + 127) / 128));
                const size_t tab_smem = sizeof(int32_t) * (size_t)(L + 1) + (size_t)h->n_buckets;
                slab_philox_kernel<<<grd, 128, tab_smem, h->stream>>>(ps, h->cl_start.as<int32_t>(), h->bucket_cluster.as<uint8_t>(), h->n_buckets,
                                                                      h->bucket_shift, L, label_scale, (int32_t)N, pb, h->slab.as<uint8_t>(), PB, cells_per_block);
            } else {
                dim3 grd(grid_for(N, 256, 1024), (unsigned)pb);
                slab_from_index_kernel<<<grd, 256, 0, h->stream>>>(ps, slab_labels, (int32_t)N, pb, h->slab.as<uint8_t>(), PB);
            }
            CK_LAUNCH();
        }
        CK(cudaEventRecord(h->events[evi++], h->stream));
        // ---- stage 2: per-(gene, cluster, permutation) means ----
        if (order_mode == LRPERM_ORDER_FAST) {
            rc = launch_fast_any(h, Q, PB, h->slab.as<uint8_t>(), h->partials.as<float>(), filt_serial);
            if (rc) return rc;
            CK(cudaEventRecord(h->events[evi++], h->stream));   // main kernel done
            fast_finalize_kernel<<<grid_for((int64_t)G * L * PB, 256, h->sm_count * 8), 256, 0, h->stream>>>(
                h->partials.as<float>(), h->tile_slot_ptr.as<int32_t>(), h->n_tiles, h->inv_n.as<float>(), G, L, PB, h->cube.as<float>(), PP, 0,
                filt_serial ? h->needed_dev.as<uint8_t>() : nullptr);
            CK_LAUNCH();
        } else {
            if ((rc = launch_exact(h, ps, pb, h->cube.as<float>(), PP))) return rc;
            CK(cudaEventRecord(h->events[evi++], h->stream));   // main kernel done
        }
        CK(cudaEventRecord(h->events[evi++], h->stream));
        // ---- stage 3: score + compare + count ----
        if (scores) {
            if (gmean_mode == LRPERM_GMEAN_LITERAL) rc = launch_score_t<true>(h, scores, h->cube.as<float>(), PP, pb);
            else rc = launch_score_t<false>(h, scores, h->cube.as<float>(), PP, pb);
            if (rc) return rc;
        }
        if (cube_out) {
            float* dst = cube_out + (size_t)p0 * L * G;
            float* dev_dst = dst;
            if (!cube_on_device) { CK(h->stage_out.ensure(sizeof(float) * (size_t)pb * L * G)); dev_dst = h->stage_out.as<float>(); }
            dim3 blk(32, 8), grd((G + 31) / 32, (pb + 31) / 32, L);
            cube_export_kernel<<<grd, blk, 0, h->stream>>>(h->cube.as<float>(), G, L, PP, pb, dev_dst);
            CK_LAUNCH();
            if (!cube_on_device) {
                CK(cudaMemcpyAsync(dst, dev_dst, sizeof(float) * (size_t)pb * L * G, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));   // staging buffer is reused by the next batch
            }
        }
        CK(cudaEventRecord(h->events[evi++], h->stream));
    }
    CK(cudaEventRecord(h->events[evi++], h->stream));   // run end (before result copies)
    if (scores & LRPERM_SCORE_CPDB) if ((rc = download(h, counts_cpdb, h->counts_c.p, sizeof(uint32_t) * (size_t)T, counts_on_device))) return rc;
    if (scores & LRPERM_SCORE_GMEAN) if ((rc = download(h, counts_gmean, h->counts_g.p, sizeof(uint32_t) * (size_t)T, counts_on_device))) return rc;
    CK(cudaStreamSynchronize(h->stream));

    h->t_setup = h->t_means = h->t_score = h->t_main = 0.f;
    h->n_main_launches = (int32_t)n_batches;
    for (int64_t b = 0; b < n_batches; ++b) {
        float ms;
        const size_t base = 1 + (size_t)b * 5;
        CK(cudaEventElapsedTime(&ms, h->events[base], h->events[base + 1])); h->t_setup += ms;
        CK(cudaEventElapsedTime(&ms, h->events[base + 1], h->events[base + 2])); h->t_main += ms;
        CK(cudaEventElapsedTime(&ms, h->events[base + 1], h->events[base + 3])); h->t_means += ms;
        CK(cudaEventElapsedTime(&ms, h->events[base + 3], h->events[base + 4])); h->t_score += ms;
    }
    CK(cudaEventElapsedTime(&h->t_total, h->events[0], h->events[evi - 1]));
    return LRPERM_OK;
}

int lrperm_run_trimean(lrperm_engine* h, int64_t n_perms, int64_t perm_offset, uint64_t seed, int perm_source,
                       const void* perm_idx, int perm_idx_is_64, int perm_idx_on_device, float recip32, double kh,
                       const double* obs_prob, int obs_on_device, uint32_t* counts, int counts_on_device,
                       double* cube_out, int cube_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("run_trimean");
    DeviceGuard guard(h->device);
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: set the matrix and labels first");
    { const int rcw = wait_matrix(h); if (rcw) return rcw; }
    if (n_perms < 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: n_perms < 0");
    if (perm_source != LRPERM_PERMS_INDEX && perm_source != LRPERM_PERMS_PHILOX) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: unknown perm_source %d", perm_source);
    if (perm_source == LRPERM_PERMS_INDEX && n_perms > 0 && !perm_idx) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: perm_idx is NULL");
    if (!(recip32 > 0.0f) || !std::isfinite(recip32)) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: recip32 must be positive and finite");
    const bool score = counts != nullptr;
    if (score && (!h->have_triples || !obs_prob)) return h->fail(LRPERM_ERR_INVALID, "lrperm_run_trimean: counts requested without triples / obs_prob");
    const int64_t T = h->T;
    h->launches = 0;
    int rc;
    if (score) {
        if ((rc = upload(h, h->thr_d, obs_prob, sizeof(double) * (size_t)T, obs_on_device))) return rc;
        CK(h->counts_t.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(T, 1)));
        CK(cudaMemsetAsync(h->counts_t.p, 0, sizeof(uint32_t) * (size_t)std::max<int64_t>(T, 1), h->stream));
    }
    TrimeanArgs a{n_perms, perm_offset, seed, perm_source, perm_idx, perm_idx_is_64, perm_idx_on_device, recip32, (double)recip32, kh,
                  score, cube_out, cube_on_device};
    if ((rc = run_trimean_batches(h, a))) { cudaStreamSynchronize(h->stream); return rc; }
    if (score && (rc = download(h, counts, h->counts_t.p, sizeof(uint32_t) * (size_t)T, counts_on_device))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return collect_trimean_timings(h);
}

int lrperm_observed_trimean(lrperm_engine* h, double recip, double* trimean, int out_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_observed_trimean: set the matrix and labels first");
    { const int rcw = wait_matrix(h); if (rcw) return rcw; }
    if (!(recip > 0.0) || !std::isfinite(recip) || !trimean) return h->fail(LRPERM_ERR_INVALID, "lrperm_observed_trimean: recip must be positive and finite, trimean non-NULL");
    const int32_t G = (int32_t)h->G, L = h->L;
    h->launches = 0;
    int rc;
    // one identity "permutation" through the same kernels; the batch is exported as [1][L][G]
    double* dev_dst = trimean;
    if (!out_on_device) { CK(h->tmp_d.ensure(sizeof(double) * (size_t)L * G)); dev_dst = h->tmp_d.as<double>(); }
    TrimeanArgs a{1, 0, 0, -1, nullptr, 0, 0, (float)recip, recip, 0.0, false, dev_dst, 1};
    if ((rc = run_trimean_batches(h, a))) { cudaStreamSynchronize(h->stream); return rc; }
    if (!out_on_device) CK(cudaMemcpyAsync(trimean, dev_dst, sizeof(double) * (size_t)L * G, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return collect_trimean_timings(h);
}

extern "C++" {
namespace {
// byte layout of the resolved rows buffer: 5 int32 columns, 2 statistic columns (ST), 2 double columns, 1 byte column
struct RowsLayout { size_t inter, src, tgt, lg, rg, ls, rs, lp, rp, keep, total; };
RowsLayout rows_layout(size_t T, size_t st_size) {
    RowsLayout l;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    l.lp = take(8 * T); l.rp = take(8 * T); l.ls = take(st_size * T); l.rs = take(st_size * T);
    l.inter = take(4 * T); l.src = take(4 * T); l.tgt = take(4 * T); l.lg = take(4 * T); l.rg = take(4 * T); l.keep = take(T);
    l.total = o;
    return l;
}

template <typename ST>
int resolve_impl(lrperm_engine* h, int32_t n, int32_t max_l, int32_t max_r, const int32_t* d_lsub, const int32_t* d_rsub,
                 const ST* d_stat, const double* d_props, double expr_prop, const uint8_t* d_mask, int first_subunit, int return_all,
                 int64_t* n_rows) {
    const int32_t L = h->L, G = (int32_t)h->G;
    const size_t nl = (size_t)n * L;
    // side tables: per side gene[nl] int32, stat[nl] ST, prop[nl] f64, pmin[nl] f64
    const size_t side_bytes = ((nl * 8 + 255) & ~(size_t)255) * 3 + ((nl * 4 + 255) & ~(size_t)255);
    CK(h->rs_side.ensure(2 * side_bytes));
    auto side = [&](int which, int32_t*& gene, ST*& stv, double*& prop, double*& pmin) {
        unsigned char* b = h->rs_side.as<unsigned char>() + (size_t)which * side_bytes;
        const size_t a8 = (nl * 8 + 255) & ~(size_t)255;
        prop = reinterpret_cast<double*>(b); pmin = reinterpret_cast<double*>(b + a8);
        stv = reinterpret_cast<ST*>(b + 2 * a8); gene = reinterpret_cast<int32_t*>(b + 3 * a8);
    };
    int32_t *gl, *gr; ST *sl, *sr; double *pl, *pr, *ml, *mr;
    side(0, gl, sl, pl, ml); side(1, gr, sr, pr, mr);
    const int g1 = grid_for((int64_t)nl, 256, 1 << 30);
    resolve_side_kernel<ST><<<g1, 256, 0, h->stream>>>(d_lsub, n, max_l, L, G, d_stat, d_props, first_subunit, gl, sl, pl, ml);
    CK_LAUNCH();
    resolve_side_kernel<ST><<<g1, 256, 0, h->stream>>>(d_rsub, n, max_r, L, G, d_stat, d_props, first_subunit, gr, sr, pr, mr);
    CK_LAUNCH();
    const int64_t F = (int64_t)L * L * n;
    if (F >= 0x7fffffffLL) return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_resolve_triples: L^2 x n_inter must be < 2^31");
    CK(h->rs_flag.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(F, 1)));
    CK(h->rs_pos.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(F, 1)));
    int64_t T = 0;
    if (F > 0) {
        const int g2 = grid_for(F, 256, 1 << 30);
        resolve_flags_kernel<<<g2, 256, 0, h->stream>>>(ml, mr, d_mask, n, L, expr_prop, return_all, h->rs_flag.as<int32_t>());
        CK_LAUNCH();
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->rs_flag.as<int32_t>(), h->rs_pos.as<int32_t>(), (int)F, h->stream));
        CK(h->sort_tmp.ensure(tmp_bytes));
        tmp_bytes = h->sort_tmp.cap;
        CK(cub::DeviceScan::ExclusiveSum(h->sort_tmp.p, tmp_bytes, h->rs_flag.as<int32_t>(), h->rs_pos.as<int32_t>(), (int)F, h->stream));
        int32_t last_pos = 0, last_flag = 0;
        CK(cudaMemcpyAsync(&last_pos, h->rs_pos.as<int32_t>() + (F - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&last_flag, h->rs_flag.as<int32_t>() + (F - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        T = (int64_t)last_pos + last_flag;
        const RowsLayout lay = rows_layout((size_t)T, sizeof(ST));
        CK(h->rs_rows.ensure(std::max<size_t>(lay.total, 256)));
        unsigned char* b = h->rs_rows.as<unsigned char>();
        if (T > 0) {
            resolve_scatter_kernel<ST><<<g2, 256, 0, h->stream>>>(
                h->rs_flag.as<int32_t>(), h->rs_pos.as<int32_t>(), n, L, expr_prop, gl, sl, pl, ml, gr, sr, pr, mr,
                reinterpret_cast<int32_t*>(b + lay.inter), reinterpret_cast<int32_t*>(b + lay.src), reinterpret_cast<int32_t*>(b + lay.tgt),
                reinterpret_cast<int32_t*>(b + lay.lg), reinterpret_cast<int32_t*>(b + lay.rg), reinterpret_cast<ST*>(b + lay.ls),
                reinterpret_cast<ST*>(b + lay.rs), reinterpret_cast<double*>(b + lay.lp), reinterpret_cast<double*>(b + lay.rp), b + lay.keep);
            CK_LAUNCH();
        }
    }
    h->rs_T = T;
    h->rs_f64 = sizeof(ST) == 8;
    if (n_rows) *n_rows = T;
    return LRPERM_OK;
}
}  // namespace
}  // extern "C++"

int lrperm_resolve_triples(lrperm_engine* h, int32_t n_inter, int32_t max_l, int32_t max_r, const int32_t* lig_sub, const int32_t* rec_sub,
                           const void* stat, int stat_is_f64, const double* props, double expr_prop, const uint8_t* pair_mask,
                           int first_subunit, int return_all, int64_t* n_rows) {
    if (!h) return LRPERM_ERR_INVALID;
    StageTimer timer("resolve_triples");
    DeviceGuard guard(h->device);
    h->rs_T = -1;
    if (!h->have_matrix || !h->have_labels) return h->fail(LRPERM_ERR_INVALID, "lrperm_resolve_triples: set the matrix and labels first");
    if (n_inter < 0 || max_l <= 0 || max_r <= 0 || (n_inter > 0 && (!lig_sub || !rec_sub)) || !stat || !props)
        return h->fail(LRPERM_ERR_INVALID, "lrperm_resolve_triples: bad arguments");
    const int32_t L = h->L, G = (int32_t)h->G;
    const size_t LG = (size_t)L * G, st_size = stat_is_f64 ? 8 : 4;
    // inputs (host) -> one device buffer: [props f64][stat][lig_sub][rec_sub][pair_mask]
    const size_t o_props = 0, o_stat = LG * 8, o_ls = o_stat + ((LG * st_size + 7) & ~(size_t)7);
    const size_t o_rs = o_ls + (((size_t)n_inter * max_l * 4 + 7) & ~(size_t)7), o_mask = o_rs + (((size_t)n_inter * max_r * 4 + 7) & ~(size_t)7);
    CK(h->rs_in.ensure(o_mask + (size_t)L * L + 16));
    unsigned char* b = h->rs_in.as<unsigned char>();
    CK(cudaMemcpyAsync(b + o_props, props, LG * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(b + o_stat, stat, LG * st_size, cudaMemcpyHostToDevice, h->stream));
    if (n_inter > 0) {
        CK(cudaMemcpyAsync(b + o_ls, lig_sub, (size_t)n_inter * max_l * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(b + o_rs, rec_sub, (size_t)n_inter * max_r * 4, cudaMemcpyHostToDevice, h->stream));
    }
    if (pair_mask) CK(cudaMemcpyAsync(b + o_mask, pair_mask, (size_t)L * L, cudaMemcpyHostToDevice, h->stream));
    const uint8_t* d_mask = pair_mask ? b + o_mask : nullptr;
    if (stat_is_f64)
        return resolve_impl<double>(h, n_inter, max_l, max_r, reinterpret_cast<int32_t*>(b + o_ls), reinterpret_cast<int32_t*>(b + o_rs),
                                    reinterpret_cast<double*>(b + o_stat), reinterpret_cast<double*>(b + o_props), expr_prop, d_mask,
                                    first_subunit, return_all, n_rows);
    return resolve_impl<float>(h, n_inter, max_l, max_r, reinterpret_cast<int32_t*>(b + o_ls), reinterpret_cast<int32_t*>(b + o_rs),
                               reinterpret_cast<float*>(b + o_stat), reinterpret_cast<double*>(b + o_props), expr_prop, d_mask,
                               first_subunit, return_all, n_rows);
}

int lrperm_fetch_triples(lrperm_engine* h, int32_t* inter, int32_t* src, int32_t* tgt, int32_t* lig_gene, int32_t* rec_gene,
                         void* lig_stat, void* rec_stat, double* lig_props, double* rec_props, uint8_t* keep) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (h->rs_T < 0) return h->fail(LRPERM_ERR_INVALID, "lrperm_fetch_triples: call lrperm_resolve_triples first");
    const size_t T = (size_t)h->rs_T, st_size = h->rs_f64 ? 8 : 4;
    if (T == 0) return LRPERM_OK;
    const RowsLayout lay = rows_layout(T, st_size);
    const unsigned char* b = h->rs_rows.as<unsigned char>();
    int rc;
    if ((rc = download(h, inter, b + lay.inter, 4 * T, 0))) return rc;
    if ((rc = download(h, src, b + lay.src, 4 * T, 0))) return rc;
    if ((rc = download(h, tgt, b + lay.tgt, 4 * T, 0))) return rc;
    if ((rc = download(h, lig_gene, b + lay.lg, 4 * T, 0))) return rc;
    if ((rc = download(h, rec_gene, b + lay.rg, 4 * T, 0))) return rc;
    if ((rc = download(h, lig_stat, b + lay.ls, st_size * T, 0))) return rc;
    if ((rc = download(h, rec_stat, b + lay.rs, st_size * T, 0))) return rc;
    if ((rc = download(h, lig_props, b + lay.lp, 8 * T, 0))) return rc;
    if ((rc = download(h, rec_props, b + lay.rp, 8 * T, 0))) return rc;
    if ((rc = download(h, keep, b + lay.keep, T, 0))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_rank_average(lrperm_engine* h, int64_t n, const double* values, int negate, double* ranks, int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (n < 0 || n >= 0x7fffffffLL || (n > 0 && (!values || !ranks))) return h->fail(LRPERM_ERR_INVALID, "lrperm_rank_average: bad arguments");
    if (n == 0) return LRPERM_OK;
    int rc;
    const size_t N = (size_t)n;
    // tmp_a: values (host input) | ranks ; tmp_b: keys, tmp_c: sorted keys ; tmp_d: idx, sorted idx, flags, gid, starts
    const double* d_vals = values;
    if (!on_device) { if ((rc = upload(h, h->tmp_a, values, sizeof(double) * N, 0))) return rc; d_vals = h->tmp_a.as<double>(); }
    CK(h->tmp_b.ensure(sizeof(unsigned long long) * N));
    CK(h->tmp_c.ensure(sizeof(unsigned long long) * N));
    CK(h->tmp_d.ensure(sizeof(int32_t) * N * 5));
    CK(h->stage_out.ensure(sizeof(double) * N));
    uint32_t* idx = h->tmp_d.as<uint32_t>();
    uint32_t* idx_sorted = idx + N;
    int32_t* flag = reinterpret_cast<int32_t*>(idx + 2 * N);
    int32_t* gid = flag + N;
    int32_t* starts = gid + N;
    const int grid = grid_for(n, 256, 1 << 30);
    rank_keys_kernel<<<grid, 256, 0, h->stream>>>(d_vals, n, negate, h->tmp_b.as<unsigned long long>(), idx);
    CK_LAUNCH();
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->tmp_b.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(), idx, idx_sorted, (int)n, 0, 64, h->stream));
    size_t scan_bytes = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flag, gid, (int)n, h->stream));
    CK(h->sort_tmp.ensure(std::max(tmp_bytes, scan_bytes)));
    tmp_bytes = h->sort_tmp.cap;
    CK(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, h->tmp_b.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(), idx, idx_sorted, (int)n, 0, 64, h->stream));
    rank_flags_kernel<<<grid, 256, 0, h->stream>>>(h->tmp_c.as<unsigned long long>(), n, flag);
    CK_LAUNCH();
    scan_bytes = h->sort_tmp.cap;
    CK(cub::DeviceScan::InclusiveSum(h->sort_tmp.p, scan_bytes, flag, gid, (int)n, h->stream));
    rank_starts_kernel<<<grid, 256, 0, h->stream>>>(flag, gid, n, starts);
    CK_LAUNCH();
    int32_t n_groups = 0;
    CK(cudaMemcpyAsync(&n_groups, gid + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double* d_ranks = on_device ? ranks : h->stage_out.as<double>();
    rank_scatter_kernel<<<grid, 256, 0, h->stream>>>(gid, starts, idx_sorted, n, n_groups, d_ranks);
    CK_LAUNCH();
    if (!on_device) CK(cudaMemcpyAsync(ranks, d_ranks, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_argsort(lrperm_engine* h, int64_t n, const double* values, int descending, int32_t* order, int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (n < 0 || n >= 0x7fffffffLL || (n > 0 && (!values || !order))) return h->fail(LRPERM_ERR_INVALID, "lrperm_argsort: bad arguments");
    if (n == 0) return LRPERM_OK;
    int rc;
    const size_t N = (size_t)n;
    const double* d_vals = values;
    if (!on_device) { if ((rc = upload(h, h->tmp_a, values, sizeof(double) * N, 0))) return rc; d_vals = h->tmp_a.as<double>(); }
    CK(h->tmp_b.ensure(sizeof(unsigned long long) * N));
    CK(h->tmp_c.ensure(sizeof(unsigned long long) * N));
    CK(h->tmp_d.ensure(sizeof(int32_t) * N * 2));
    uint32_t* idx = h->tmp_d.as<uint32_t>();
    uint32_t* idx_sorted = on_device ? reinterpret_cast<uint32_t*>(order) : idx + N;
    // stable ascending sort of the key of -v puts the largest v first and keeps equal values in their original order
    rank_keys_kernel<<<grid_for(n, 256, 1 << 30), 256, 0, h->stream>>>(d_vals, n, descending, h->tmp_b.as<unsigned long long>(), idx);
    CK_LAUNCH();
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->tmp_b.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(), idx, idx_sorted, (int)n, 0, 64, h->stream));
    CK(h->sort_tmp.ensure(tmp_bytes));
    tmp_bytes = h->sort_tmp.cap;
    CK(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp_bytes, h->tmp_b.as<unsigned long long>(), h->tmp_c.as<unsigned long long>(), idx, idx_sorted, (int)n, 0, 64, h->stream));
    if (!on_device) CK(cudaMemcpyAsync(order, idx_sorted, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_rra(lrperm_engine* h, int64_t n_rows, int32_t n_cols, const double* rmat, const double* col_max, double* rho, int on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (n_rows < 0 || n_cols <= 0 || n_cols > kRraMaxCols || (n_rows > 0 && (!rmat || !col_max || !rho)))
        return h->fail(LRPERM_ERR_INVALID, "lrperm_rra: bad arguments (at most %d score columns)", kRraMaxCols);
    if (n_rows == 0) return LRPERM_OK;
    int rc;
    const size_t cells = (size_t)n_rows * (size_t)n_cols;
    const double *d_r = rmat, *d_m = col_max;
    if (!on_device) {
        if ((rc = upload(h, h->tmp_a, rmat, sizeof(double) * cells, 0))) return rc;
        if ((rc = upload(h, h->small_a, col_max, sizeof(double) * (size_t)n_cols, 0))) return rc;
        d_r = h->tmp_a.as<double>(); d_m = h->small_a.as<double>();
        CK(h->stage_out.ensure(sizeof(double) * (size_t)n_rows));
    }
    double* d_out = on_device ? rho : h->stage_out.as<double>();
    rra_kernel<<<grid_for(n_rows, 128, 1 << 30), 128, 0, h->stream>>>(d_r, d_m, n_rows, n_cols, d_out);
    CK_LAUNCH();
    if (!on_device) CK(cudaMemcpyAsync(rho, d_out, sizeof(double) * (size_t)n_rows, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_debug_bounds(lrperm_engine* h, uint32_t* flags, int64_t* first2) {
    if (!h || !flags) return LRPERM_ERR_INVALID;
#ifdef LRPERM_BOUNDS_CHECK
    DeviceGuard guard(h->device);
    CK(cudaStreamSynchronize(h->stream));
    unsigned int f = 0;
    long long first[2] = {0, 0};
    CK(cudaMemcpyFromSymbol(&f, g_dbg_flags, sizeof(f)));
    CK(cudaMemcpyFromSymbol(first, g_dbg_first, sizeof(first)));
    const unsigned int zero = 0;
    CK(cudaMemcpyToSymbol(g_dbg_flags, &zero, sizeof(zero)));
    *flags = f;
    if (first2) { first2[0] = first[0]; first2[1] = first[1]; }
    return LRPERM_OK;
#else
    (void)first2;
    return h->fail(LRPERM_ERR_UNSUPPORTED, "lrperm_debug_bounds: this library was built without -DLRPERM_BOUNDS_CHECK (python -m liana_py_b200.build --debug)");
#endif
}

int lrperm_run_stats(lrperm_engine* h, int64_t* out4) {
    if (!h || !out4) return LRPERM_ERR_INVALID;
    out4[0] = h->stat_genes_scored; out4[1] = h->stat_nnz_scored; out4[2] = h->stat_chunks_run; out4[3] = h->nnz;
    return LRPERM_OK;
}

int lrperm_matrix_info(lrperm_engine* h, int64_t* out3) {
    if (!h || !out3) return LRPERM_ERR_INVALID;
    if (!h->have_matrix) return h->fail(LRPERM_ERR_INVALID, "lrperm_matrix_info: no matrix set");
    out3[0] = h->N; out3[1] = h->G; out3[2] = h->nnz;
    return LRPERM_OK;
}

int lrperm_export_matrix_csr(lrperm_engine* h, int32_t* indptr, int32_t* indices, float* data, int out_on_device) {
    if (!h) return LRPERM_ERR_INVALID;
    DeviceGuard guard(h->device);
    if (!h->have_matrix) return h->fail(LRPERM_ERR_INVALID, "lrperm_export_matrix_csr: no matrix set");
    if (!indptr || (h->nnz > 0 && (!indices || !data))) return h->fail(LRPERM_ERR_INVALID, "lrperm_export_matrix_csr: NULL output");
    int rc;
    if ((rc = wait_matrix(h, false))) return rc;
    // cudaMemcpyDefault: with unified addressing the destination may be a buffer of ANOTHER device (peer copy over NVLink)
    const cudaMemcpyKind kind = out_on_device ? cudaMemcpyDefault : cudaMemcpyDeviceToHost;
    CK(cudaMemcpyAsync(indptr, h->indptr.p, sizeof(int32_t) * (size_t)(h->N + 1), kind, h->stream));
    if (h->nnz > 0) {
        const size_t n = (size_t)h->nnz;
        const int32_t* src_i; const float* src_v;
        if (!h->csr_packed) { src_i = h->st_indices.as<int32_t>(); src_v = h->st_data.as<float>(); }     // raw arrays of the upload
        else {
            CK(h->tmp_c.ensure(sizeof(int32_t) * n));
            CK(h->tmp_d.ensure(sizeof(float) * n));
            unpack_pairs_kernel<<<grid_for((int64_t)n, 256, h->sm_count * 16), 256, 0, h->stream>>>(h->csr_pairs.as<int2>(), (int64_t)n, h->tmp_c.as<int32_t>(), h->tmp_d.as<float>());
            CK_LAUNCH();
            src_i = h->tmp_c.as<int32_t>(); src_v = h->tmp_d.as<float>();
        }
        CK(cudaMemcpyAsync(indices, src_i, sizeof(int32_t) * n, kind, h->stream));
        CK(cudaMemcpyAsync(data, src_v, sizeof(float) * n, kind, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return LRPERM_OK;
}

int lrperm_trimean_timings(lrperm_engine* h, float* out6) {
    if (!h || !out6) return LRPERM_ERR_INVALID;
    for (int i = 0; i < 4; ++i) out6[i] = h->tm_ms[i];
    out6[4] = (float)h->tm_walked_warps;
    out6[5] = (float)h->tm_select_warps;
    return LRPERM_OK;
}

}  // extern "C"
