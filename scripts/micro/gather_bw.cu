// Microbenchmark: what HBM bandwidth does the EXACT kernel's access pattern allow on B200?
// Every warp reads a sequence of random "rows" (contiguous runs of int2, lengths like the CSR rows of
// configs[2]: ~166 entries = 1.3 KB) from a 665 MB array and only sums them - no shared-memory
// accumulation.  Variants: rows per warp in flight (1 or 2), entry size 8 B (int2) or 6 B (u16 + f32 SoA).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu && ./gather_bw
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#include <cuda_runtime.h>

__device__ __forceinline__ int2 ld_stream_int2(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

template <int INFLIGHT>
__global__ void gather_rows(const int2* __restrict__ pairs, const int32_t* __restrict__ row_start, const int32_t* __restrict__ row_len,
                            const int32_t* __restrict__ seq, int rows_per_warp, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int32_t* my = seq + (int64_t)warp * rows_per_warp;
    float acc = 0.f;
    constexpr int U = 8;
    for (int r = 0; r < rows_per_warp; r += INFLIGHT) {
        int2 q[INFLIGHT][U];
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f) {
            const int row = my[r + f];
            const int s = row_start[row], e = s + row_len[row];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int o = s + lane + 32 * u;
                q[f][u] = (o < e) ? ld_stream_int2(pairs + o) : make_int2(0, 0);
            }
        }
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f)
#pragma unroll
            for (int u = 0; u < U; ++u) acc += __int_as_float(q[f][u].y) * (float)(q[f][u].x & 1);
    }
    if (acc == 123.456f) out[warp] = acc;
}

int main() {
    const int64_t N = 500000;
    std::mt19937 rng(1);
    std::vector<int32_t> start(N), len(N);
    int64_t nnz = 0;
    for (int64_t i = 0; i < N; ++i) { len[i] = 120 + rng() % 93; start[i] = (int32_t)nnz; nnz += len[i]; }
    int2* pairs; cudaMalloc(&pairs, nnz * 8); cudaMemset(pairs, 1, nnz * 8);
    int32_t *d_start, *d_len, *d_seq; float* d_out;
    cudaMalloc(&d_start, N * 4); cudaMalloc(&d_len, N * 4);
    cudaMemcpy(d_start, start.data(), N * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_len, len.data(), N * 4, cudaMemcpyHostToDevice);
    for (int warps_per_sm : {16, 24, 32, 48, 64}) {
        const int n_warps = 148 * warps_per_sm, rows_per_warp = 2048;
        std::vector<int32_t> seq((size_t)n_warps * rows_per_warp);
        for (auto& v : seq) v = rng() % N;
        cudaMalloc(&d_seq, seq.size() * 4); cudaMalloc(&d_out, n_warps * 4);
        cudaMemcpy(d_seq, seq.data(), seq.size() * 4, cudaMemcpyHostToDevice);
        double bytes = 0;
        for (auto v : seq) bytes += 8.0 * len[v];
        for (int inflight : {1, 2}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                if (inflight == 1) gather_rows<1><<<n_warps / 8, 256>>>(pairs, d_start, d_len, d_seq, rows_per_warp, d_out);
                else gather_rows<2><<<n_warps / 8, 256>>>(pairs, d_start, d_len, d_seq, rows_per_warp, d_out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("warps/SM %2d rows-in-flight %d: %.2f ms  %.0f GB/s (err %s)\n", warps_per_sm, inflight, best, bytes / best / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(d_seq); cudaFree(d_out);
    }
    return 0;
}
