// host microbenchmark: copy+narrow (today's staging) vs read-only filter pass; build: g++ -O3 -march=native -pthread
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
using clk = std::chrono::steady_clock;
int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const int64_t N = argc > 2 ? atoll(argv[2]) : 80000;
    const int G = 30000, per_row = 2760;
    const int64_t nnz = N * per_row;
    std::vector<int64_t> indptr(N + 1);
    std::vector<int32_t> idx(nnz);
    std::vector<float> val(nnz);
    std::vector<int32_t> sel(G, -1);
    { uint64_t s = 12345; int k = 0; for (int g = 0; g < G; ++g) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; if ((s >> 33) % 16 == 0) sel[g] = k++; } printf("selected genes %d\n", k); }
    {   // sorted columns per row: stride walk with jitter
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
            uint64_t s = 777 + t;
            for (int64_t r = N * t / T; r < N * (t + 1) / T; ++r) {
                indptr[r] = r * per_row;
                double pos = 0, step = (double)G / per_row;
                for (int j = 0; j < per_row; ++j) { s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                    idx[r * per_row + j] = std::min<int>(G - 1, (int)(pos + (double)((s >> 40) & 1023) / 1024.0 * step * 0.9)); pos += step;
                    val[r * per_row + j] = 0.5f + (float)((s >> 20) & 0xffff) / 65536.0f; }
            }
        });
        for (auto& x : th) x.join();
        indptr[N] = nnz;
    }
    std::vector<uint16_t> d_idx(nnz);
    std::vector<float> d_val(nnz);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = clk::now();
        {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
                const int64_t a = nnz * t / T, b = nnz * (t + 1) / T;
                memcpy(d_val.data() + a, val.data() + a, (b - a) * 4);
                uint32_t bad = 0;
                for (int64_t o = a; o < b; ++o) { const uint32_t c = (uint32_t)idx[o]; bad |= c; d_idx[o] = (uint16_t)c; }
                if (bad >> 16) printf("bad\n");
            });
            for (auto& x : th) x.join();
        }
        double dt = std::chrono::duration<double>(clk::now() - t0).count();
        printf("copy+narrow : %.1f ms  %.1f GB/s of input (8 B per entry)\n", dt * 1e3, nnz * 8.0 / dt / 1e9);
    }
    std::vector<int32_t> o_idx(nnz / 8 + 1024 * T);
    std::vector<float> o_val(nnz / 8 + 1024 * T);
    std::vector<int32_t> row_cnt(N);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = clk::now();
        std::atomic<int64_t> total_out{0};
        std::vector<std::vector<double>> sums(T, std::vector<double>(G, 0.0));
        std::vector<double> tot(T, 0.0);
        {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
                double* cs = sums[t].data();
                int64_t k = (nnz / 8 / T + 1024) * t, k0 = k;
                int64_t bad = 0, unord = 0, empty = 0; double total = 0;
                for (int64_t r = N * t / T; r < N * (t + 1) / T; ++r) {
                    const int64_t s = indptr[r], e = indptr[r + 1];
                    double row = 0; int prev = -1; const int64_t kr = k;
                    for (int64_t o = s; o < e; ++o) {
                        const int c = idx[o]; const float v = val[o];
                        unord += (c <= prev); prev = c;
                        if (!std::isfinite(v)) { ++bad; continue; }
                        cs[c] += (double)v; row += (double)v;
                        if (sel[c] >= 0) { o_idx[k] = c; o_val[k] = v; ++k; }
                    }
                    row_cnt[r] = (int32_t)(k - kr);
                    empty += (row == 0.0); total += row;
                }
                tot[t] = total + (double)bad + (double)unord + (double)empty; total_out += k - k0;
            });
            for (auto& x : th) x.join();
        }
        double dt = std::chrono::duration<double>(clk::now() - t0).count();
        printf("filter pass : %.1f ms  %.1f GB/s of input, %.1f %% of the entries kept (%g)\n", dt * 1e3, nnz * 8.0 / dt / 1e9, 100.0 * total_out / nnz, tot[0]);
    }
    return 0;
}
