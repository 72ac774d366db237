// Raw scaling of the upload narrowing (hp_hostpack.cpp) and of memcpy over host threads, no Python in the loop.
//   g++ -O2 -pthread -I hicpeaks_b200/csrc scratch/pack_bw.cpp hicpeaks_b200/csrc/hp_hostpack.cpp -o scratch/pack_bw
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "hp_hostpack.h"
int main() {
    const size_t n = 10u << 20;                         // 40 MB of int32 per thread
    const unsigned hw = std::thread::hardware_concurrency();
    printf("{\"host_threads\": %u, \"rows\": [", hw);
    bool first = true;
    for (unsigned T : {1u, 2u, 4u, 8u, 12u, 16u, 24u, 32u}) {
        if (T > hw) break;
        std::vector<std::vector<int32_t>> src(T);
        std::vector<std::vector<uint8_t>> dst(T);
        for (unsigned t = 0; t < T; ++t) {
            src[t].resize(n);
            dst[t].resize(n * 4);
            for (size_t i = 0; i < n; ++i) src[t][i] = (int32_t)((i * 2654435761u) >> 29);
        }
        for (int mode = 0; mode < 2; ++mode) {
            const int reps = 8;
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    for (int r = 0; r < reps; ++r) {
                        if (mode == 0) for (size_t o = 0; o < n; o += 20000) hp::narrow_diagonal(src[t].data() + o, std::min<size_t>(20000, n - o), dst[t].data() + o);
                        else memcpy(dst[t].data(), src[t].data(), n * 4);
                    }
                });
            for (auto& x : th) x.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%s{\"threads\": %u, \"mode\": \"%s\", \"int32_read_GBps\": %.1f}", first ? "" : ", ", T, mode == 0 ? "narrow" : "memcpy",
                   (double)T * reps * n * 4 / s / 1e9);
            first = false;
        }
    }
    printf("]}\n");
    return 0;
}
