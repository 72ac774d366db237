#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
int main() {
    const size_t N = 160u << 20;
    char* src = (char*)malloc(N); memset(src, 1, N);
    char* dst; cudaHostAlloc((void**)&dst, N, cudaHostAllocDefault); memset(dst, 2, N);
    for (int nt : {1, 2, 4, 8, 16}) {
        auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 5; ++rep) {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back([&, t]() { memcpy(dst + N / nt * t, src + N / nt * t, N / nt); });
            for (auto& x : th) x.join();
        }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 5;
        printf("threads %2d: pageable->pinned memcpy %.1f GB/s\n", nt, N / s / 1e9);
    }
    return 0;
}
