// Host-side staging copy: glibc memcpy vs AVX2 streaming stores, T threads x 6.2 MB frames (CPU only; run on the GPU box to see ITS cores).
// build: g++ -O2 -mavx2 -pthread scripts/ubench/host_copy.cpp -o build/host_copy; usage: build/host_copy [threads]
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void nt_copy(char* d, const char* s, size_t n) {
    // assumes 32-byte aligned d
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    memcpy(d + i, s + i, n - i);
    _mm_sfence();
}
int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 8;
    size_t N = 6220800;
    int frames = 16;
    std::vector<char*> src(T * frames), dst(T * frames);
    for (auto& p : src) { p = (char*)aligned_alloc(4096, N + 4096); memset(p, 1, N); }
    for (auto& p : dst) { p = (char*)aligned_alloc(4096, N + 4096); memset(p, 2, N); }
    for (int mode = 0; mode < 2; ++mode) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
            for (int rep = 0; rep < 4; ++rep)
                for (int f = 0; f < frames; ++f) {
                    if (mode == 0) memcpy(dst[t * frames + f], src[t * frames + f], N);
                    else nt_copy(dst[t * frames + f], src[t * frames + f], N);
                }
        });
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%s T=%d: %.1f GB/s copied (%.2f ms per 6.2MB per thread)\n", mode ? "nt_copy" : "memcpy ", T, T * 4.0 * frames * N / dt / 1e9, dt / (4.0 * frames) * 1e3);
    }
}
