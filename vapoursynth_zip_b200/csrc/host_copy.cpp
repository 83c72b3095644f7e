// host_copy.cpp — the CPU side of the staging path: pageable VapourSynth planes <-> the slots' pinned buffers.
//
// Compiled by the host compiler alone (no CUDA in here).  With many getFrame threads copying at once the box is DRAM-bound, and a
// plain memcpy pays a read-for-ownership of every destination line on top of the read and the write.  Streaming (non-temporal)
// stores drop that third of the traffic: measured on the 16-core host of a B200 box with 6.2 MB frames (scripts/ubench/host_copy.cpp),
// 16 threads move 76 GB/s with streaming stores against 50 GB/s with memcpy, 8 threads tie at 51 GB/s and a single thread is
// 17 % slower (9.0 -> 7.5 GB/s) - so the runtime asks for streaming only while at least kStreamingBusySlots requests are in flight.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define VSZ_HAVE_X86 1
#endif

namespace vsz {

#ifdef VSZ_HAVE_X86
__attribute__((target("avx2"))) static void stream_copy_avx2(char* d, const char* s, size_t n) {
    // head: up to the first 32-byte boundary of the destination
    const size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head) {
        const size_t h = head < n ? head : n;
        memcpy(d, s, h);
        d += h; s += h; n -= h;
    }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
    }
    if (i < n) memcpy(d + i, s + i, n - i);
    _mm_sfence();  // the DMA engine (or the caller's consumer thread) must see the lines once this returns
}
static bool have_avx2() {
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
}
#endif

// rows of `row_bytes` from src (pitch spitch) to dst (pitch dpitch); streaming = bypass the cache on the store side
void copy_rows(char* dst, ptrdiff_t dpitch, const char* src, ptrdiff_t spitch, size_t row_bytes, int rows, bool streaming) {
    const bool one_block = dpitch == spitch && (size_t)spitch == row_bytes;
#ifdef VSZ_HAVE_X86
    if (streaming && have_avx2() && row_bytes >= 1024) {
        if (one_block) { stream_copy_avx2(dst, src, row_bytes * (size_t)rows); return; }
        for (int y = 0; y < rows; ++y) stream_copy_avx2(dst + dpitch * y, src + spitch * y, row_bytes);
        return;
    }
#endif
    (void)streaming;
    if (one_block) { memcpy(dst, src, row_bytes * (size_t)rows); return; }
    for (int y = 0; y < rows; ++y) memcpy(dst + dpitch * y, src + spitch * y, row_bytes);
}

}  // namespace vsz
