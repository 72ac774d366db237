// Host side of the count upload: narrows one int32 count diagonal to the smallest of u8 / u16 / i32 that holds
// every value, writing straight into the pinned staging buffer.
//
// Why: the worker-level entry (hp_band_upload_counts) is bound by the host memory system and by PCIe, not by the
// GPU.  Hi-C counts beyond the first few diagonals are small (cfg2: 99 % of the band fits a byte), so reading the
// caller's int32 arrays once and writing 1 B per pixel cuts the host write traffic and the PCIe bytes 4x; the GPU
// widens them again at HBM speed (k_unpack_counts).  Values are preserved exactly -- a diagonal holding one value
// above the range simply takes the next wider format -- so nothing about the results changes.
//
// Compiled by the host compiler (g++) as a separate translation unit so that the AVX2 intrinsics and the
// per-function target attribute stay away from nvcc's front end.  Runtime dispatch: AVX2 when the CPU has it,
// plain C++ otherwise.
#include "hp_hostpack.h"

#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace hp {

namespace {

// ---- portable versions -----------------------------------------------------------------------
bool narrow8_scalar(const int32_t* s, size_t n, uint8_t* d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = (uint32_t)s[i];
        acc |= v;
        d[i] = (uint8_t)v;
    }
    return (acc & ~0xFFu) == 0;
}
bool narrow16_scalar(const int32_t* s, size_t n, uint16_t* d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = (uint32_t)s[i];
        acc |= v;
        d[i] = (uint16_t)v;
    }
    return (acc & ~0xFFFFu) == 0;
}

#if defined(__x86_64__)
// ---- AVX2: 32 counts per iteration -----------------------------------------------------------
// NT: the destination is 16-byte aligned -> non-temporal stores.  The staging buffer is written once and next read by the
// copy engine, never by this core: a regular store would first read the line it is about to overwrite (read for
// ownership), 1 B per pixel of host DRAM traffic on a path that is bound by exactly that (7 -> 6 B per pixel).
template <bool NT>
__attribute__((target("avx2"))) static inline void put256(void* p, __m256i v) {
    if (NT) {
        _mm_stream_si128((__m128i*)p, _mm256_castsi256_si128(v));
        _mm_stream_si128((__m128i*)p + 1, _mm256_extracti128_si256(v, 1));
    } else {
        _mm256_storeu_si256((__m256i*)p, v);
    }
}
template <bool NT>
__attribute__((target("avx2"))) bool narrow8_avx2(const int32_t* s, size_t n, uint8_t* d) {
    __m256i acc = _mm256_setzero_si256();
    const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);      // undo the per-lane interleave of the two packs
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(s + i + 8));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 16));
        const __m256i e = _mm256_loadu_si256((const __m256i*)(s + i + 24));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, e)));
        const __m256i ab = _mm256_packus_epi32(a, b);                    // values out of range saturate; acc notices
        const __m256i ce = _mm256_packus_epi32(c, e);
        const __m256i o = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, ce), fix);
        put256<NT>(d + i, o);
    }
    alignas(32) uint32_t t[8];
    _mm256_store_si256((__m256i*)t, acc);
    uint32_t r = t[0] | t[1] | t[2] | t[3] | t[4] | t[5] | t[6] | t[7];
    for (; i < n; ++i) {
        const uint32_t v = (uint32_t)s[i];
        r |= v;
        d[i] = (uint8_t)v;
    }
    return (r & ~0xFFu) == 0;
}
template <bool NT>
__attribute__((target("avx2"))) bool narrow16_avx2(const int32_t* s, size_t n, uint16_t* d) {
    __m256i acc = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(s + i + 8));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        const __m256i o = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
        put256<NT>(d + i, o);
    }
    alignas(32) uint32_t t[8];
    _mm256_store_si256((__m256i*)t, acc);
    uint32_t r = t[0] | t[1] | t[2] | t[3] | t[4] | t[5] | t[6] | t[7];
    for (; i < n; ++i) {
        const uint32_t v = (uint32_t)s[i];
        r |= v;
        d[i] = (uint16_t)v;
    }
    return (r & ~0xFFFFu) == 0;
}
bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#endif

constexpr size_t kBlock = 8192;      // range check granularity: a diagonal that does not fit is abandoned early

}  // namespace

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(unsigned char* d, const unsigned char* s, size_t n) {
    const size_t head = (32 - ((uintptr_t)d & 31)) & 31;               // up to the first 32-byte boundary of the destination
    if (head >= n) { memcpy(d, s, n); return; }
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a);
        _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c);
        _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    memcpy(d + i, s + i, n - i);
    _mm_sfence();
}
#endif
void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
    if (have_avx2() && bytes >= 4096) { stream_copy_avx2((unsigned char*)dst, (const unsigned char*)src, bytes); return; }
#endif
    memcpy(dst, src, bytes);
}

int narrow_diagonal(const int32_t* src, size_t len, void* dst) {
#if defined(__x86_64__)
    const bool simd = have_avx2();
#else
    const bool simd = false;
#endif
    bool ok = true;
#if defined(__x86_64__)
    const bool nt = simd && ((uintptr_t)dst & 15) == 0;      // (kBlock keeps every block's start 16-byte aligned)
    struct Fence { bool on; ~Fence() { if (on) _mm_sfence(); } } fence{nt};      // the copy engine reads what was streamed
#endif
    for (size_t i = 0; i < len && ok; i += kBlock) {
        const size_t m = len - i < kBlock ? len - i : kBlock;
#if defined(__x86_64__)
        ok = !simd ? narrow8_scalar(src + i, m, (uint8_t*)dst + i)
                   : nt ? narrow8_avx2<true>(src + i, m, (uint8_t*)dst + i) : narrow8_avx2<false>(src + i, m, (uint8_t*)dst + i);
#else
        ok = narrow8_scalar(src + i, m, (uint8_t*)dst + i);
#endif
    }
    if (ok) return 1;
    ok = true;
    for (size_t i = 0; i < len && ok; i += kBlock) {
        const size_t m = len - i < kBlock ? len - i : kBlock;
#if defined(__x86_64__)
        ok = !simd ? narrow16_scalar(src + i, m, (uint16_t*)dst + i)
                   : nt ? narrow16_avx2<true>(src + i, m, (uint16_t*)dst + i) : narrow16_avx2<false>(src + i, m, (uint16_t*)dst + i);
#else
        ok = narrow16_scalar(src + i, m, (uint16_t*)dst + i);
#endif
    }
    if (ok) return 2;
    stream_copy(dst, src, len * sizeof(int32_t));
    return 4;
}

}  // namespace hp
