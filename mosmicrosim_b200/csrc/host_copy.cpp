// host_copy.cpp -- the row copies between caller (pageable) memory and the host pipeline's pinned bounce buffers
// (host_pipeline.cu), and the split of a chunk's columns among the workers: the CUDA-free pieces of the pipeline, so
// that tests/c_driver/host_copy_test.cpp can check them on a CPU.  Compiled by g++ alone (build.py) so that it can use the AVX2 streaming stores: a bounce copy
// is written once and next read by the DMA engine (upload) or never again by this thread (download), so the
// destination lines should not be read for ownership nor stay in the cache -- about twice the bandwidth of a
// cached memcpy per thread on the hosts measured.  Falls back to memcpy on CPUs without AVX2.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define MMS_X86 1
#else
#define MMS_X86 0
#endif

namespace mms_api {

#if MMS_X86
__attribute__((target("avx2"))) static void stream_copy_avx2(char *dst, const char *src, size_t n)
{
    // head: up to the first 32-byte boundary of the destination
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head) {
        const size_t h = head < n ? head : n;
        std::memcpy(dst, src, h);
        dst += h; src += h; n -= h;
    }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    for (; i + 32 <= n; i += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i),
                            _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)));
    if (i < n) std::memcpy(dst + i, src + i, n - i);
}

__attribute__((target("avx2"))) static void stream_fence() { _mm_sfence(); }

static bool have_avx2()
{
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
}
#endif

// which columns of a chunk a worker moves (host_pipeline.h)
void window_of_worker(int64_t c0, int64_t n_cells, int w, int workers, int64_t unit, int64_t &w0, int64_t &wn)
{
    const int64_t units = (n_cells + unit - 1) / unit;
    const int64_t per = (units + workers - 1) / workers;
    const int64_t a = (int64_t)w * per * unit < n_cells ? (int64_t)w * per * unit : n_cells;
    const int64_t b = (int64_t)(w + 1) * per * unit < n_cells ? (int64_t)(w + 1) * per * unit : n_cells;
    w0 = c0 + a;
    wn = b - a;
}

// rows x width bytes from src (pitch spitch) to dst (pitch dpitch)
void copy_rows(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows)
{
#if MMS_X86
    if (have_avx2() && width >= 256) {
        if (dpitch == width && spitch == width) {
            stream_copy_avx2(dst, src, width * rows);
        } else {
            for (size_t r = 0; r < rows; ++r) stream_copy_avx2(dst + r * dpitch, src + r * spitch, width);
        }
        stream_fence();
        return;
    }
#endif
    if (dpitch == width && spitch == width) { std::memcpy(dst, src, width * rows); return; }
    for (size_t r = 0; r < rows; ++r) std::memcpy(dst + r * dpitch, src + r * spitch, width);
}

}  // namespace mms_api
