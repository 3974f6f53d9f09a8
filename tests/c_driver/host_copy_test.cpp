// host_copy_test.cpp -- CPU check of the two pieces of the host pipeline that decide which bytes go where:
// copy_rows() (mosmicrosim_b200/csrc/host_copy.cpp: AVX2 streaming stores with unaligned heads and tails) against
// memcpy on ragged widths, pitches and alignments, and window_of_worker() (same file: the windows of a chunk's workers
// must tile the chunk exactly, in whole units except the last).
// Built and run by tests/test_host.py::test_host_copy_and_worker_windows (g++, no CUDA).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mms_api {
void copy_rows(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows);
void window_of_worker(int64_t c0, int64_t n_cells, int w, int workers, int64_t unit, int64_t &w0, int64_t &wn);
}

int main()
{
    uint64_t seed = 12345;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(seed >> 33); };
    int checked = 0;
    for (int trial = 0; trial < 4000; ++trial) {
        const size_t width = trial < 40 ? (size_t)trial * 37 : 1 + rnd() % 5000;
        const size_t rows = 1 + rnd() % 5;
        const bool packed = rnd() % 3 == 0;
        const size_t spitch = packed ? width : width + rnd() % 100, dpitch = packed ? width : width + rnd() % 100;
        const size_t so = rnd() % 64, dof = rnd() % 64;                  // misalign both ends
        std::vector<char> src(so + spitch * rows + 64), dst(dof + dpitch * rows + 64, (char)0x5a), ref;
        for (char &c : src) c = (char)rnd();
        ref = dst;
        for (size_t r = 0; r < rows; ++r) std::memcpy(ref.data() + dof + r * dpitch, src.data() + so + r * spitch, width);
        mms_api::copy_rows(dst.data() + dof, dpitch, src.data() + so, spitch, width, rows);
        if (dst != ref) { std::printf("copy_rows mismatch: width %zu rows %zu pitches %zu %zu offsets %zu %zu\n", width, rows, spitch, dpitch, so, dof); return 1; }
        ++checked;
    }
    for (int trial = 0; trial < 20000; ++trial) {
        const int64_t c0 = rnd() % 1000 * 32, n = 1 + rnd() % 200000, unit = (rnd() % 2) ? 32 : 1;
        const int workers = 1 + rnd() % 16;
        int64_t next = c0;
        for (int w = 0; w < workers; ++w) {
            int64_t w0, wn;
            mms_api::window_of_worker(c0, n, w, workers, unit, w0, wn);
            if (wn < 0 || (wn > 0 && w0 != next) || (wn > 0 && (w0 - c0) % unit != 0)) { std::printf("window mismatch\n"); return 1; }
            next += wn;
        }
        if (next != c0 + n) { std::printf("windows do not tile the chunk: %lld of %lld\n", (long long)(next - c0), (long long)n); return 1; }
    }
    std::printf("ok %d copies\n", checked);
    return 0;
}
