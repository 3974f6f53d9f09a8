// chain_ws_tu.cu -- translation unit of the warp-pair form of the fused kernel (chain_ws.cuh; fast math only).
// Separate from chain_tu.cu so that the two long compilations run side by side (build.py).
#include "chain_ws.cuh"

namespace mms {

using namespace MMS_MODE;

namespace {

// the specialised instantiations of the pair kernel: two settings per thread, summaries only.  spec = density-
// dependence form | (shape of setting 0 + 1) << 1 | (shape of setting 1 + 1) << 3 (chain_ws.cuh); the shapes are
// compiled in only in the kChainMaxPairs instantiation and when every unit has the same pair of shapes.
template <int kNS, bool kFull, bool kHourly, int kMaxPairs>
auto pick_kernel_ws(const ChainParams &p) -> void (*)(ChainParams)
{
    constexpr bool kSpecialise = kNS == 2 && !kFull && !kHourly;
    constexpr bool kShapes = kSpecialise && kMaxPairs == kChainMaxPairs;
    if (!kSpecialise) return chain_kernel_ws<kNS, kFull, kHourly, kMaxPairs, -1>;
    const int dens = p.densmod_type == 1 ? 1 : 0;
    bool periodic = kShapes;
    for (int s = kNS; s < p.S && periodic; ++s) periodic = p.hab[s].shape == p.hab[s % kNS].shape;
    if (periodic) {
        switch (dens | (p.hab[0].shape + 1) << 1 | (p.hab[kNS > 1 ? 1 : 0].shape + 1) << 3) {
#define MMS_WS_CASE(v) case v: return chain_kernel_ws<kNS, kFull, kHourly, kMaxPairs, kShapes ? v : -1>;
            MMS_WS_CASE(10) MMS_WS_CASE(11) MMS_WS_CASE(12) MMS_WS_CASE(13)      // (cone | cylinder, cone)
            MMS_WS_CASE(18) MMS_WS_CASE(19) MMS_WS_CASE(20) MMS_WS_CASE(21)      // (cone | cylinder, cylinder)
#undef MMS_WS_CASE
        default: break;
        }
    }
    return dens ? chain_kernel_ws<kNS, kFull, kHourly, kMaxPairs, kSpecialise ? 1 : -1>
                : chain_kernel_ws<kNS, kFull, kHourly, kMaxPairs, kSpecialise ? 0 : -1>;
}

template <int kNS, bool kFull, bool kHourly, int kMaxPairs>
cudaError_t launch_kernel_ws(const ChainParams &p, int grid, int pairs, cudaStream_t stream)
{
    auto kern = pick_kernel_ws<kNS, kFull, kHourly, kMaxPairs>(p);
    const size_t smem = sizeof(double) * ((size_t)p.spline.n * (kSplineRow + 1) + p.ds.nx + p.ds.ny + 1 + (size_t)pairs * kWsRingDoubles) +
                        sizeof(int) * (size_t)pairs;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, pairs * 64, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int kNS, bool kFull, bool kHourly>
cudaError_t launch_pairs(const ChainParams &p, int grid, int pairs, cudaStream_t stream)
{
    // fewer pairs -> the instantiation with the larger register budget
    if (pairs <= kChainFewPairs) return launch_kernel_ws<kNS, kFull, kHourly, kChainFewPairs>(p, grid, pairs, stream);
    return launch_kernel_ws<kNS, kFull, kHourly, kChainMaxPairs>(p, grid, pairs, stream);
}

template <int kNS>
cudaError_t launch_pairs_ns(const ChainParams &p, bool full, bool hourly, int grid, int pairs, cudaStream_t stream)
{
    if (full) return hourly ? launch_pairs<kNS, true, true>(p, grid, pairs, stream) : launch_pairs<kNS, true, false>(p, grid, pairs, stream);
    return hourly ? launch_pairs<kNS, false, true>(p, grid, pairs, stream) : launch_pairs<kNS, false, false>(p, grid, pairs, stream);
}

template <bool kFull, bool kHourly>
cudaError_t launch_fan(const ChainParams &p, int grid, int teams, int groups, cudaStream_t stream)
{
    // summaries only: the density-dependence form compiled in (the shapes differ between the groups: run time)
    constexpr bool kSpecialise = !kFull && !kHourly;
    auto kern = chain_kernel_fan<4, kFull, kHourly, -1>;
    if (kSpecialise) kern = p.densmod_type == 1 ? chain_kernel_fan<4, kFull, kHourly, kSpecialise ? 1 : -1>
                                                : chain_kernel_fan<4, kFull, kHourly, kSpecialise ? 0 : -1>;
    const size_t smem = sizeof(double) * ((size_t)p.spline.n * (kSplineRow + 1) + p.ds.nx + p.ds.ny + 1 + (size_t)teams * kWsRingDoubles) +
                        sizeof(int) * (size_t)teams;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, teams * (groups + 1) * 32, smem, stream>>>(p, groups);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_chain_fan_fast(const ChainParams &p, bool full, bool hourly, int grid, int teams, int groups,
                                  cudaStream_t stream)
{
    if (full) return hourly ? launch_fan<true, true>(p, grid, teams, groups, stream) : launch_fan<true, false>(p, grid, teams, groups, stream);
    return hourly ? launch_fan<false, true>(p, grid, teams, groups, stream) : launch_fan<false, false>(p, grid, teams, groups, stream);
}

// settings_per_thread: 1, 2 or 4 (what launch_chain chose); grid / pairs: the persistent-CTA shape chosen by launch_one
cudaError_t launch_chain_ws_fast(const ChainParams &p, int settings_per_thread, bool full, bool hourly, int grid,
                                 int pairs, cudaStream_t stream)
{
    if (settings_per_thread == 4) return launch_pairs_ns<4>(p, full, hourly, grid, pairs, stream);
    if (settings_per_thread == 2) return launch_pairs_ns<2>(p, full, hourly, grid, pairs, stream);
    return launch_pairs_ns<1>(p, full, hourly, grid, pairs, stream);
}

}  // namespace mms
