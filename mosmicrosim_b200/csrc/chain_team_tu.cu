// chain_team_tu.cu -- translation unit of the team form of the fused kernel (chain_team.cuh; fast math only).
#include "chain_team.cuh"

namespace mms {

using namespace MMS_MODE;

namespace {

template <int kG, bool kFull, bool kHourly>
cudaError_t launch_team(const ChainParams &p, int grid, int teams, cudaStream_t stream)
{
    constexpr int kMaxTeams = kG == 1 ? kChainTeams1 : kChainTeams2;
    // summaries only: the density-dependence form compiled in, and the water-body shape when all settings share it
    constexpr bool kSpecialise = !kFull && !kHourly;
    auto kern = chain_kernel_team<kG, kFull, kHourly, kMaxTeams, -1>;
    if (kSpecialise) {
        bool same = true;
        for (int s = 1; s < p.S; ++s) same = same && p.hab[s].shape == p.hab[0].shape;
        const int spec = (p.densmod_type == 1 ? 1 : 0) | (same ? (p.hab[0].shape + 1) << 1 : 0);
        switch (spec) {
#define MMS_TEAM_CASE(v) case v: kern = chain_kernel_team<kG, kFull, kHourly, kMaxTeams, kSpecialise ? v : -1>; break;
            MMS_TEAM_CASE(0) MMS_TEAM_CASE(1) MMS_TEAM_CASE(2) MMS_TEAM_CASE(3) MMS_TEAM_CASE(4) MMS_TEAM_CASE(5)
#undef MMS_TEAM_CASE
        default: break;
        }
    }
    const size_t smem = sizeof(double) * ((size_t)p.spline.n * (kSplineRow + 1) + p.ds.nx + p.ds.ny + (size_t)teams * kTeamRingDoubles) +
                        sizeof(int) * (size_t)teams;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, teams * (2 + kG) * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int kG>
cudaError_t launch_team_g(const ChainParams &p, bool full, bool hourly, int grid, int teams, cudaStream_t stream)
{
    if (full) return hourly ? launch_team<kG, true, true>(p, grid, teams, stream) : launch_team<kG, true, false>(p, grid, teams, stream);
    return hourly ? launch_team<kG, false, true>(p, grid, teams, stream) : launch_team<kG, false, false>(p, grid, teams, stream);
}

}  // namespace

// settings_per_unit: 1 or 2 = the state warps of a team (one setting each)
cudaError_t launch_chain_team_fast(const ChainParams &p, int settings_per_unit, bool full, bool hourly, int grid,
                                   int teams, cudaStream_t stream)
{
    if (settings_per_unit == 2) return launch_team_g<2>(p, full, hourly, grid, teams, stream);
    return launch_team_g<1>(p, full, hourly, grid, teams, stream);
}

}  // namespace mms
