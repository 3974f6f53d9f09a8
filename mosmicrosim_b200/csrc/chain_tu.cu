// chain_tu.cu -- kernel translation unit, compiled twice by build.py:
//   -DMMS_STRICT=0              -> chain_fast.o   (launch_*_fast)
//   -DMMS_STRICT=1 -fmad=false  -> chain_strict.o (launch_*_strict)
// (the warp-pair form of the fused kernel is a translation unit of its own, chain_ws_tu.cu, so that the
// two long compilations run side by side)
#include "chain_kernel.cuh"
#include "stage_kernels.cuh"
#include "verb_kernels.cuh"

namespace mms {

#if MMS_STRICT
#define MMS_LAUNCH(name) name##_strict
#else
#define MMS_LAUNCH(name) name##_fast
#endif

using namespace MMS_MODE;

namespace {

template <int kNS, bool kFull, bool kHourly, int kMaxWarps>
cudaError_t launch_kernel(const ChainParams &p, int grid, int warps, cudaStream_t stream)
{
    // summaries only (the production shape of the call): the density-dependence form is compiled in, and with
    // one setting per thread also the water-body shape when all settings share it (spec = dens | (shape + 1) << 1)
    constexpr bool kSpecialise = !kFull && !kHourly;
    constexpr bool kShapeToo = kSpecialise && kNS == 1;
    int shape_all = p.hab[0].shape;
    for (int s = 1; s < p.S; ++s)
        if (p.hab[s].shape != shape_all) shape_all = -1;
    const int dens = p.densmod_type == 1 ? 1 : 0;
    auto kern = chain_kernel<kNS, kFull, kHourly, kMaxWarps, -1>;
    // the water-temperature EXTENSION: one setting per thread, the all-variables instantiation (launch_chain sees to both)
    if (p.wt.enabled) kern = chain_kernel<kNS, kFull, kHourly, kMaxWarps, -1, (kNS == 1 && kFull)>;
    if (kSpecialise) {
        kern = dens ? chain_kernel<kNS, kFull, kHourly, kMaxWarps, kSpecialise ? 1 : -1>
                    : chain_kernel<kNS, kFull, kHourly, kMaxWarps, kSpecialise ? 0 : -1>;
        if (kShapeToo && shape_all == 0)
            kern = dens ? chain_kernel<kNS, kFull, kHourly, kMaxWarps, kShapeToo ? 3 : -1>
                        : chain_kernel<kNS, kFull, kHourly, kMaxWarps, kShapeToo ? 2 : -1>;
        if (kShapeToo && shape_all == 1)
            kern = dens ? chain_kernel<kNS, kFull, kHourly, kMaxWarps, kShapeToo ? 5 : -1>
                        : chain_kernel<kNS, kFull, kHourly, kMaxWarps, kShapeToo ? 4 : -1>;
    }
    const size_t smem = sizeof(double) * ((size_t)p.spline.n * (kSplineRow + 1) + p.ds.nx + p.ds.ny);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, warps * 32, smem, stream>>>(p);
    return cudaGetLastError();
}


template <int kNS, bool kFull, bool kHourly>
cudaError_t launch_one(ChainParams &p, int n_sm, cudaStream_t stream)
{
    // work units: 32 cells x one group of kNS settings.  Pick the number of warps
    // per persistent CTA so that the rounds come out even: the fewest rounds that
    // fit the largest CTA, then the fewest warps that still need only that many.
    p.n_cell_blocks = (int32_t)(p.n_cols / 32);
    p.n_units = p.n_cell_blocks * (p.S / kNS);
#if !MMS_STRICT
    // Habitat ensembles: the fan-out form shares the forcing half among all settings of a cell (chain_ws.cuh).
    // MMS_CHAIN_FAN = 0 / 1 forces it off / on (tests, measurements).
    {
        int groups = kNS == 4 ? chain_fan_groups(p.S, false) : 0;
        if (p.wt.enabled || ((p.mean_mask | p.hourly_mask) & (1u << MMS_V_WATER_TEMPERATURE)) || p.warp_specialised == 0) groups = 0;
        if (const char *e = getenv("MMS_CHAIN_FAN")) { if (atoi(e) == 0) groups = 0; }
        if (groups > 0) {
            const int teams_max = chain_fan_teams_max(groups);
            const int64_t per_round = (int64_t)n_sm * teams_max;
            const int64_t rounds = (p.n_cell_blocks + per_round - 1) / per_round;
            int teams = (int)((p.n_cell_blocks + (int64_t)n_sm * rounds - 1) / ((int64_t)n_sm * rounds));
            teams = teams < 1 ? 1 : (teams > teams_max ? teams_max : teams);
            const int64_t ctas = (p.n_cell_blocks + teams - 1) / teams;
            return launch_chain_fan_fast(p, kFull, kHourly, (int)(ctas < n_sm ? ctas : n_sm), teams, groups, stream);
        }
    }
    // Team form (chain_team.cuh): climate / rates / state warps at <= 64 registers, 30-32 warps per SM.
    // MMS_CHAIN_TEAM = 1 selects it (one or two settings per unit, no water-temperature model).
    if ((kNS == 1 || kNS == 2) && !p.wt.enabled && !((p.mean_mask | p.hourly_mask) & (1u << MMS_V_WATER_TEMPERATURE))) {
        const char *e = getenv("MMS_CHAIN_TEAM");
        if (e && atoi(e) != 0) {
            const int teams_cap = kNS == 1 ? kChainTeams1 : kChainTeams2;
            const int teams_max = p.force_warps > 0 && p.force_warps < teams_cap ? p.force_warps : teams_cap;
            const int64_t per_round = (int64_t)n_sm * teams_max;
            const int64_t rounds = (p.n_units + per_round - 1) / per_round;
            int teams = (int)((p.n_units + (int64_t)n_sm * rounds - 1) / ((int64_t)n_sm * rounds));
            teams = teams < 1 ? 1 : (teams > teams_max ? teams_max : teams);
            const int64_t ctas = (p.n_units + teams - 1) / teams;
            return launch_chain_team_fast(p, kNS, kFull, kHourly, (int)(ctas < n_sm ? ctas : n_sm), teams, stream);
        }
    }
    // Which form?  Measured on B200 (profiles/r01_*): the pair form wins when the launch is
    // short of warps -- several settings per cell and at most two rounds of units per SM (the
    // 100k-cell tile of BASELINE configs[1]: 58.5 vs 67.4 ms) -- and loses a little when the
    // monolithic kernel can already fill 16 warps per SM, with one setting per cell (100k cells:
    // 54.0 vs 50.2 ms) or with four settings per thread (132 vs 118 ms).
    bool ws = p.warp_specialised >= 0 ? p.warp_specialised != 0
                                      : (kNS == 2 && p.n_units <= (int64_t)2 * n_sm * kChainMaxPairs);
    // the water temperature (as an output, or modelled) lives in the monolithic form only
    if (p.wt.enabled || ((p.mean_mask | p.hourly_mask) & (1u << MMS_V_WATER_TEMPERATURE))) ws = false;
    if (ws) {
        // warp-specialised form: a pair of warps (forcing + state) per unit, chain_ws.cuh
        const int pairs_max = p.force_warps > 0 ? p.force_warps : kChainMaxPairs;
        const int64_t per_round = (int64_t)n_sm * pairs_max;
        const int64_t rounds = (p.n_units + per_round - 1) / per_round;
        int pairs = (int)((p.n_units + (int64_t)n_sm * rounds - 1) / ((int64_t)n_sm * rounds));
        pairs = pairs < 1 ? 1 : (pairs > pairs_max ? pairs_max : pairs);
        const int64_t ctas = (p.n_units + pairs - 1) / pairs;
        const int grid = (int)(ctas < n_sm ? ctas : n_sm);
        return launch_chain_ws_fast(p, kNS, kFull, kHourly, grid, pairs, stream);
    }
#endif
    const int w_max = p.force_warps > 0 ? p.force_warps : kChainMaxWarps;
    const int64_t per_round = (int64_t)n_sm * w_max;
    const int64_t rounds = (p.n_units + per_round - 1) / per_round;
    int warps = (int)((p.n_units + (int64_t)n_sm * rounds - 1) / ((int64_t)n_sm * rounds));
    warps = warps < 1 ? 1 : (warps > w_max ? w_max : warps);
    const int64_t ctas = (p.n_units + warps - 1) / warps;
    const int grid = (int)(ctas < n_sm ? ctas : n_sm);
    // fewer warps -> the instantiation with the larger register budget
    if (warps <= kChainFewWarps) return launch_kernel<kNS, kFull, kHourly, kChainFewWarps>(p, grid, warps, stream);
    return launch_kernel<kNS, kFull, kHourly, kChainMaxWarps>(p, grid, warps, stream);
}

template <int kNS>
cudaError_t launch_ns(ChainParams &p, int n_sm, bool full, bool hourly, cudaStream_t stream)
{
    if (full) return hourly ? launch_one<kNS, true, true>(p, n_sm, stream) : launch_one<kNS, true, false>(p, n_sm, stream);
    return hourly ? launch_one<kNS, false, true>(p, n_sm, stream) : launch_one<kNS, false, false>(p, n_sm, stream);
}

__global__ void log1p_rows_kernel(double *rows, int64_t width, int64_t pitch)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double *row = rows + (int64_t)blockIdx.y * pitch;
    if (i < width) row[i] = log1p(row[i]);                // R/sim_microclimate.R:33
}

}  // namespace

cudaError_t MMS_LAUNCH(launch_chain)(ChainParams &p, int n_sm, cudaStream_t stream)
{
    const uint32_t all = p.mean_mask | p.hourly_mask;
    const bool full = (all & ~MMS_PIPELINE_VAR_MASK) != 0 || p.wt.enabled;
    const bool hourly = p.hourly_mask != 0 || p.wet_flags != nullptr;
    if (p.wt.enabled) return launch_ns<1>(p, n_sm, true, hourly, stream);      // host checked S == 1
    // settings per thread: the largest of 4, 2, 1 that divides S
    if (p.S % 4 == 0) return launch_ns<4>(p, n_sm, full, hourly, stream);
    if (p.S % 2 == 0) return launch_ns<2>(p, n_sm, full, hourly, stream);
    return launch_ns<1>(p, n_sm, full, hourly, stream);
}

cudaError_t MMS_LAUNCH(launch_log1p_rows)(double *rows, int64_t n_rows, int64_t width, int64_t pitch, cudaStream_t stream)
{
    if (n_rows <= 0 || width <= 0) return cudaSuccess;
    dim3 grid((unsigned)((width + 255) / 256), (unsigned)n_rows);
    log1p_rows_kernel<<<grid, 256, 0, stream>>>(rows, width, pitch);
    return cudaGetLastError();
}

cudaError_t MMS_LAUNCH(launch_daily)(const ChainParams &p, double *out_daily, cudaStream_t stream)
{
    dim3 grid((unsigned)(p.c_stride / 128), (unsigned)p.D);
    daily_kernel<<<grid, 128, 0, stream>>>(p, out_daily);
    return cudaGetLastError();
}

cudaError_t MMS_LAUNCH(launch_hourly)(const ChainParams &p, const double *daily, double *out_hourly, cudaStream_t stream)
{
    dim3 grid((unsigned)(p.c_stride / 128), (unsigned)p.D);
    hourly_kernel<<<grid, 128, 0, stream>>>(p, daily, out_hourly);
    return cudaGetLastError();
}

cudaError_t MMS_LAUNCH(launch_lifehistory)(const ChainParams &p, int64_t n, const double *water_t, const double *air_t,
                                           const double *humidity, double *mdr, double *efd, double *das0, double *ds,
                                           int n_sm, cudaStream_t stream)
{
    const size_t smem = sizeof(double) * ((size_t)p.spline.n * (kSplineRow + 1) + p.ds.nx + p.ds.ny);
    cudaError_t err = cudaFuncSetAttribute(lifehistory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int64_t want = (n + 511) / 512;
    const int grid = (int)(want < 2 * (int64_t)n_sm ? want : 2 * (int64_t)n_sm);   // 2 CTAs of 256 per SM (smem bound)
    lifehistory_kernel<<<grid < 1 ? 1 : grid, 512, smem, stream>>>(p, n, water_t, air_t, humidity, mdr, efd, das0, ds);
    return cudaGetLastError();
}

cudaError_t MMS_LAUNCH(launch_summarise)(const double *hourly, const int32_t *period_of_day, int64_t c_stride, int32_t D,
                                         int32_t P, int32_t n_vars, double *means, cudaStream_t stream)
{
    dim3 grid((unsigned)(c_stride / 128), (unsigned)n_vars);
    summarise_kernel<<<grid, 128, 0, stream>>>(hourly, period_of_day, c_stride, D, P, means);
    return cudaGetLastError();
}

cudaError_t MMS_LAUNCH(launch_terraclimate)(const TerraParams &p, cudaStream_t stream)
{
    dim3 grid((unsigned)((p.n_cols + 127) / 128), (unsigned)p.M);
    terraclimate_kernel<<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

namespace {

// widest access the series allow: 4 = 32-byte stores, 2 = 16-byte copies, 1 = 8-byte
int stage_vec(int64_t H, uintptr_t all_pointers)
{
    if (H % 4 == 0 && (all_pointers & 31) == 0) return 4;
    if (H % 2 == 0 && (all_pointers & 15) == 0) return 2;
    return 1;
}

// persistent CTAs, one per SM.  The recursions are bound by FP64 latency (an exp / log / pow or three per
// step in a serial chain per pixel), so what counts is independent chains in flight per SM sub-partition
// within the register budget.  Measured on the 100k-pixel population twin (ms): 2 even rounds x 11 warps at
// 168 registers 29.0, at 128 registers 31.6; 16 + 5 warps (uneven rounds) 41.5; 22 warps at 80 registers
// (spills) 58.4; two pixels per thread (kPx = 2, 12 warps) 52.2.  Hence: one pixel per thread, the warp
// count that makes the rounds of 32-pixel units even, and the 168-register instantiation whenever that
// count is <= 12.
struct StagePlan { int grid, warps; size_t smem; };

StagePlan stage_plan(int64_t C, int n_inputs, int n_sm)
{
    bool even = true;
    int w_max = 0;
    if (const char *e = getenv("MMS_STAGE_WARPS")) w_max = atoi(e);       // tuning aids
    if (const char *e = getenv("MMS_STAGE_EVEN")) even = atoi(e) != 0;
    if (w_max < 1 || w_max > kStageMaxWarps) w_max = kStageMaxWarps;
    const int w_smem = (int)(kStageSmemBudget / stage_smem_bytes(n_inputs, 1));
    if (w_max > w_smem) w_max = w_smem;
    const int64_t n_units = (C + 31) / 32;
    const int64_t per_round = (int64_t)n_sm * w_max;
    const int64_t rounds = even ? (n_units + per_round - 1) / per_round : 1;
    int64_t warps = (n_units + (int64_t)n_sm * rounds - 1) / ((int64_t)n_sm * rounds);
    warps = warps < 1 ? 1 : (warps > w_max ? w_max : warps);
    const int64_t ctas = (n_units + warps - 1) / warps;
    return StagePlan{(int)(ctas < n_sm ? ctas : n_sm), (int)warps, stage_smem_bytes(n_inputs, (int)warps)};
}

template <typename Kern, typename Params>
cudaError_t launch_stage_kernel(Kern kern, const Params &p, const StagePlan &pl, cudaStream_t stream)
{
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (err != cudaSuccess) return err;
    kern<<<pl.grid, pl.warps * 32, pl.smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t device_sm_count(int &n_sm)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    return err;
}

// spec: the water-body shape (water twin) / the density-dependence form (population twin), compiled in
#define MMS_STAGE_DISPATCH1(kernel, V, SPEC, pl)                                                          \
    ((pl).warps <= kStageFewWarps ? launch_stage_kernel(kernel<V, kStageFewWarps, SPEC>, p, pl, stream)   \
                                  : launch_stage_kernel(kernel<V, kStageMaxWarps, SPEC>, p, pl, stream))
#define MMS_STAGE_DISPATCH(kernel, V, spec, pl)                                                           \
    ((spec) == 1 ? MMS_STAGE_DISPATCH1(kernel, V, 1, pl) : MMS_STAGE_DISPATCH1(kernel, V, 0, pl))

}  // namespace

cudaError_t MMS_LAUNCH(launch_water_stage)(const WaterStageParams &p, cudaStream_t stream)
{
    if (p.C <= 0) return cudaSuccess;
    int n_sm = 0;
    cudaError_t err = device_sm_count(n_sm);
    if (err != cudaSuccess) return err;
    const int vec = stage_vec(p.H, (uintptr_t)p.rainfall | (uintptr_t)p.air_temperature | (uintptr_t)p.humidity |
                                       (uintptr_t)p.windspeed | (uintptr_t)p.water_surface_area |
                                       (uintptr_t)p.water_volume | (uintptr_t)p.wet);
    const StagePlan pl = stage_plan(p.C, kWaterInputs, n_sm);
    if (vec == 4) return MMS_STAGE_DISPATCH(water_stage_kernel, 4, p.hab.shape, pl);
    if (vec == 2) return MMS_STAGE_DISPATCH(water_stage_kernel, 2, p.hab.shape, pl);
    return MMS_STAGE_DISPATCH(water_stage_kernel, 1, p.hab.shape, pl);
}

cudaError_t MMS_LAUNCH(launch_pop_stage)(const PopStageParams &p, cudaStream_t stream)
{
    if (p.C <= 0) return cudaSuccess;
    int n_sm = 0;
    cudaError_t err = device_sm_count(n_sm);
    if (err != cudaSuccess) return err;
    const int vec = stage_vec(p.H, (uintptr_t)p.larval_habitat_area | (uintptr_t)p.mdr | (uintptr_t)p.efd |
                                       (uintptr_t)p.das_zerodensity | (uintptr_t)p.ds | (uintptr_t)p.aquatic |
                                       (uintptr_t)p.adult);
    const StagePlan pl = stage_plan(p.C, kPopInputs, n_sm);
    if (vec == 4) return MMS_STAGE_DISPATCH(pop_stage_kernel, 4, p.densmod_type, pl);
    if (vec == 2) return MMS_STAGE_DISPATCH(pop_stage_kernel, 2, p.densmod_type, pl);
    return MMS_STAGE_DISPATCH(pop_stage_kernel, 1, p.densmod_type, pl);
}

}  // namespace mms
