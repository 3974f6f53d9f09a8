// mms_common.cuh -- structures shared by the host side of libmmsgpu and its
// sm_100a kernels.  All citations are relative to the reference root.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mmsgpu.h"

namespace mms {

constexpr int kBlock = 128;          // cell padding granule of a shard (keeps [period][cell] rows 1 KB aligned)
#ifndef MMS_CHAIN_MAX_WARPS
#define MMS_CHAIN_MAX_WARPS 16
#endif
constexpr int kChainMaxWarps = MMS_CHAIN_MAX_WARPS;   // largest persistent CTA of the fused kernel: 512 threads, <= 128 registers / thread
constexpr int kChainFewWarps = 12;   // second instantiation: <= 384 threads, <= 168 registers / thread
#ifndef MMS_CHAIN_MAX_PAIRS
#define MMS_CHAIN_MAX_PAIRS 11
#endif
constexpr int kChainMaxPairs = MMS_CHAIN_MAX_PAIRS;   // warp-specialised kernel: 11 forcing + 11 state warps, <= 88 registers / thread
constexpr int kChainFewPairs = 10;   // second instantiation: 640 threads, <= 96 registers / thread
constexpr int kChainTeams1 = 10;     // team form, one setting: 10 teams of climate + rates + state warp = 960 threads, <= 64 registers
constexpr int kChainTeams2 = 8;      // team form, two settings: 8 teams of climate + rates + 2 state warps = 1024 threads, <= 64 registers
constexpr int kChainFanWarps = 17;   // fan-out form: 1 forcing + up to 16 state warps per team, <= 112 registers / thread
// doubles per knot of the shared-memory copy of the spline table: [mdr, efd, das][y, b, c, d] = 12, padded so that
// the rows of 32 lanes with unrelated temperatures spread over all eight 16-byte bank groups (a 96-byte stride reaches
// only four of them; 112 bytes = 7 groups, coprime with 8).  MEASURED, one B200: 12 -> 14 takes the pair kernel on
// BASELINE configs[1] from 51.39 to 50.51 ms and the life-history verb (100k cells x 8760 h) from 18.46 to 14.91 ms.
#ifndef MMS_SPLINE_ROW
#define MMS_SPLINE_ROW 14
#endif
constexpr int kSplineRow = MMS_SPLINE_ROW;
constexpr int kHoursPerYear = 8760;  // n_times_year <- 365 * 24, R/sim_water.R:229, R/sim_pop_dynamics.R:222
constexpr int kDaysPerYear = 365;

// Hour-of-day constants of the SINEC / VSINE curves.  They depend only on the
// hour index, so they are evaluated once on the host with the host libm (the
// same one R uses) and handed to the kernel in parameter (constant) space.
struct HourTables {
    double z_light[24];    // sin((360*(t-950)/1400)/57.29577), R/sim_microclimate.R:255
    double exp_e[24];      // exp(tau * time_since_ss), :260-262
    double predawn[24];    // 1 - t/600, :275
    double w_vsine[24];    // vsine_weights(), :431-454
    double w_vsine_inv[24];// 1 - weight, :498-500
    double air_coef[24];   // fast mode: predawn[h] | z_light[h] | 1 / exp_e[h] by segment
    double z_night;        // :210
    double cos_hour_angle[24];  // water-temperature EXTENSION: cos(15 (hour - 12) degrees), query_terraclimate.R:459-466
};

// Water-temperature model (EXTENSION): per-launch constants and per-day / per-cell tables
struct WaterTempDev {
    int32_t enabled; int32_t pad;
    double solar_gain;           // (1 - shade)(1 - albedo)
    double sigma_eps;            // 5.67e-8 * emissivity
    double convection;           // W / m2 / K
    const double2 *sun_day;      // [D] (sin delta, cos delta) of each day
    const int32_t *month_of_day; // [D] 1..12
    const double *sin_lat, *cos_lat;   // [c_stride]
    const double *clear_sky;     // [12][c_stride]
};

// Per-setting water-body constants, derived on the host exactly as
// simulate_ephemeral_habitat_vectorised() derives them (R/sim_water.R:257-261).
struct HabitatDev {
    double max_volume;     // max_cone_volume
    double catchment_area; // inflow_multiplier * surface(max_volume)
    double leak_per_hour;  // EXTENSION
    double initial_volume;
    double cyl_area;       // pi r^2 (cylinder, EXTENSION)
    int32_t shape;
    int32_t pad;
};

struct DayEntry {          // approx1() interval of one prediction day (same for all cells)
    int32_t idx;           // i: knots i, i+1
    int32_t flag;          // 0 = interpolate, 1 = v == x[j] -> y[j], 2 = v == x[i] -> y[i]
    double  w;             // (v - x[i]) / (x[j] - x[i]), R approx1()
};

// Device-resident spline tables: one knot vector shared by the three
// positive_spline() closures, coefficients interleaved [knot][function][y,b,c,d].
struct SplineDev {
    const double *x;       // [n]
    const double *coef;    // [n][3][4]  function order: mdr, efd, das
    int32_t n;
    int32_t uniform;       // 1: x[i] == x0 + i*step bit for bit (then x is not read)
    double x0, step, inv_step;
};

// Device-resident ds(T, RH) surface: quad table, one 32-byte record per
// bilinear cell (z00, z10, z01, z11), record (i, j) at [i + (nx-1)*j].
struct SurfaceDev {
    const double *x, *y;   // [nx], [ny]
    const double *quad;    // [(nx-1)*(ny-1)][4]
    int32_t nx, ny;
    double x0, inv_dx, y0, inv_dy;   // strict: index guess only (intervals are verified against x, y);
                                     // fast: the index itself (host verified the axes are equally spaced)
    double x1, y1;                   // last grid values
};

struct ChainParams {
    // per-cell inputs (device), cell index fastest, row stride = c_stride
    const double *monthly;     // [9][M][c_stride]; rainfall row holds log1p(rainfall_daily)
    const double *altitude;    // [c_stride]
    const DayEntry *days;      // [D]
    const int32_t *period_of_day; // [D] or nullptr
    int64_t c_stride;          // row stride of every [...][cell] array of this shard (its padded number of cells)
    int64_t n_cells;           // real cells of this launch (a launch covers a column window of the shard's arrays:
                               // all pointers below are already offset to the window's first cell)
    int64_t n_cols;            // columns this launch integrates (n_cells rounded up to whole 32-cell units, <= c_stride)
    int32_t M, D, S, P;
    int32_t burnin_years_water, burnin_years_pop;
    double initial_adult, initial_aquatic;
    // density dependence
    int32_t densmod_type; int32_t pad0;
    double dd_effect, surface_area_cm2;
    SplineDev spline;
    SurfaceDev ds;
    // outputs (device) -- same layouts as the host ones with C -> out_stride (= c_stride unless the outputs are
    // scattered into raster cell order: then cell c of the launch writes column out_col[c] of rows out_stride long)
    int64_t out_stride;
    const int64_t *out_col;    // [n_cols] or nullptr (column = cell)
    double *mean_state, *mean_forcing, *hourly_state, *hourly_forcing;
    int32_t *wet_hours;
    uint8_t *wet_flags;
    uint32_t mean_mask, hourly_mask;
    int32_t *error_flag;       // set non-zero by the kernel on non-finite forcing
    int32_t *unit_counter;     // work-unit dispenser, zeroed before every launch
    int32_t n_units;           // n_cell_blocks * (S / settings per thread)
    int32_t n_cell_blocks;     // n_cols / 32
    int32_t force_warps;       // > 0: cap on warps (pairs) per CTA (tuning aid, env MMS_CHAIN_WARPS)
    int32_t warp_specialised;  // fast mode: 1 = chain_ws.cuh (forcing / state warp pairs), 0 = monolithic, -1 = auto
    WaterTempDev wt;
    HourTables hr;
    HabitatDev hab[MMS_MAX_SETTINGS];
};

// launchers, one translation unit per math mode (chain_fast.cu / chain_strict.cu)
// (p.unit_counter must be zero when the kernel starts; n_units / n_cell_blocks are filled in here)
cudaError_t launch_chain_fast(ChainParams &p, int n_sm, cudaStream_t stream);
cudaError_t launch_chain_strict(ChainParams &p, int n_sm, cudaStream_t stream);
// the warp-pair form (chain_ws_tu.cu), called by launch_chain_fast when it picks that form
cudaError_t launch_chain_ws_fast(const ChainParams &p, int settings_per_thread, bool full, bool hourly, int grid,
                                 int pairs, cudaStream_t stream);
// the team form (chain_team_tu.cu): `teams` teams of a climate, a rates and settings_per_unit (1 or 2) state warps per CTA
cudaError_t launch_chain_team_fast(const ChainParams &p, int settings_per_unit, bool full, bool hourly, int grid,
                                   int teams, cudaStream_t stream);
// the fan-out form (chain_ws_tu.cu): `teams` teams of 1 forcing + `groups` state warps (4 settings each) per CTA
cudaError_t launch_chain_fan_fast(const ChainParams &p, bool full, bool hourly, int grid, int teams, int groups,
                                  cudaStream_t stream);
// monthly rainfall -> log1p(rainfall) in place (R/sim_microclimate.R:33): n_rows rows of `width` doubles, `pitch` apart
cudaError_t launch_log1p_rows_fast(double *rows, int64_t n_rows, int64_t width, int64_t pitch, cudaStream_t stream);
cudaError_t launch_log1p_rows_strict(double *rows, int64_t n_rows, int64_t width, int64_t pitch, cudaStream_t stream);

// work units (32 cells x one group of settings) one round of the persistent fused kernel holds on a device with
// n_sm SMs, and the settings a thread integrates -- what the host pipeline sizes its chunks with (chain_tu.cu
// launch_one picks the same forms)
inline int chain_settings_per_thread(int S) { return S % 4 == 0 ? 4 : (S % 2 == 0 ? 2 : 1); }
// fan-out form (chain_ws.cuh chain_kernel_fan): fast mode, four settings per state warp, at least two state warps
inline int chain_fan_groups(int S, bool strict) { return (!strict && S % 4 == 0 && S >= 8) ? S / 4 : 0; }
inline int chain_fan_teams_max(int groups) { return groups > 0 ? (kChainFanWarps / (groups + 1) < 15 ? kChainFanWarps / (groups + 1) : 15) : 0; }
inline int chain_round_units(int S, bool strict, int n_sm)
{
    const int groups = chain_fan_groups(S, strict);
    if (groups > 0) return n_sm * chain_fan_teams_max(groups) * groups;    // units are (cell block, group of 4 settings)
    const bool pairs = !strict && chain_settings_per_thread(S) == 2;
    return n_sm * (pairs ? kChainMaxPairs : kChainMaxWarps);
}

// state-free verbs (verb_kernels.cuh); p carries monthly / days / hr / tables as for the chain
cudaError_t launch_daily_fast(const ChainParams &p, double *out_daily, cudaStream_t stream);
cudaError_t launch_daily_strict(const ChainParams &p, double *out_daily, cudaStream_t stream);
cudaError_t launch_hourly_fast(const ChainParams &p, const double *daily, double *out_hourly, cudaStream_t stream);
cudaError_t launch_hourly_strict(const ChainParams &p, const double *daily, double *out_hourly, cudaStream_t stream);
cudaError_t launch_lifehistory_fast(const ChainParams &p, int64_t n, const double *water_t, const double *air_t,
                                    const double *humidity, double *mdr, double *efd, double *das0, double *ds,
                                    int n_sm, cudaStream_t stream);
cudaError_t launch_lifehistory_strict(const ChainParams &p, int64_t n, const double *water_t, const double *air_t,
                                      const double *humidity, double *mdr, double *efd, double *das0, double *ds,
                                      int n_sm, cudaStream_t stream);

// summarise_vectors() / aggregate_vectors(), R/pipeline_functions.R:995-1091 (verb_kernels.cuh, summarise_kernel)
cudaError_t launch_summarise_fast(const double *hourly, const int32_t *period_of_day, int64_t c_stride, int32_t D,
                                  int32_t P, int32_t n_vars, double *means, cudaStream_t stream);
cudaError_t launch_summarise_strict(const double *hourly, const int32_t *period_of_day, int64_t c_stride, int32_t D,
                                    int32_t P, int32_t n_vars, double *means, cudaStream_t stream);

// process_terraclimate_pixel_vars(), R/terraclimate.R:581-643 (verb_kernels.cuh, terraclimate_kernel)
struct TerraParams {
    const double *terraclimate;   // [6][M][c_stride]: tmax, tmin, ppt, ws, vpd, srad
    const double *clear_sky;      // [12][c_stride]: clear_sky_SOLR by month_id
    const int32_t *month_id;      // [M], 1..12
    const double *n_days;         // [M], as.numeric(end - start)
    double *monthly;              // [9][M][c_stride], MMS_F_* order
    int64_t c_stride, n_cells;    // row stride; real cells of this launch (pointers offset to its first cell)
    int64_t n_cols;               // columns written (cells beyond n_cells get zeros), <= c_stride
    int32_t M;
    int32_t log1p_rainfall;       // 1: store log1p(rainfall_daily) (what the fused kernel reads, R/sim_microclimate.R:33)
    double wind_factor;           // (height_required / height_measured) ^ exponent, host libm like R (R/utils.R:132)
};
cudaError_t launch_terraclimate_fast(const TerraParams &p, cudaStream_t stream);
cudaError_t launch_terraclimate_strict(const TerraParams &p, cudaStream_t stream);

struct WaterStageParams {
    const double *rainfall, *air_temperature, *humidity, *windspeed; // [C][H] (R column-major)
    const double *altitude;   // [C]
    double *water_surface_area, *water_volume;  // [C][H]
    uint8_t *wet;             // [C][H]
    int64_t H, C;
    int64_t burnin_steps;     // rows repeated per burn-in pass: 8760 (vectorised rule) or H (scalar twins)
    int32_t burnin_years;
    HabitatDev hab;
};
struct PopStageParams {
    const double *larval_habitat_area, *mdr, *efd, *das_zerodensity, *ds; // [C][H]
    double *aquatic, *adult;
    int64_t H, C;
    int64_t burnin_steps;     // as above
    int32_t burnin_years;
    int32_t densmod_type;
    double dd_effect, surface_area_cm2, initial_adult, initial_aquatic;
};
cudaError_t launch_water_stage_fast(const WaterStageParams &p, cudaStream_t stream);
cudaError_t launch_water_stage_strict(const WaterStageParams &p, cudaStream_t stream);
cudaError_t launch_pop_stage_fast(const PopStageParams &p, cudaStream_t stream);
cudaError_t launch_pop_stage_strict(const PopStageParams &p, cudaStream_t stream);

}  // namespace mms
