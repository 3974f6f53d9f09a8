// chain_team.cuh -- team form of the fused chain (fast math mode only): the forcing half of chain_ws.cuh split once
// more, so that every role fits 64 registers and 30-32 warps are resident per SM instead of 22.
//
// Why: the pair form is bound by the dependent-issue latency of two in-order instruction streams per 32 cells
// (profiles/r02_pair_kernel_binder.md: a warp issues once per ~8 cycles, 5.5 warps per SM sub-partition fill 65 % of
// the issue slots, and 80 registers per thread -- 22 warps x 2560 -- is the cap).  The only levers left are more
// resident warps and shorter streams.  A TEAM serves one 32-cell unit with 2 + kG warps:
//   * climate warp : monthly -> daily -> hourly microclimate, evaporation rate, the ds(T, RH) look-up
//   * rates warp   : the three life-history splines of the hour's temperature, ds^(1/24), the emergence
//                    fraction, the density-free link of das0
//   * kG state warps, ONE habitat setting each: water body and population recursions, summaries of the state
// The hand-off is a software pipeline of depth three behind ONE named barrier per team and hour (bar.sync id,
// (2 + kG) * 32): in step t the climate warp fills slot t of the climate ring (4 slots), the rates warp turns hour
// t - 1 of it into slot t - 1 of the rates ring (2 slots), the state warps consume hour t - 2 of both.  Every role
// makes n + 2 barrier calls in a pass of n hours (the lagging roles idle at the start, the leading ones at the end),
// so the warps of a team stay aligned over passes and units.
//
// Arithmetic is exactly that of the other fast forms (same device functions on the same values; the results are
// bit for bit those of the pair form, tests/test_gpu_chain.py::test_team_form).
//
// MEASURED (one B200, BASELINE configs[1] shapes at each form's own even-round size, cells per ms of kernel):
//   two settings: pair form 104 192 cells 50.64 ms = 2058 / ms;  team form 75 776 cells 37.36 ms = 2028 / ms
//   one setting : monolithic 151 552 cells 55.80 ms = 2716 / ms; team form 94 720 cells 39.10 ms = 2423 / ms;
//                 pair form 104 192 cells 45.23 ms = 2304 / ms
// (and 100k cells x 2 settings: 52.7 ms in three rounds of eight teams against 50.6 ms in two rounds of eleven pairs).
// No hour loop spills at 64 registers (tools/sass_loops.py: climate 153-167, rates 166-181, state 185-200
// instructions per hour), the issue-slot utilisation rises from 0.67 to 0.71 -- and the 8 % more instructions per
// cell-hour (three loops, two hand-offs) and the coarser rounds take it back.  The same holds for the monolithic
// kernel at 20 warps x 96 registers instead of 16 x 128 (one setting: 2501 against 2716 cells / ms): more resident
// warps do not buy throughput here.  NOT the default: selected with MMS_CHAIN_TEAM=1 only (kept, parity-tested, as
// the record of the experiment; profiles/r02_pair_kernel_binder.md).
#pragma once

#include "chain_ws.cuh"

#if !MMS_STRICT

namespace mms {
namespace MMS_MODE {

static_assert(kWsK == 1, "the team form hands over one hour at a time");

// bit 0: the state warps (not the rates warp) evaluate the density-free link of das0 -- the balance between the
// rates role and the state role
#ifndef MMS_TEAM_SPLIT
#define MMS_TEAM_SPLIT 0
#endif
constexpr int kTeamSplit = MMS_TEAM_SPLIT;

constexpr int kTeamClimateSlots = 4, kTeamClimateFields = 4;      // air temperature, ds, evaporation rate, rain gain
constexpr int kTeamRatesSlots = 2, kTeamRatesFields = 4;          // ds^(1/24), exp(-mdr/24), eggs, link of das0
constexpr int kTeamRingDoubles = (kTeamClimateSlots * kTeamClimateFields + kTeamRatesSlots * kTeamRatesFields) * 32;

struct TeamEnd {
    double *climate;        // [4 slots][4 fields][32], this lane's column
    double *rates;          // [2 slots][4 fields][32], this lane's column
    int bar_id, bar_threads;
    __device__ __forceinline__ void step() const { pair_barrier(bar_id, bar_threads); }
};

// what the state warps see of a team
struct TeamLink {
    const TeamEnd &t;
    template <int kNS, bool kFull, bool kHourly, bool kRecord, int kDens, int kSplit>
    __device__ __forceinline__ void pop_pass(const ChainParams &p, const UnitCtx &u, UnitState<kNS> &st, int n_days)
    {
        state_pass<kNS, kFull, kHourly, false, kRecord, kDens, kSplit>(p, u, st, *this, n_days);
    }
    __device__ __forceinline__ void begin_pass() { t.step(); }
    __device__ __forceinline__ void end_pass() { t.step(); }
    __device__ __forceinline__ void fetch(int h, int, double &evap, PopHour &ph, double &gain_m_h)
    {
        t.step();
        const double *c = t.climate + (size_t)(h & (kTeamClimateSlots - 1)) * (kTeamClimateFields * 32);
        const double *r = t.rates + (size_t)(h & (kTeamRatesSlots - 1)) * (kTeamRatesFields * 32);
        evap = c[2 * 32];
        gain_m_h = c[3 * 32];
        ph.surv_ad = r[0 * 32];
        ph.one_m_emf = r[1 * 32];
        ph.egg = r[2 * 32];
        ph.link0 = r[3 * 32];
    }
};

// ---- climate warp ----------------------------------------------------------------------------------------------
template <bool kFull, bool kHourly, bool kWaterOnly, bool kRecord>
__device__ __forceinline__ void climate_pass(const ChainParams &p, const UnitCtx &u, const TeamEnd &team, int n_days,
                                             int &bad)
{
    constexpr int kNF = kFull ? 6 : 1;                 // ds, then air temperature, humidity, wind speed, rainfall, cloud cover
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    const int vid[6] = {MMS_V_DS, MMS_V_AIR_TEMPERATURE, MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
    double acc[kNF];
#pragma unroll
    for (int k = 0; k < kNF; ++k) acc[k] = 0.0;
    PeriodCursor pc;
    pc.init();

    auto flush = [&](int period) {
        if (!u.live) return;
        const int64_t at = (int64_t)period * u.os + u.ocell;
        const double n = (double)pc.acc_n;
#pragma unroll
        for (int k = 0; k < kNF; ++k) {
            if (p.mean_mask & (1u << vid[k])) {
                const int64_t row = out_slot(p.mean_mask, ~kState, vid[k]);
                p.mean_forcing[row * p.P * u.os + at] = acc[k] / n;
            }
        }
    };

    DayCtx dc;
    dc.cur = load_daily<kFull>(p, u.cell, 0, bad);
    dc.nxt = load_daily<kFull>(p, u.cell, (p.D > 1) ? 1 : 0, bad);
    dc.A = (dc.cur.tmax - dc.cur.tmin) * 0.5;                              // R/sim_microclimate.R:217
    dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;                // :220
    dc.t_midnight_prev = (dc.t_sunset - dc.nxt.tmin) / p.hr.exp_e[23] + dc.nxt.tmin;   // :271-272
    dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;

    for (int day = 0; day < n_days; ++day) {
        if (kRecord) {
            bool fresh;
            const int to_flush = pc.begin_day(p, day, fresh);
            if (fresh) {
                if (to_flush >= 0) flush(to_flush);
#pragma unroll
                for (int k = 0; k < kNF; ++k) acc[k] = 0.0;
                pc.acc_n = 0;
            }
        }
        const double rain_h = dc.cur.rain * (1.0 / 24.0);                 // R/sim_microclimate.R:132
        const double gain_m_h = rain_h * (1.0 / 1000.0);                  // R/sim_water.R:122
        double air_t_last = 0.0;
#pragma unroll 1
        for (int h = 0; h < 24; ++h) {
            HourF f;
            hour_microclimate<kFull>(p.hr, dc, h, f);
            DsFetch dsf;
            if (!kWaterOnly) dsf = ds_fetch(p.ds, f.air_t, f.rh);
            f.evap = evaporation_rate(f.air_t, f.rh, u.p_a, f.wind);      // R/sim_water.R:264-270 (air temperature)
            f.ds = kWaterOnly ? 0.0 : ds_blend(dsf);
            double *c = team.climate + (size_t)(h & (kTeamClimateSlots - 1)) * (kTeamClimateFields * 32);
            c[0 * 32] = f.air_t;
            c[1 * 32] = f.ds;
            c[2 * 32] = f.evap;
            c[3 * 32] = gain_m_h;
            air_t_last = f.air_t;
            if (kRecord) {
                acc[0] += (pc.acc_on ? f.ds : 0.0);
                if (kFull) {
                    acc[1] += (pc.acc_on ? f.air_t : 0.0); acc[2] += (pc.acc_on ? f.rh : 0.0);
                    acc[3] += (pc.acc_on ? f.wind : 0.0); acc[4] += (pc.acc_on ? rain_h : 0.0);
                    acc[5] += (pc.acc_on ? f.cloud : 0.0);
                }
                pc.acc_n += pc.acc_inc;
                if (kHourly && u.live) {
                    const int64_t tc = (int64_t)(day * 24 + h) * u.os + u.ocell;
                    const double val[6] = {f.ds, f.air_t, f.rh, f.wind, rain_h, f.cloud};
#pragma unroll
                    for (int q = 0; q < kNF; ++q) {
                        if (p.hourly_mask & (1u << vid[q])) {
                            const int64_t row = out_slot(p.hourly_mask, ~kState, vid[q]);
                            p.hourly_forcing[row * u.H * u.os + tc] = val[q];
                        }
                    }
                }
            }
            team.step();                             // hour h is in its slot
        }
        // ---- roll the day window
        dc.t_midnight_prev = air_t_last;             // hour-24 value, R/sim_microclimate.R:271
        dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;
        dc.cur = dc.nxt;
        const int d2 = day + 2;
        dc.nxt = load_daily<kFull>(p, u.cell, d2 < p.D ? d2 : p.D - 1, bad);
        dc.A = (dc.cur.tmax - dc.cur.tmin) * 0.5;
        dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;
    }
    team.step();                                     // the two steps the other roles lag behind
    team.step();
    if (kRecord && pc.cur_period >= 0) flush(pc.cur_period);
}

// ---- rates warp ------------------------------------------------------------------------------------------------
template <bool kHourly, bool kWaterOnly, bool kRecord, int kDens, int kSplit>
__device__ __forceinline__ void rates_pass(const ChainParams &p, const SmemTables &tb, const UnitCtx &u,
                                           const TeamEnd &team, int n_days)
{
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    const int vid[3] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY};
    const int dens_type = kDens >= 0 ? (kDens & 1) : p.densmod_type;
    double acc[3] = {0.0, 0.0, 0.0};
    PeriodCursor pc;
    pc.init();

    auto flush = [&](int period) {
        if (!u.live) return;
        const int64_t at = (int64_t)period * u.os + u.ocell;
        const double n = (double)pc.acc_n;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (p.mean_mask & (1u << vid[k])) {
                const int64_t row = out_slot(p.mean_mask, ~kState, vid[k]);
                p.mean_forcing[row * p.P * u.os + at] = acc[k] / n;
            }
        }
    };

    for (int day = 0; day < n_days; ++day) {
        if (kRecord) {
            bool fresh;
            const int to_flush = pc.begin_day(p, day, fresh);
            if (fresh) {
                if (to_flush >= 0) flush(to_flush);
                acc[0] = acc[1] = acc[2] = 0.0;
                pc.acc_n = 0;
            }
        }
#pragma unroll 1
        for (int h = 0; h < 24; ++h) {
            team.step();                             // the climate warp has filled hour h
            if (kWaterOnly) continue;                // water burn-in: no life history (the state warps read the climate ring only)
            const double *c = team.climate + (size_t)(h & (kTeamClimateSlots - 1)) * (kTeamClimateFields * 32);
            const double air_t = c[0 * 32], ds = c[1 * 32];
            // water_temperature = air_temperature, R/sim_microclimate.R:163; R/sim_pop_dynamics.R:100-135
            const int i = spline_interval(p.spline, tb.sx, air_t);
            const double dx = air_t - tb.sx[i];
            const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
            const double4 c0 = cf[0], c1 = cf[1], c2 = cf[2];
            const double mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
            const double efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
            const double das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
            double *r = team.rates + (size_t)(h & (kTeamRatesSlots - 1)) * (kTeamRatesFields * 32);
            r[0 * 32] = r_pow_pos(ds, kLitC[3]);
            r[1 * 32] = pop_emergence_complement(mdr);          // exp(-timestep * mdr); emf = 1 - it
            r[2 * 32] = efd * kLitC[3];
            r[3 * 32] = (kSplit & 1) ? das0 : pop_link0(dens_type, das0);
            if (kRecord) {
                acc[0] += (pc.acc_on ? mdr : 0.0); acc[1] += (pc.acc_on ? efd : 0.0); acc[2] += (pc.acc_on ? das0 : 0.0);
                pc.acc_n += pc.acc_inc;
                if (kHourly && u.live) {
                    const int64_t tc = (int64_t)(day * 24 + h) * u.os + u.ocell;
                    const double val[3] = {mdr, efd, das0};
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (p.hourly_mask & (1u << vid[q])) {
                            const int64_t row = out_slot(p.hourly_mask, ~kState, vid[q]);
                            p.hourly_forcing[row * u.H * u.os + tc] = val[q];
                        }
                    }
                }
            }
        }
    }
    team.step();                                     // hour n - 1 is in its slot
    team.step();
    if (kRecord && pc.cur_period >= 0) flush(pc.cur_period);
}

// ---- the kernel: warps [team * (2 + kG), ...) = climate, rates, kG state warps of one team --------------------
// kDens: -1, or what the launcher knows at compile time: bit 0 the density-dependence form, then two bits with the
// water-body shape + 1 of every setting when all share one (0 = read hab[].shape)
template <int kG, bool kFull, bool kHourly, int kMaxTeams, int kDens>
__global__ void __launch_bounds__(kMaxTeams * (2 + kG) * 32, 1)
chain_kernel_team(const __grid_constant__ ChainParams p)
{
    constexpr int kTeamWarps = 2 + kG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_coef = reinterpret_cast<double *>(smem_raw);              // [n][kSplineRow]
    double *s_x = s_coef + (size_t)p.spline.n * kSplineRow;             // [n]
    double *s_gx = s_x + p.spline.n;                                    // [nx]
    double *s_gy = s_gx + p.ds.nx;                                      // [ny]
    double *s_ring = s_gy + p.ds.ny;                                    // [teams][kTeamRingDoubles]
    const int n_teams = (blockDim.x >> 5) / kTeamWarps;
    int *s_unit = reinterpret_cast<int *>(s_ring + (size_t)n_teams * kTeamRingDoubles);   // [teams]
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[(k / 12) * kSplineRow + k % 12] = p.spline.coef[k];
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.ds.nx; k += blockDim.x) s_gx[k] = p.ds.x[k];
    for (int k = threadIdx.x; k < p.ds.ny; k += blockDim.x) s_gy[k] = p.ds.y[k];
    ::mms::fm::math_tables_init();
    __syncthreads();
    SmemTables tb{s_x, s_coef, s_gx, s_gy};

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team_id = warp / kTeamWarps;
    if (team_id >= n_teams) return;
    // the roles of consecutive teams are rotated over the warp slots, so that every SM sub-partition (warp % 4)
    // hosts the same mix of the (unequal) roles
    const int role = (warp % kTeamWarps + team_id) % kTeamWarps;        // 0 climate, 1 rates, 2.. state
    TeamEnd team;
    team.climate = s_ring + (size_t)team_id * kTeamRingDoubles + lane;
    team.rates = team.climate + kTeamClimateSlots * kTeamClimateFields * 32;
    team.bar_id = team_id;
    team.bar_threads = kTeamWarps * 32;

    for (;;) {
        if (role == 0 && lane == 0) s_unit[team_id] = atomicAdd(p.unit_counter, 1);
        team.step();
        const int unit = s_unit[team_id];
        team.step();                                 // everybody has read it before it is overwritten
        if (unit >= p.n_units) break;
        const int cb = unit % p.n_cell_blocks, g = unit / p.n_cell_blocks;

        UnitCtx u;
        u.cell = (int64_t)cb * 32 + lane; u.cs = p.c_stride; u.H = 24 * p.D;
        u.set0 = g * kG + (role >= 2 ? role - 2 : 0);
        u.live = u.cell < p.n_cells; u.write_forcing = g == 0;
        u.os = p.out_stride;
        u.ocell = (p.out_col && u.live) ? __ldg(p.out_col + u.cell) : u.cell;
        u.n_mean_state = __popc(p.mean_mask & MMS_STATE_VAR_MASK);
        u.n_hourly_state = __popc(p.hourly_mask & MMS_STATE_VAR_MASK);
        u.p_a = 0.0;
        if (role == 0) {
            u.p_a = air_pressure_kpa(__ldg(p.altitude + u.cell));
            u.live = u.live && u.write_forcing;      // the summaries of the forcing are written once per cell
            int bad = 0;
            for (int y = 0; y < p.burnin_years_water; ++y) climate_pass<kFull, kHourly, true, false>(p, u, team, kDaysPerYear, bad);
            for (int y = 0; y < p.burnin_years_pop; ++y) climate_pass<kFull, kHourly, false, false>(p, u, team, kDaysPerYear, bad);
            climate_pass<kFull, kHourly, false, true>(p, u, team, p.D, bad);
            if (bad && u.cell < p.n_cells) atomicOr(p.error_flag, 1);
        } else if (role == 1) {
            u.live = u.live && u.write_forcing;
            for (int y = 0; y < p.burnin_years_water; ++y) rates_pass<kHourly, true, false, kDens, kTeamSplit>(p, tb, u, team, kDaysPerYear);
            for (int y = 0; y < p.burnin_years_pop; ++y) rates_pass<kHourly, false, false, kDens, kTeamSplit>(p, tb, u, team, kDaysPerYear);
            rates_pass<kHourly, false, true, kDens, kTeamSplit>(p, tb, u, team, p.D);
        } else {
            TeamLink link{team};
            // one setting per state warp: its shape is the first (only) setting's two bits of kDens
            state_unit<1, kFull, kHourly, kDens, kTeamSplit>(p, u, link);
        }
    }
}

}  // namespace MMS_MODE
}  // namespace mms

#endif  // !MMS_STRICT
