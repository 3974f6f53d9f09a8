// chain_kernel.cuh -- the fused per-cell chain: monthly forcing -> daily
// (approx) -> hourly (SINEC / VSINE) -> evaporation -> water body ->
// life-history parameters -> population -> period summaries.  One thread per
// cell, the whole hourly time loop inside the kernel, state in registers.
//
// Execution shape (B200: 148 SMs, FP64 pipe = 1 warp instruction / 2 cycles / SMSP):
//   * one persistent CTA per SM; the three spline tables (96 KB) and the ds grid
//     axes are staged in shared memory ONCE per SM,
//   * work unit = 32 adjacent cells x one group of kNS habitat settings; warps
//     fetch units from a global counter, so the number of warps per CTA can be
//     chosen on the host to make the rounds come out even,
//   * the hour loop is software pipelined: the state-independent half of hour
//     h+1 (microclimate, evaporation rate, life-history tables, ds^(1/24), ...)
//     is issued next to the state-dependent half of hour h (water body and
//     population recursions), giving every warp two independent FP64 chains.
//
// Included by chain_tu.cu for MMS_STRICT=0 and 1.  Citations: reference root.
#pragma once

#include "mms_math.cuh"

namespace mms {
namespace MMS_MODE {

// ---- monthly -> daily: spline_seasonal() in-range branch, R/spline_interpolation.R:25-27,
// i.e. stats::approx / approx1() with the interval found once per day on the host side
struct Daily {
    double tmin, tmax, rhmin, rhmax, wsmin, wsmax, rain;
    double ccmin, ccmax;
};

__device__ __forceinline__ double knot_interp(const double *__restrict__ col, int64_t stride,
                                              const DayEntry &de, int &bad)
{
    const double yi = __ldg(col + (int64_t)de.idx * stride);
    const double yj = __ldg(col + (int64_t)(de.idx + 1) * stride);
    bad |= !(isfinite(yi) && isfinite(yj));
    if (de.flag == 1) return yj;                         // v == x[j]
    if (de.flag == 2) return yi;                         // v == x[i]
    return yi + (yj - yi) * de.w;                        // approx1(), linear
}

template <bool kCloud>
__device__ __forceinline__ Daily load_daily(const ChainParams &p, int64_t cell, int day, int &bad)
{
    const DayEntry de = p.days[day];                     // uniform across the warp
    const int64_t cs = p.c_stride;
    const double *base = p.monthly + cell;
    const int64_t vs = (int64_t)p.M * cs;                // variable stride
    Daily d;
    d.tmin = knot_interp(base + MMS_F_TMIN * vs, cs, de, bad);
    d.tmax = knot_interp(base + MMS_F_TMAX * vs, cs, de, bad);
    d.rhmin = knot_interp(base + MMS_F_RHMIN * vs, cs, de, bad);
    d.rhmax = knot_interp(base + MMS_F_RHMAX * vs, cs, de, bad);
    d.wsmin = knot_interp(base + MMS_F_WSMIN * vs, cs, de, bad);
    d.wsmax = knot_interp(base + MMS_F_WSMAX * vs, cs, de, bad);
    if (kCloud) {
        d.ccmin = knot_interp(base + MMS_F_CCMIN * vs, cs, de, bad);
        d.ccmax = knot_interp(base + MMS_F_CCMAX * vs, cs, de, bad);
    } else {
        d.ccmin = d.ccmax = 0.0;
    }
    // the rainfall rows hold log1p(rainfall_daily) (R/sim_microclimate.R:33);
    // back-transform pmax(0, expm1(.)), :61
    const double r = expm1(knot_interp(base + MMS_F_RAINFALL_DAILY * vs, cs, de, bad));
    d.rain = (r != r) ? r : ((0.0 < r) ? r : 0.0);
    return d;
}

// ---- life-history tables (shared memory) -------------------------------------
struct SmemTables {
    const double *sx;      // [n] spline knots
    const double *coef;    // [n][3][4]
    const double *gx, *gy; // ds grid axes
};

// spline_eval() (base R splines.c) interval: largest i with x[i] <= u, 0 below the
// first knot, n-1 at or beyond the last one.
__device__ __forceinline__ int spline_interval(const SplineDev &sp, const double *__restrict__ xs, double u)
{
    const int n = sp.n;
#if MMS_STRICT
    double g = (u - sp.x0) * sp.inv_step;
    g = fmin(fmax(g, 0.0), (double)(n - 1));             // NaN -> 0
    int i = (int)g;
    while (i > 0 && u < xs[i]) --i;
    while (i < n - 1 && !(u < xs[i + 1])) ++i;
    return i;
#else
    // fast mode runs only on knots the host verified to be equally spaced: direct index,
    // no search.  A last-bit slip to the neighbouring interval changes the value by
    // O(dx^3 * 1e-16) because the spline is C2.
    (void)xs;
    const int i = __double2int_rd((u - sp.x0) * sp.inv_step);   // NaN -> 0
    return min(max(i, 0), n - 1);
#endif
}

// fields::interp.surface() on the ds(T, RH) grid, R/lifehistory_estimation_functions.R:1144-1190
// lx = approx(x, 1:nx, v): 1-based fractional index, NaN outside [x[0], x[n-1]]
__device__ __forceinline__ double approx_index(const double *__restrict__ x, int n, double x0, double inv_dx, double v)
{
    if (!(v >= x[0]) || !(v <= x[n - 1])) return __longlong_as_double(0x7ff8000000000000LL);
    double g = (v - x0) * inv_dx;
    g = fmin(fmax(g, 0.0), (double)(n - 2));
    int i = (int)g;
    while (i > 0 && v < x[i]) --i;
    while (i < n - 2 && !(v < x[i + 1])) ++i;            // bisection result: x[i] <= v, i <= n-2
    const double xi = x[i], xj = x[i + 1];
    if (v == xj) return (double)(i + 2);
    if (v == xi) return (double)(i + 1);
    return (double)(i + 1) + 1.0 * ((v - xi) / (xj - xi));
}

__device__ __forceinline__ double ds_surface(const SurfaceDev &g, const SmemTables &tb, double temperature, double humidity)
{
#if MMS_STRICT
    double lx = approx_index(tb.gx, g.nx, g.x0, g.inv_dx, temperature);
    double ly = approx_index(tb.gy, g.ny, g.y0, g.inv_dy, humidity);
    if (lx != lx || ly != ly) return __longlong_as_double(0x7ff8000000000000LL);   // NA outside the grid
    double lx1 = floor(lx), ly1 = floor(ly);
    double ex = lx - lx1, ey = ly - ly1;
    if (lx1 == (double)g.nx) { ex = 1.0; lx1 = (double)(g.nx - 1); }
    if (ly1 == (double)g.ny) { ey = 1.0; ly1 = (double)(g.ny - 1); }
    const int i = (int)lx1 - 1, j = (int)ly1 - 1;
    const double2 *q = reinterpret_cast<const double2 *>(g.quad) + ((int64_t)i + (int64_t)(g.nx - 1) * j) * 2;
    const double2 a = __ldg(q), b = __ldg(q + 1);        // (z00, z10), (z01, z11)
    return a.x * (1.0 - ex) * (1.0 - ey) + a.y * ex * (1.0 - ey) +
           b.x * (1.0 - ex) * ey + b.y * ex * ey;
#else
    // fast mode runs only on axes the host verified to be equally spaced: the fractional
    // index is (v - x0) / dx directly, no search and no axis loads; branch free, NaN
    // (the reference's NA) outside the grid is selected at the end.
    (void)tb;
    const bool inside = temperature >= g.x0 && temperature <= g.x1 && humidity >= g.y0 && humidity <= g.y1;
    const double fx = (temperature - g.x0) * g.inv_dx, fy = (humidity - g.y0) * g.inv_dy;
    int i = __double2int_rd(fx), j = __double2int_rd(fy);     // NaN -> 0
    i = min(max(i, 0), g.nx - 2);
    j = min(max(j, 0), g.ny - 2);
    const double ex = fx - (double)i, ey = fy - (double)j;    // == 1 on the last grid line
    const double2 *q = reinterpret_cast<const double2 *>(g.quad) + (i + (g.nx - 1) * j) * 2;
    const double2 a = __ldg(q), b = __ldg(q + 1);             // (z00, z10), (z01, z11)
    const double mx = 1.0 - ex, my = 1.0 - ey;
    const double v = (a.x * mx + a.y * ex) * my + (b.x * mx + b.y * ex) * ey;
    return inside ? v : __longlong_as_double(0x7ff8000000000000LL);
#endif
}

#if !MMS_STRICT
// the same look-up in two steps, so that the caller can put independent arithmetic between
// the (L2-latency) loads and their first use
struct DsFetch {
    double2 a, b;          // (z00, z10), (z01, z11)
    double ex, ey;
    bool inside;
};

__device__ __forceinline__ DsFetch ds_fetch(const SurfaceDev &g, double temperature, double humidity)
{
    DsFetch d;
    d.inside = temperature >= g.x0 && temperature <= g.x1 && humidity >= g.y0 && humidity <= g.y1;
    const double fx = (temperature - g.x0) * g.inv_dx, fy = (humidity - g.y0) * g.inv_dy;
    int i = __double2int_rd(fx), j = __double2int_rd(fy);     // NaN -> 0
    i = min(max(i, 0), g.nx - 2);
    j = min(max(j, 0), g.ny - 2);
    d.ex = fx - (double)i;
    d.ey = fy - (double)j;                                    // == 1 on the last grid line
    const double2 *q = reinterpret_cast<const double2 *>(g.quad) + (i + (g.nx - 1) * j) * 2;
    d.a = __ldg(q);
    d.b = __ldg(q + 1);
    return d;
}

__device__ __forceinline__ double ds_blend(const DsFetch &d)
{
    const double mx = 1.0 - d.ex, my = 1.0 - d.ey;
    const double v = (d.a.x * mx + d.a.y * d.ex) * my + (d.b.x * mx + d.b.y * d.ex) * d.ey;
    return d.inside ? v : __longlong_as_double(0x7ff8000000000000LL);
}
#endif

// ---- state-independent half of one hour ---------------------------------------
struct DayCtx {          // everything the hourly curves need about yesterday / today / tomorrow
    Daily cur, nxt;
    double A, t_sunset, t_midnight_prev, rhmin_prev, wsmax_prev, ccmin_prev;
    // water-temperature EXTENSION: clear-sky irradiance of hour h = sol_scale * max(0, sol_a * cos_hour_angle[h] + sol_b)
    double sol_a, sol_b, sol_scale;
};

struct HourF {
    double air_t, rh, wind, cloud;     // hourly microclimate
    double evap;                       // evaporation rate per unit area
    double mdr, efd, das0, ds;         // life-history parameters
    PopHour ph;
};

// hourly microclimate of hour index h (0..23): SINEC air temperature, VSINE humidity / wind
// speed [/ cloud cover]; hourly_from_daily_climate_ambient(), R/sim_microclimate.R:126-173
template <bool kFull>
__device__ __forceinline__ void hour_microclimate(const HourTables &hr, const DayCtx &dc, int h, HourF &f)
{
    // ---- SINEC air temperature, R/sim_microclimate.R:185-285 (hour = h + 1, times = 100 * hour)
#if !MMS_STRICT
    {   // the three segments as one fma with warp-uniform selects (no branch in the hour loop):
        // pre-dawn  tmin + (Tmidnight_yesterday - tmin) * (1 - t/600)            :275-277
        // daylight  A * Z(t) + tmin + A                                          :255-256
        // night     (Tsunset - tmin_tomorrow) * (1 / exp(tau * tss)) + tmin_tomorrow   :260-262
        const bool pre = h < 5, night = h > 17;
        const double base = night ? dc.nxt.tmin : dc.cur.tmin;
        const double amp = pre ? (dc.t_midnight_prev - dc.cur.tmin) : (night ? (dc.t_sunset - dc.nxt.tmin) : dc.A);
        const double add = (pre || night) ? 0.0 : dc.A;
        f.air_t = fma(amp, hr.air_coef[h], base) + add;
    }
#else
    if (h < 5) {                     // times < time_sunrise: pre-dawn ramp, :275-277
        f.air_t = dc.cur.tmin + (dc.t_midnight_prev - dc.cur.tmin) * hr.predawn[h];
    } else if (h <= 17) {            // 600 <= times <= 1800: sine, :255-256
        f.air_t = dc.A * hr.z_light[h] + dc.cur.tmin + dc.A;
    } else {                         // night: exponential decay, :260-262
        f.air_t = (dc.t_sunset - dc.nxt.tmin) / hr.exp_e[h] + dc.nxt.tmin;
    }
#endif
    // ---- VSINE humidity / wind speed [/ cloud cover], :466-566
    const bool before_sunrise = h < 5;          // times < 6
    const bool after_inflection = h > 12;       // times > 13
    double inf_last = before_sunrise ? dc.rhmin_prev : dc.cur.rhmin;       // :529-530
    double sun_next = after_inflection ? dc.nxt.rhmax : dc.cur.rhmax;      // :548-549
    f.rh = inf_last + hr.w_vsine[h] * (sun_next - inf_last);             // :556-557
    inf_last = before_sunrise ? dc.wsmax_prev : dc.cur.wsmax;
    sun_next = after_inflection ? dc.nxt.wsmin : dc.cur.wsmin;
    f.wind = sun_next + hr.w_vsine_inv[h] * (inf_last - sun_next);       // :553-554
    f.cloud = 0.0;
    if (kFull) {
        inf_last = before_sunrise ? dc.ccmin_prev : dc.cur.ccmin;
        sun_next = after_inflection ? dc.nxt.ccmax : dc.cur.ccmax;
        f.cloud = inf_last + hr.w_vsine[h] * (sun_next - inf_last);
    }
}

// dens_type: the density-dependence form; callers that know it at compile time pass the constant, so that the
// other form's code (and the branch around it) disappears from their loop
// kWT (water-temperature EXTENSION): the two rates that take the water temperature (mdr, das_zerodensity) and what
// follows from them are left to the water step, which knows that temperature; the hourly cloud cover is needed
template <bool kFull, bool kWaterOnly, bool kWT = false>
__device__ __forceinline__ HourF hour_forcing(const ChainParams &p, const SmemTables &tb, const DayCtx &dc,
                                              double p_a, int h, int dens_type)
{
    HourF f;
    hour_microclimate<kFull || kWT>(p.hr, dc, h, f);
    // ---- life-history parameters, R/sim_pop_dynamics.R:100-135
    // water_temperature = air_temperature, R/sim_microclimate.R:163.  The table look-ups
    // (L2-resident ds quad table, shared-memory spline rows) are issued first so that their
    // latency overlaps the evaporation arithmetic below.
    f.mdr = f.efd = f.das0 = f.ds = 0.0;
    double dx = 0.0;
    double4 c0 = make_double4(0, 0, 0, 0), c1 = c0, c2 = c0;
#if !MMS_STRICT
    DsFetch dsf;
#endif
    if (!kWaterOnly) {
#if MMS_STRICT
        f.ds = ds_surface(p.ds, tb, f.air_t, f.rh);
#else
        dsf = ds_fetch(p.ds, f.air_t, f.rh);
#endif
        const int i = spline_interval(p.spline, tb.sx, f.air_t);
        dx = f.air_t - tb.sx[i];
        const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
        c0 = cf[0]; c1 = cf[1]; c2 = cf[2];
    }
    // ---- evaporation rate per unit area, R/sim_water.R:264-270 (air temperature)
    f.evap = evaporation_rate(f.air_t, f.rh, p_a, f.wind);
    if (!kWaterOnly) {
        // what does not need the ds quad first: its first use comes as late as the hour allows
        f.efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
        if (!kWT) {
            f.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
            f.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
            pop_hour_water(f.ph, dens_type, f.mdr, f.das0);
        }
#if !MMS_STRICT
        f.ds = ds_blend(dsf);
#endif
        pop_hour_air(f.ph, f.efd, f.ds);
    }
    return f;
}

#if !MMS_STRICT
// hour_forcing() in two phases, so that the monolithic hour loop can put the population update of the previous hour
// and the water step between the request of the ds quad (an L2 access: ncu had 10-18 % of the loop's stall samples
// on its first use, profiles/r02_ncu_chain_mono.txt) and that first use.  MEASURED (one B200, ms per launch,
// MMS_MONO_LATE_DS 0 -> 1): 100k cells x 1 setting 41.02 -> 37.87, x 4 settings 105.5 -> 97.3, 151 552 cells x 1 setting
// x 2 years 84.98 -> 80.34.  Same operations on the same values: the results do not change.
#ifndef MMS_MONO_LATE_DS
#define MMS_MONO_LATE_DS 1
#endif
template <bool kFull, bool kWaterOnly>
__device__ __forceinline__ HourF hour_forcing_early(const ChainParams &p, const DayCtx &dc, double p_a, int h, DsFetch &dsf)
{
    HourF f;
    hour_microclimate<kFull>(p.hr, dc, h, f);
    f.mdr = f.efd = f.das0 = f.ds = 0.0;
    if (!kWaterOnly) dsf = ds_fetch(p.ds, f.air_t, f.rh);
    f.evap = evaporation_rate(f.air_t, f.rh, p_a, f.wind);
    return f;
}
template <bool kWaterOnly>
__device__ __forceinline__ void hour_forcing_late(const ChainParams &p, const SmemTables &tb, HourF &f, const DsFetch &dsf,
                                                  int dens_type)
{
    if (kWaterOnly) return;
    const int i = spline_interval(p.spline, tb.sx, f.air_t);
    const double dx = f.air_t - tb.sx[i];
    const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
    const double4 c0 = cf[0], c1 = cf[1], c2 = cf[2];
    f.efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
    f.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
    f.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
    pop_hour_water(f.ph, dens_type, f.mdr, f.das0);
    f.ds = ds_blend(dsf);
    pop_hour_air(f.ph, f.efd, f.ds);
}
#endif

// ---- per-unit state carried across passes ---------------------------------------
template <int kNS>
struct UnitState {
    double volume[kNS], surface[kNS], aquatic[kNS], adult[kNS];
    double water_t[kNS];      // water temperature: the hour's air temperature unless the EXTENSION models it
};

struct UnitCtx {
    int64_t cell, cs;       // the thread's cell (column of the input arrays) and their row stride
    int64_t ocell, os;      // its column in the OUTPUT arrays and their row stride: the same, unless the outputs are
                            // scattered into raster cell order (ChainParams::out_col)
    int set0, H;
    bool live, write_forcing;
    double p_a;
    int n_mean_state, n_hourly_state;
};

__device__ __forceinline__ int out_slot(uint32_t mask, uint32_t class_mask, int v)
{
    return __popc(mask & class_mask & ((1u << v) - 1u));   // position among the selected variables
}

// ---- one pass over n_days days ----------------------------------------------------
// kWaterOnly : water burn-in year (R/sim_water.R:228-233): no life history, no population
// kRecord    : the final pass over the full series: period summaries / hourly series
template <int kNS, bool kFull, bool kHourly, bool kWaterOnly, bool kRecord, int kDens, bool kWT = false>
__device__ __forceinline__ void run_pass(const ChainParams &p, const SmemTables &tb, const UnitCtx &u,
                                         UnitState<kNS> &st, int n_days, int &bad)
{
    static_assert(!kWT || (kNS == 1 && kFull), "the water-temperature model runs one setting per thread, all variables");
    constexpr int kNF = kFull ? 9 : 4;
    constexpr uint32_t kState = MMS_STATE_VAR_MASK;
    // kDens packs what the launcher knows at compile time: bit 0 the density-dependence form, bits 1-2 the
    // water-body shape when all settings share one (1 cone, 2 cylinder; 0 = per setting, from the parameters)
    const int dens_type = kDens >= 0 ? (kDens & 1) : p.densmod_type;
    constexpr int kShapeAll = kDens >= 0 ? ((kDens >> 1) & 3) - 1 : -1;
    const int64_t cs = u.cs, cell = u.cell;
    const int set0 = u.set0;

    // period accumulators (aggregate_vectors(), R/pipeline_functions.R:1030-1091)
    double acc_state[kNS][5];      // adult, aquatic, area, volume, water temperature
    double acc_forc[kNF];          // mdr, efd, das0, ds [, airT, RH, wind, rain, cloud]
    int acc_wet[kNS];
    int acc_n = 0, cur_period = -1;
    auto reset_acc = [&]() {
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            acc_wet[s] = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) acc_state[s][k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < kNF; ++k) acc_forc[k] = 0.0;
        acc_n = 0;
    };
    reset_acc();

    auto flush = [&](int period) {
        if (!u.live) return;
        const int64_t pc = (int64_t)period * u.os + u.ocell;
        const double n = (double)acc_n;
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            const int vid[5] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME,
                                MMS_V_WATER_TEMPERATURE};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                if (!kFull && k >= 2) continue;
                if (p.mean_mask & (1u << vid[k])) {
                    const int64_t row = (int64_t)(set0 + s) * u.n_mean_state + out_slot(p.mean_mask, kState, vid[k]);
                    p.mean_state[row * p.P * u.os + pc] = acc_state[s][k] / n;
                }
            }
            if (p.wet_hours) p.wet_hours[(int64_t)(set0 + s) * p.P * u.os + pc] = acc_wet[s];
        }
        if (u.write_forcing) {
            const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                                MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
            for (int k = 0; k < kNF; ++k) {
                if (p.mean_mask & (1u << vid[k])) {
                    const int64_t row = out_slot(p.mean_mask, ~kState, vid[k]);
                    p.mean_forcing[row * p.P * u.os + pc] = acc_forc[k] / n;
                }
            }
        }
    };

    // day window: today, tomorrow; carried values of yesterday
    DayCtx dc;
    dc.cur = load_daily<kFull>(p, cell, 0, bad);
    dc.nxt = load_daily<kFull>(p, cell, (p.D > 1) ? 1 : 0, bad);
    dc.A = MMS_DIVC(dc.cur.tmax - dc.cur.tmin, 2.0);                       // R/sim_microclimate.R:217
    dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;                // :220
    // :271-272 (first day: its own midnight = night curve at t = 2400)
    dc.t_midnight_prev = (dc.t_sunset - dc.nxt.tmin) / p.hr.exp_e[23] + dc.nxt.tmin;
    dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;   // :522

    for (int day = 0; day < n_days; ++day) {
        bool acc_on = false;               // true while this day's hours go into the current period
        int acc_inc = 0;
        if (kRecord && p.period_of_day) {
            const int period = p.period_of_day[day];
            if (period >= 0 && period != cur_period) {
                if (cur_period >= 0) flush(cur_period);
                reset_acc();
                cur_period = period;
            }
            acc_on = period >= 0;
            acc_inc = (period >= 0) ? 1 : 0;
        }

        if (kWT) {
            // the day's clear-sky curve: the pixel's clear-sky radiation of the calendar month (a daily mean) spread
            // over the 24 hours in proportion to max(0, cos Z), cos Z = cos(lat) cos(delta) cos(h) + sin(lat) sin(delta)
            const double2 sun = p.wt.sun_day[day];
            dc.sol_a = __ldg(p.wt.cos_lat + cell) * sun.y;
            dc.sol_b = __ldg(p.wt.sin_lat + cell) * sun.x;
            double w = 0.0;
#pragma unroll 4
            for (int k = 0; k < 24; ++k) {
                const double cz = dc.sol_a * p.hr.cos_hour_angle[k] + dc.sol_b;
                w += (cz > 0.0) ? cz : 0.0;
            }
            const double clear = __ldg(p.wt.clear_sky + (int64_t)(p.wt.month_of_day[day] - 1) * cs + cell);
            dc.sol_scale = (w > 0.0) ? clear * 24.0 / w : 0.0;
        }
        const double rain_h = MMS_DIVC(dc.cur.rain, 24.0);                // R/sim_microclimate.R:132
        const double gain_m_h = MMS_DIVC(rain_h, 1000.0);                 // R/sim_water.R:122
        double gain[kNS];
#pragma unroll
        for (int s = 0; s < kNS; ++s) gain[s] = p.hab[set0 + s].catchment_area * gain_m_h;   // :123

        // the water recursion of one hour: volume, surface <- hour h (needs that hour's evaporation rate only)
        auto water_step = [&](HourF &fx, int h) {
#if !MMS_STRICT
            if (!kWT) {
                // all settings of the thread stage by stage (mms_fastmath.cuh: why)
                int shape[kNS];
                MMS_EACH(s, kNS) shape[s] = kShapeAll >= 0 ? kShapeAll : p.hab[set0 + s].shape;
                water_step_n<kNS>(p.hab + set0, st.volume, st.surface, fx.evap, gain, shape);   // R/sim_water.R:279-294
                if (kFull) MMS_EACH(s, kNS) st.water_t[s] = fx.air_t;   // water_temperature = air_temperature, R/sim_microclimate.R:163
                return;
            }
#endif
#pragma unroll
            for (int s = 0; s < kNS; ++s) {
                const HabitatDev &hb = p.hab[set0 + s];
                if (kWT) {
                    // EXTENSION: volume and temperature together, then the rates that take the water temperature
                    const double cz = dc.sol_a * p.hr.cos_hour_angle[h] + dc.sol_b;
                    const double i_clear = dc.sol_scale * ((cz > 0.0) ? cz : 0.0);
                    double tw = st.water_t[s];
                    tw = (tw != tw) ? fx.air_t : tw;             // first step: the water starts at the air's temperature
                    st.volume[s] = iterate_volume_temperature(hb, p.wt, st.volume[s], st.surface[s], tw, fx.evap, gain[s],
                                                              fx.air_t, fx.cloud, i_clear);
                    st.water_t[s] = tw;
                    if (!kWaterOnly) {
                        const int i = spline_interval(p.spline, tb.sx, tw);        // R/sim_pop_dynamics.R:108,114
                        const double dx = tw - tb.sx[i];
                        const double4 *cf = reinterpret_cast<const double4 *>(tb.coef + (size_t)i * kSplineRow);
                        const double4 c0 = cf[0], c2 = cf[2];
                        fx.mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
                        fx.das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
                        pop_hour_water(fx.ph, dens_type, fx.mdr, fx.das0);
                    }
                } else {
                    st.volume[s] = iterate_volume(hb, st.volume[s], st.surface[s], fx.evap, gain[s]);     // R/sim_water.R:279-287
                    if (kFull) st.water_t[s] = fx.air_t;         // water_temperature = air_temperature, R/sim_microclimate.R:163
                }
                st.surface[s] = habitat_surface(hb, st.volume[s], kShapeAll >= 0 ? kShapeAll : hb.shape);   // :294
            }
        };
        // the population update of hour h (volume, surface are already at hour h) + the outputs of hour h
        auto pop_and_record = [&](const HourF &f, int h) {
            if (!kWaterOnly) {
#if !MMS_STRICT
                // larval_habitat_area = water_surface_area, R/sim_pop_dynamics.R:123; all settings stage by stage
                iterate_state_n<kNS>(st.aquatic, st.adult, f.ph, dens_type, p.dd_effect, p.surface_area_cm2, st.surface);
#else
#pragma unroll
                for (int s = 0; s < kNS; ++s)   // larval_habitat_area = water_surface_area, R/sim_pop_dynamics.R:123
                    iterate_state(st.aquatic[s], st.adult[s], f.ph, dens_type, p.dd_effect,
                                  p.surface_area_cm2, st.surface[s]);
#endif
            }
            if (kRecord) {
                // branch free: days outside included_dates add 0 (their sums are never written)
#pragma unroll
                for (int s = 0; s < kNS; ++s) {
                    acc_state[s][0] += (acc_on ? st.adult[s] : 0.0);
                    acc_state[s][1] += (acc_on ? st.aquatic[s] : 0.0);
                    if (kFull) {
                        acc_state[s][2] += (acc_on ? st.surface[s] : 0.0);
                        acc_state[s][3] += (acc_on ? st.volume[s] : 0.0);
                        acc_state[s][4] += (acc_on ? st.water_t[s] : 0.0);
                    }
                    acc_wet[s] += (st.volume[s] > 0.0) ? acc_inc : 0;
                }
                acc_forc[0] += (acc_on ? f.mdr : 0.0); acc_forc[1] += (acc_on ? f.efd : 0.0);
                acc_forc[2] += (acc_on ? f.das0 : 0.0); acc_forc[3] += (acc_on ? f.ds : 0.0);
                if (kFull) {
                    acc_forc[4] += (acc_on ? f.air_t : 0.0); acc_forc[5] += (acc_on ? f.rh : 0.0); acc_forc[6] += (acc_on ? f.wind : 0.0);
                    acc_forc[7] += (acc_on ? rain_h : 0.0); acc_forc[8] += (acc_on ? f.cloud : 0.0);
                }
                acc_n += acc_inc;
                if (kHourly && u.live) {
                    const int64_t tc = (int64_t)(day * 24 + h) * u.os + u.ocell;
#pragma unroll
                    for (int s = 0; s < kNS; ++s) {
                        const double val[5] = {st.adult[s], st.aquatic[s], st.surface[s], st.volume[s], st.water_t[s]};
                        const int vid[5] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME,
                                            MMS_V_WATER_TEMPERATURE};
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (!kFull && k >= 2) continue;
                            if (p.hourly_mask & (1u << vid[k])) {
                                const int64_t row = (int64_t)(set0 + s) * u.n_hourly_state + out_slot(p.hourly_mask, kState, vid[k]);
                                p.hourly_state[row * u.H * u.os + tc] = val[k];
                            }
                        }
                        if (p.wet_flags)
                            p.wet_flags[(int64_t)(set0 + s) * u.H * u.os + tc] = (st.volume[s] > 0.0) ? 1 : 0;
                    }
                    if (u.write_forcing) {
                        const double val[9] = {f.mdr, f.efd, f.das0, f.ds, f.air_t, f.rh, f.wind, rain_h, f.cloud};
                        const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                                            MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
                        for (int k = 0; k < kNF; ++k) {
                            if (p.hourly_mask & (1u << vid[k])) {
                                const int64_t row = out_slot(p.hourly_mask, ~kState, vid[k]);
                                p.hourly_forcing[row * u.H * u.os + tc] = val[k];
                            }
                        }
                    }
                }
            }
        };

        // software pipeline over the 24 hours, three independent chains per basic block: the
        // state-independent half of hour h + 1, the water recursion of hour h + 1 (it needs only the
        // evaporation rate of that hour) and the population update of hour h (it needs the surface of
        // hour h, which the previous iteration left behind).  Same arithmetic, same results.
        HourF f = hour_forcing<kFull, kWaterOnly, kWT>(p, tb, dc, u.p_a, 0, dens_type);
        water_step(f, 0);
#pragma unroll 1
        for (int h = 0; h < 23; ++h) {
#if !MMS_STRICT && MMS_MONO_LATE_DS
            if (!kWT) {
                DsFetch dsf;
                // (measured, 100k cells x 1 setting, ms: this order 37.7; evaporation rate after the population update
                // 39.5; late part before the water step 37.7; both 38.8)
                HourF fn = hour_forcing_early<kFull, kWaterOnly>(p, dc, u.p_a, h + 1, dsf);   // requests the ds quad
                pop_and_record(f, h);
                water_step(fn, h + 1);
                hour_forcing_late<kWaterOnly>(p, tb, fn, dsf, dens_type);                     // first use of the quad
                f = fn;
                continue;
            }
#endif
            HourF fn = hour_forcing<kFull, kWaterOnly, kWT>(p, tb, dc, u.p_a, h + 1, dens_type);
            pop_and_record(f, h);
            water_step(fn, h + 1);
            f = fn;
        }
        pop_and_record(f, 23);

        // ---- roll the day window
        dc.t_midnight_prev = f.air_t;                // hour-24 value, R/sim_microclimate.R:271
        dc.rhmin_prev = dc.cur.rhmin; dc.wsmax_prev = dc.cur.wsmax; dc.ccmin_prev = dc.cur.ccmin;
        dc.cur = dc.nxt;
        const int d2 = day + 2;
        dc.nxt = load_daily<kFull>(p, cell, d2 < p.D ? d2 : p.D - 1, bad);
        dc.A = MMS_DIVC(dc.cur.tmax - dc.cur.tmin, 2.0);
        dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;
    }   // days
    if (kRecord && cur_period >= 0) flush(cur_period);
}

// ---- one work unit: 32 cells x kNS settings, all passes ---------------------------
template <int kNS, bool kFull, bool kHourly, int kDens, bool kWT = false>
__device__ __forceinline__ void simulate_unit(const ChainParams &p, const SmemTables &tb, int64_t cell, int set0,
                                              bool write_forcing)
{
    UnitCtx u;
    u.cell = cell; u.cs = p.c_stride; u.set0 = set0; u.H = 24 * p.D;
    u.live = cell < p.n_cells; u.write_forcing = write_forcing;
    u.os = p.out_stride;
    u.ocell = (p.out_col && u.live) ? __ldg(p.out_col + cell) : cell;
    u.p_a = air_pressure_kpa(__ldg(p.altitude + cell));
    u.n_mean_state = __popc(p.mean_mask & MMS_STATE_VAR_MASK);
    u.n_hourly_state = __popc(p.hourly_mask & MMS_STATE_VAR_MASK);
    int bad = 0;

    UnitState<kNS> st;
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
        st.aquatic[s] = p.initial_aquatic;               // R/sim_pop_dynamics.R:249-252
        st.adult[s] = p.initial_adult;
        st.volume[s] = p.hab[set0 + s].initial_volume;   // R/sim_water.R:277
        st.surface[s] = habitat_surface(p.hab[set0 + s], st.volume[s]);
        st.water_t[s] = __longlong_as_double(0x7ff8000000000000LL);   // "not set yet": the first step takes the air's temperature
    }
    // water burn-in years: first 8760 steps, water only (R/sim_water.R:228-233)
    for (int y = 0; y < p.burnin_years_water; ++y)
        run_pass<kNS, kFull, kHourly, true, false, kDens, kWT>(p, tb, u, st, kDaysPerYear, bad);
    // every later pass re-runs the water model from this volume [and temperature]: the reference computes the
    // water series once and the population burn-in re-indexes its first year
    double vol0[kNS], tw0[kNS];
#pragma unroll
    for (int s = 0; s < kNS; ++s) { vol0[s] = st.volume[s]; tw0[s] = st.water_t[s]; }
    // population burn-in years: first 8760 steps (R/sim_pop_dynamics.R:221-226)
    for (int y = 0; y < p.burnin_years_pop; ++y) {
        run_pass<kNS, kFull, kHourly, false, false, kDens, kWT>(p, tb, u, st, kDaysPerYear, bad);
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            st.volume[s] = vol0[s];
            st.surface[s] = habitat_surface(p.hab[set0 + s], vol0[s]);
            st.water_t[s] = tw0[s];
        }
    }
    // the full series, recorded
    run_pass<kNS, kFull, kHourly, false, true, kDens, kWT>(p, tb, u, st, p.D, bad);

    if (bad && u.live) atomicOr(p.error_flag, 1);
}

// ---- the kernel -----------------------------------------------------------
// kNS     : habitat settings integrated per thread (they share all forcing-derived work)
// kFull   : false = only the reference's six vector_variables can be summarised / written;
//           true  = all MMS_NVARS (adds water body and microclimate variables, cloud cover)
// kHourly : write full hourly series
// kMaxWarps: largest CTA this instantiation may be launched with (register budget = 64K / threads)
// kDens: what the launcher knows at compile time (bit 0: 0 logit / 1 cox_ph; bits 1-2: common shape + 1), else -1
// kWT: the water-temperature EXTENSION (one setting per thread, all variables)
template <int kNS, bool kFull, bool kHourly, int kMaxWarps, int kDens, bool kWT = false>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
chain_kernel(const __grid_constant__ ChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_coef = reinterpret_cast<double *>(smem_raw);              // [n][3][4]
    double *s_x = s_coef + (size_t)p.spline.n * kSplineRow;                     // [n]
    double *s_gx = s_x + p.spline.n;                                    // [nx]
    double *s_gy = s_gx + p.ds.nx;                                      // [ny]
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[(k / 12) * kSplineRow + k % 12] = p.spline.coef[k];
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.ds.nx; k += blockDim.x) s_gx[k] = p.ds.x[k];
    for (int k = threadIdx.x; k < p.ds.ny; k += blockDim.x) s_gy[k] = p.ds.y[k];
    ::mms::fm::math_tables_init();                                      // exp / log tables of the fast math
    __syncthreads();
    SmemTables tb{s_x, s_coef, s_gx, s_gy};

    const int lane = threadIdx.x & 31;
    for (;;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(p.unit_counter, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= p.n_units) break;
        const int cb = unit % p.n_cell_blocks, g = unit / p.n_cell_blocks;
        simulate_unit<kNS, kFull, kHourly, kDens, kWT>(p, tb, (int64_t)cb * 32 + lane, g * kNS, g == 0);
    }
}

}  // namespace MMS_MODE
}  // namespace mms
