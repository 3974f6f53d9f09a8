// chain_kernel.cuh -- the fused per-cell chain: monthly forcing -> daily
// (approx) -> hourly (SINEC / VSINE) -> evaporation -> water body ->
// life-history parameters -> population -> period summaries, one thread per
// cell, the whole hourly time loop inside the kernel, state in registers.
//
// Included by chain_fast.cu (MMS_STRICT=0) and chain_strict.cu (MMS_STRICT=1).
// Citations are relative to the reference root.
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
    const DayEntry de = p.days[day];                     // uniform across the block
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

// ---- life-history tables ------------------------------------------------
// spline_eval() (base R splines.c) interval: largest i with x[i] <= u, 0 below the
// first knot, n-1 at or beyond the last one.
__device__ __forceinline__ int spline_interval(const SplineDev &sp, const double *__restrict__ xs, double u)
{
    const int n = sp.n;
    double g = (u - sp.x0) * sp.inv_step;
    g = fmin(fmax(g, 0.0), (double)(n - 1));             // NaN -> 0
    int i = (int)g;
    while (i > 0 && u < xs[i]) --i;
    while (i < n - 1 && !(u < xs[i + 1])) ++i;
    return i;
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

__device__ __forceinline__ double ds_surface(const SurfaceDev &g, double temperature, double humidity)
{
    double lx = approx_index(g.x, g.nx, g.x0, g.inv_dx, temperature);
    double ly = approx_index(g.y, g.ny, g.y0, g.inv_dy, humidity);
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
}

// ---- the kernel -----------------------------------------------------------
// kNS     : habitat settings integrated per thread (they share all forcing-derived work)
// kFull   : false = only the reference's six vector_variables can be summarised / written;
//           true  = all MMS_NVARS (adds water body and microclimate variables, cloud cover)
// kHourly : write full hourly series
template <int kNS, bool kFull, bool kHourly>
__global__ void __launch_bounds__(kBlock)
chain_kernel(const __grid_constant__ ChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_x = reinterpret_cast<double *>(smem_raw);                 // [n]
    double *s_coef = s_x + ((p.spline.n + 1) & ~1);                     // [n][3][4]
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[k] = p.spline.coef[k];
    __syncthreads();

    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // < c_stride (padded)
    const int set0 = blockIdx.y * kNS;
    const bool live = cell < p.n_cells;
    const int64_t cs = p.c_stride;
    const int H = 24 * p.D;

    const double p_a = air_pressure_kpa(__ldg(p.altitude + cell));
    int bad = 0;

    // output slot of each variable (position among the selected ones)
    auto slot = [](uint32_t mask, uint32_t class_mask, int v) {
        return __popc(mask & class_mask & ((1u << v) - 1u));
    };
    const uint32_t kState = MMS_STATE_VAR_MASK;
    const int n_mean_state = __popc(p.mean_mask & kState);
    const int n_hourly_state = __popc(p.hourly_mask & kState);

    double volume[kNS], surface[kNS], aquatic[kNS], adult[kNS];
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
        aquatic[s] = p.initial_aquatic;                  // R/sim_pop_dynamics.R:249-252
        adult[s] = p.initial_adult;
        volume[s] = p.hab[set0 + s].initial_volume;      // R/sim_water.R:277
    }

    // passes: water burn-in years (water only), population burn-in years (first
    // 8760 steps, R/sim_pop_dynamics.R:221-226), then the full series
    const int n_pass = p.burnin_years_water + p.burnin_years_pop + 1;
    double vol_after_water_burnin[kNS];
#pragma unroll
    for (int s = 0; s < kNS; ++s) vol_after_water_burnin[s] = volume[s];

    for (int pass = 0; pass < n_pass; ++pass) {
        const bool water_only = pass < p.burnin_years_water;
        const bool final_pass = pass == n_pass - 1;
        const int n_days = final_pass ? p.D : kDaysPerYear;
        if (pass == p.burnin_years_water) {
#pragma unroll
            for (int s = 0; s < kNS; ++s) vol_after_water_burnin[s] = volume[s];
        }
        if (!water_only) {
            // each population pass re-runs the water model from its post-burn-in volume:
            // the reference computes the water series once and re-indexes its first year
#pragma unroll
            for (int s = 0; s < kNS; ++s) volume[s] = vol_after_water_burnin[s];
        }
#pragma unroll
        for (int s = 0; s < kNS; ++s) surface[s] = habitat_surface(p.hab[set0 + s], volume[s]);

        // period accumulators
        double acc_state[kNS][4];      // adult, aquatic, area, volume
        double acc_forc[kFull ? 9 : 4];// mdr, efd, das0, ds [, airT, RH, wind, rain, cloud]
        int acc_wet[kNS];
        int acc_n = 0, cur_period = -1;
#pragma unroll
        for (int s = 0; s < kNS; ++s) {
            acc_wet[s] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc_state[s][k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < (kFull ? 9 : 4); ++k) acc_forc[k] = 0.0;

        auto flush = [&](int period) {
            if (!live) return;
            const double inv_n = 1.0 / (double)acc_n;
            (void)inv_n;
            const int64_t pc = (int64_t)period * cs + cell;
#pragma unroll
            for (int s = 0; s < kNS; ++s) {
                const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!kFull && k >= 2) continue;
                    if (p.mean_mask & (1u << vid[k])) {
                        const int64_t row = (int64_t)(set0 + s) * n_mean_state + slot(p.mean_mask, kState, vid[k]);
                        p.mean_state[row * p.P * cs + pc] = acc_state[s][k] / (double)acc_n;
                    }
                }
                if (p.wet_hours) p.wet_hours[(int64_t)(set0 + s) * p.P * cs + pc] = acc_wet[s];
            }
            if (blockIdx.y == 0) {
                const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                                    MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
                for (int k = 0; k < (kFull ? 9 : 4); ++k) {
                    if (p.mean_mask & (1u << vid[k])) {
                        const int64_t row = slot(p.mean_mask, ~kState, vid[k]);
                        p.mean_forcing[row * p.P * cs + pc] = acc_forc[k] / (double)acc_n;
                    }
                }
            }
        };

        // day window: today, tomorrow; carried values of yesterday
        Daily cur = load_daily<kFull>(p, cell, 0, bad);
        Daily nxt = load_daily<kFull>(p, cell, (p.D > 1) ? 1 : 0, bad);
        double A = MMS_DIVC(cur.tmax - cur.tmin, 2.0);                        // R/sim_microclimate.R:217
        double t_sunset = A * p.hr.z_night + cur.tmin + A;                    // :220
        double t_sunset_prev = t_sunset;                                      // :223 (first day: its own)
        // :271-272 (first day: its own midnight = night curve at t = 2400)
        double t_midnight_prev = (t_sunset - nxt.tmin) / p.hr.exp_e[23] + nxt.tmin;
        double rhmin_prev = cur.rhmin, wsmax_prev = cur.wsmax, ccmin_prev = cur.ccmin;   // :522

        for (int day = 0; day < n_days; ++day) {
            const bool record = final_pass;
            int period = -1;
            if (record && p.period_of_day) {
                period = p.period_of_day[day];
                if (period >= 0 && period != cur_period) {
                    if (cur_period >= 0) flush(cur_period);
#pragma unroll
                    for (int s = 0; s < kNS; ++s) {
                        acc_wet[s] = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc_state[s][k] = 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < (kFull ? 9 : 4); ++k) acc_forc[k] = 0.0;
                    acc_n = 0;
                    cur_period = period;
                }
            }
            const bool accumulate = period >= 0;

            const double rain_h = MMS_DIVC(cur.rain, 24.0);                   // R/sim_microclimate.R:132
            const double gain_m_h = MMS_DIVC(rain_h, 1000.0);                 // R/sim_water.R:122
            double gain[kNS];
#pragma unroll
            for (int s = 0; s < kNS; ++s) gain[s] = p.hab[set0 + s].catchment_area * gain_m_h;   // :123

            double air_t = 0.0;
#pragma unroll 1
            for (int h = 0; h < 24; ++h) {       // hour = h + 1, times = 100 * (h + 1)
                // ---- SINEC air temperature, R/sim_microclimate.R:185-285
                if (h < 5) {                     // times < time_sunrise: pre-dawn ramp, :275-277
                    air_t = cur.tmin + (t_midnight_prev - cur.tmin) * p.hr.predawn[h];
                } else if (h <= 17) {            // 600 <= times <= 1800: sine, :255-256
                    air_t = A * p.hr.z_light[h] + cur.tmin + A;
                } else {                         // night: exponential decay, :260-262
                    air_t = (t_sunset - nxt.tmin) / p.hr.exp_e[h] + nxt.tmin;
                }
                // ---- VSINE humidity / wind speed [/ cloud cover], :466-566
                const bool before_sunrise = h < 5;          // times < 6
                const bool after_inflection = h > 12;       // times > 13
                double inf_last = before_sunrise ? rhmin_prev : cur.rhmin;       // :529-530
                double sun_next = after_inflection ? nxt.rhmax : cur.rhmax;      // :548-549
                const double rh = inf_last + p.hr.w_vsine[h] * (sun_next - inf_last);       // :556-557
                inf_last = before_sunrise ? wsmax_prev : cur.wsmax;
                sun_next = after_inflection ? nxt.wsmin : cur.wsmin;
                const double wind = sun_next + p.hr.w_vsine_inv[h] * (inf_last - sun_next); // :553-554
                double cloud = 0.0;
                if (kFull) {
                    inf_last = before_sunrise ? ccmin_prev : cur.ccmin;
                    sun_next = after_inflection ? nxt.ccmax : cur.ccmax;
                    cloud = inf_last + p.hr.w_vsine[h] * (sun_next - inf_last);
                }

                // ---- evaporation rate per unit area, R/sim_water.R:264-270 (air temperature)
                const double evap = evaporation_rate(air_t, rh, p_a, wind);

                // ---- life-history parameters, R/sim_pop_dynamics.R:100-135
                // water_temperature = air_temperature, R/sim_microclimate.R:163
                double mdr = 0.0, efd = 0.0, das0 = 0.0, ds = 0.0;
                PopHour ph;
                if (!water_only) {
                    const int i = spline_interval(p.spline, s_x, air_t);
                    const double dx = air_t - s_x[i];
                    const double4 *cf = reinterpret_cast<const double4 *>(s_coef + (size_t)i * 12);
                    const double4 c0 = cf[0], c1 = cf[1], c2 = cf[2];
                    mdr = ensure_positive(c0.x + dx * (c0.y + dx * (c0.z + dx * c0.w)));
                    efd = ensure_positive(c1.x + dx * (c1.y + dx * (c1.z + dx * c1.w)));
                    das0 = ensure_positive(c2.x + dx * (c2.y + dx * (c2.z + dx * c2.w)));
                    ds = ds_surface(p.ds, air_t, rh);
                    ph = pop_hour(p.densmod_type, mdr, efd, das0, ds);
                }

                // ---- water body and population, per habitat setting
#pragma unroll
                for (int s = 0; s < kNS; ++s) {
                    const HabitatDev &hb = p.hab[set0 + s];
                    volume[s] = iterate_volume(hb, volume[s], surface[s], evap, gain[s]);   // R/sim_water.R:279-287
                    surface[s] = habitat_surface(hb, volume[s]);                            // :294
                    if (!water_only)   // larval_habitat_area = water_surface_area, R/sim_pop_dynamics.R:123
                        iterate_state(aquatic[s], adult[s], ph, p.densmod_type, p.dd_effect,
                                      p.surface_area_cm2, surface[s]);
                }

                // ---- outputs
                if (record) {
                    if (accumulate) {
#pragma unroll
                        for (int s = 0; s < kNS; ++s) {
                            acc_state[s][0] += adult[s];
                            acc_state[s][1] += aquatic[s];
                            if (kFull) {
                                acc_state[s][2] += surface[s];
                                acc_state[s][3] += volume[s];
                            }
                            acc_wet[s] += (volume[s] > 0.0) ? 1 : 0;
                        }
                        acc_forc[0] += mdr; acc_forc[1] += efd; acc_forc[2] += das0; acc_forc[3] += ds;
                        if (kFull) {
                            acc_forc[4] += air_t; acc_forc[5] += rh; acc_forc[6] += wind;
                            acc_forc[7] += rain_h; acc_forc[8] += cloud;
                        }
                        ++acc_n;
                    }
                    if (kHourly && live) {
                        const int64_t tc = (int64_t)(day * 24 + h) * cs + cell;
#pragma unroll
                        for (int s = 0; s < kNS; ++s) {
                            const double val[4] = {adult[s], aquatic[s], surface[s], volume[s]};
                            const int vid[4] = {MMS_V_ADULT, MMS_V_AQUATIC, MMS_V_LARVAL_HABITAT_AREA, MMS_V_WATER_VOLUME};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (!kFull && k >= 2) continue;
                                if (p.hourly_mask & (1u << vid[k])) {
                                    const int64_t row = (int64_t)(set0 + s) * n_hourly_state + slot(p.hourly_mask, kState, vid[k]);
                                    p.hourly_state[row * H * cs + tc] = val[k];
                                }
                            }
                            if (p.wet_flags)
                                p.wet_flags[(int64_t)(set0 + s) * H * cs + tc] = (volume[s] > 0.0) ? 1 : 0;
                        }
                        if (blockIdx.y == 0) {
                            const double val[9] = {mdr, efd, das0, ds, air_t, rh, wind, rain_h, cloud};
                            const int vid[9] = {MMS_V_MDR, MMS_V_EFD, MMS_V_DAS_ZERODENSITY, MMS_V_DS, MMS_V_AIR_TEMPERATURE,
                                                MMS_V_HUMIDITY, MMS_V_WINDSPEED, MMS_V_RAINFALL, MMS_V_CLOUDCOVER};
#pragma unroll
                            for (int k = 0; k < (kFull ? 9 : 4); ++k) {
                                if (p.hourly_mask & (1u << vid[k])) {
                                    const int64_t row = slot(p.hourly_mask, ~kState, vid[k]);
                                    p.hourly_forcing[row * H * cs + tc] = val[k];
                                }
                            }
                        }
                    }
                }
            }   // hours

            // ---- roll the day window
            t_midnight_prev = air_t;                     // hour-24 value, R/sim_microclimate.R:271
            t_sunset_prev = t_sunset;
            rhmin_prev = cur.rhmin; wsmax_prev = cur.wsmax; ccmin_prev = cur.ccmin;
            cur = nxt;
            const int d2 = day + 2;
            nxt = load_daily<kFull>(p, cell, d2 < p.D ? d2 : p.D - 1, bad);
            A = MMS_DIVC(cur.tmax - cur.tmin, 2.0);
            t_sunset = A * p.hr.z_night + cur.tmin + A;
        }   // days
        (void)t_sunset_prev;
        if (final_pass && cur_period >= 0) flush(cur_period);
    }   // passes

    if (bad && live) atomicOr(p.error_flag, 1);
}

}  // namespace MMS_MODE
}  // namespace mms
