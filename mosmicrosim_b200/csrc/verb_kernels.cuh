// verb_kernels.cuh -- stage-level twins of the three state-free pipeline verbs, for callers that
// want the intermediate products the reference materialises between its verbs:
//   interpolate_daily_climate()     R/pipeline_functions.R:258-275  -> daily_kernel
//   simulate_hourly_microclimate()  R/pipeline_functions.R:284-301  -> hourly_kernel
//   simulate_hourly_lifehistory()   R/pipeline_functions.R:634-683  -> lifehistory_kernel
// (the fused chain computes the same quantities on the fly and never stores them).
// All three are element-wise in the cell index: HBM-bound streaming kernels, layout
// [variable][time][cell] with the cell index fastest (coalesced loads and stores).
// They share the device functions of the fused kernel, so the arithmetic is identical.
#pragma once

#include "chain_kernel.cuh"

namespace mms {
namespace MMS_MODE {

// daily_from_monthly_climate(), R/sim_microclimate.R:7-64: out[9][D][cs]
// grid (cells / 128, D); p.monthly rows hold log1p(rainfall) already
__global__ void __launch_bounds__(128) daily_kernel(const __grid_constant__ ChainParams p, double *__restrict__ out)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int day = blockIdx.y;
    if (cell >= p.c_stride) return;
    int bad = 0;
    const Daily d = load_daily<true>(p, cell, day, bad);
    const int64_t plane = (int64_t)p.D * p.c_stride, at = (int64_t)day * p.c_stride + cell;
    out[MMS_F_TMIN * plane + at] = d.tmin;
    out[MMS_F_TMAX * plane + at] = d.tmax;
    out[MMS_F_RHMIN * plane + at] = d.rhmin;
    out[MMS_F_RHMAX * plane + at] = d.rhmax;
    out[MMS_F_WSMIN * plane + at] = d.wsmin;
    out[MMS_F_WSMAX * plane + at] = d.wsmax;
    out[MMS_F_CCMIN * plane + at] = d.ccmin;
    out[MMS_F_CCMAX * plane + at] = d.ccmax;
    out[MMS_F_RAINFALL_DAILY * plane + at] = d.rain;
    if (bad && cell < p.n_cells) atomicOr(p.error_flag, 1);
}

__device__ __forceinline__ Daily read_daily(const double *__restrict__ daily, int64_t plane, int64_t at)
{
    Daily d;
    d.tmin = daily[MMS_F_TMIN * plane + at];
    d.tmax = daily[MMS_F_TMAX * plane + at];
    d.rhmin = daily[MMS_F_RHMIN * plane + at];
    d.rhmax = daily[MMS_F_RHMAX * plane + at];
    d.wsmin = daily[MMS_F_WSMIN * plane + at];
    d.wsmax = daily[MMS_F_WSMAX * plane + at];
    d.ccmin = daily[MMS_F_CCMIN * plane + at];
    d.ccmax = daily[MMS_F_CCMAX * plane + at];
    d.rain = daily[MMS_F_RAINFALL_DAILY * plane + at];
    return d;
}

// hourly_from_daily_climate_ambient(), R/sim_microclimate.R:126-173: daily[9][D][cs] ->
// out[6][24 D][cs] = air_temperature, humidity, cloudcover, windspeed, water_temperature, rainfall.
// Every (cell, day) is independent: yesterday's midnight temperature only needs yesterday's and
// today's daily values (R/sim_microclimate.R:271-275).  grid (cells / 128, D).
__global__ void __launch_bounds__(128) hourly_kernel(const __grid_constant__ ChainParams p,
                                                     const double *__restrict__ daily, double *__restrict__ out)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int day = blockIdx.y;
    if (cell >= p.c_stride) return;
    const int64_t cs = p.c_stride, plane = (int64_t)p.D * cs;
    const int dprev = day > 0 ? day - 1 : 0, dnext = day + 1 < p.D ? day + 1 : p.D - 1;
    const Daily prv = read_daily(daily, plane, (int64_t)dprev * cs + cell);
    DayCtx dc;
    dc.cur = read_daily(daily, plane, (int64_t)day * cs + cell);
    dc.nxt = read_daily(daily, plane, (int64_t)dnext * cs + cell);
    dc.A = MMS_DIVC(dc.cur.tmax - dc.cur.tmin, 2.0);                       // R/sim_microclimate.R:217
    dc.t_sunset = dc.A * p.hr.z_night + dc.cur.tmin + dc.A;                // :220
    // yesterday's hour-24 value: night curve from yesterday's sunset towards today's minimum;
    // the first day repeats its own (:271-272), i.e. towards tomorrow's minimum
    const double a_prev = MMS_DIVC(prv.tmax - prv.tmin, 2.0);
    const double sunset_prev = a_prev * p.hr.z_night + prv.tmin + a_prev;
    const double tmin_after_prev = day > 0 ? dc.cur.tmin : dc.nxt.tmin;
    dc.t_midnight_prev = (sunset_prev - tmin_after_prev) / p.hr.exp_e[23] + tmin_after_prev;
    dc.rhmin_prev = prv.rhmin; dc.wsmax_prev = prv.wsmax; dc.ccmin_prev = prv.ccmin;   // :522
    const double rain_h = MMS_DIVC(dc.cur.rain, 24.0);                     // :132
    const int64_t hplane = (int64_t)24 * p.D * cs;
#pragma unroll 4
    for (int h = 0; h < 24; ++h) {
        HourF f;
        hour_microclimate<true>(p.hr, dc, h, f);
        const int64_t at = ((int64_t)day * 24 + h) * cs + cell;
        out[0 * hplane + at] = f.air_t;
        out[1 * hplane + at] = f.rh;
        out[2 * hplane + at] = f.cloud;
        out[3 * hplane + at] = f.wind;
        out[4 * hplane + at] = f.air_t;                                    // water_temperature = air_temperature, :163
        out[5 * hplane + at] = rain_h;
    }
}

// ---- process_terraclimate_pixel_vars(), R/terraclimate.R:581-643 -----------------
// NicheMapR::VAPPRS as restated in oracle/mms_oracle.c (third party, parity unpinned): Goff-Gratch
// saturation vapour pressure in Pa, over ice at or below the triple point and over water above it.
// CUDA libm log10 / pow in both math modes: this runs once per monthly knot, not per hour.
__device__ __forceinline__ double vapprs_pa(double db)
{
    const double t = db + 273.16;
    double loge;
    if (t <= 273.16) {
        loge = -9.09718 * (273.16 / t - 1.) - 3.56654 * log10(273.16 / t) +
               .876793 * (1. - MMS_DIVC(t, 273.16)) + 0.7858350313586662;           // + log10(6.1071)
    } else {
        loge = -7.90298 * (373.16 / t - 1.) + 5.02808 * log10(373.16 / t) -
               1.3816E-07 * (pow(10., 11.344 * (1. - MMS_DIVC(t, 373.16))) - 1.) +
               8.1328E-03 * (pow(10., -3.49149 * (373.16 / t - 1.)) - 1.) + 3.0057148979490314;   // + log10(1013.246)
    }
    return pow(10., loge) * 100;
}

// relative_humidity(), R/utils.R:80-89 (pmin / pmax keep NaN)
__device__ __forceinline__ double relative_humidity_pct(double temperature, double vapour_pressure)
{
    double rh = vapour_pressure / vapprs_pa(temperature) * 100;
    rh = rh > 100 ? 100 : rh;
    rh = rh < 0.01 ? 0.01 : rh;
    return rh;
}

// cloud_cover(), R/utils.R:141-147
__device__ __forceinline__ double cloud_cover_pct(double srad, double clear_sky, double multiplier)
{
    double cloud = (1 - srad / clear_sky) * 100;
    cloud = cloud * multiplier;
    cloud = cloud > 100 ? 100 : cloud;
    cloud = cloud < 0 ? 0 : cloud;
    return cloud;
}

// grid (ceil(n_cols / 128), M).  Padding cells (>= n_cells) get zeros.
__global__ void __launch_bounds__(128) terraclimate_kernel(const __grid_constant__ TerraParams p)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (cell >= p.n_cols) return;
    const int64_t cs = p.c_stride, plane = (int64_t)p.M * cs, at = (int64_t)m * cs + cell;
    double o[MMS_NFORCING] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cell < p.n_cells) {
        const double tmax = p.terraclimate[0 * plane + at], tmin = p.terraclimate[1 * plane + at];
        const double ppt = p.terraclimate[2 * plane + at], ws = p.terraclimate[3 * plane + at];
        const double vpd = p.terraclimate[4 * plane + at], srad = p.terraclimate[5 * plane + at];
        const double clear = p.clear_sky[(int64_t)(p.month_id[m] - 1) * cs + cell];      // R/terraclimate.R:600-603
        const double tmean = MMS_DIVC(tmax + tmin, 2.0);                                  // :607
        // vapour_pressure(), R/utils.R:64-78: WETAIR(db, rh = 100)$e = esat * rh / 100, minus vpd in Pa
        const double vp = MMS_DIVC(vapprs_pa(tmean) * 100, 100.0) - vpd * 1000;
        o[MMS_F_TMIN] = tmin;
        o[MMS_F_TMAX] = tmax;
        o[MMS_F_RHMAX] = relative_humidity_pct(tmin, vp);                                 // :609
        o[MMS_F_RHMIN] = relative_humidity_pct(tmax, vp);                                 // :610
        o[MMS_F_WSMAX] = ws * p.wind_factor;                                              // :615-617
        o[MMS_F_WSMIN] = o[MMS_F_WSMAX] * 0.1;                                            // :618
        o[MMS_F_CCMAX] = cloud_cover_pct(srad, clear, 1.2);                               // :624
        o[MMS_F_CCMIN] = cloud_cover_pct(srad, clear, 0.8);                               // :625
        const double rain = ppt / p.n_days[m];                                            // :630
        o[MMS_F_RAINFALL_DAILY] = p.log1p_rainfall ? log1p(rain) : rain;                  // R/sim_microclimate.R:33
    }
#pragma unroll
    for (int v = 0; v < MMS_NFORCING; ++v) p.monthly[v * plane + at] = o[v];
}

// ---- summarise_vectors() -> aggregate_vectors(), R/pipeline_functions.R:995-1091 ---------------------------
// Period means of hourly series: hourly [n_vars][24 D][cs] -> means [n_vars][P][cs].  One thread per (cell,
// variable) walks its series once (a warp reads 256 contiguous bytes per hour: an HBM-bound streaming kernel) and
// accumulates hour by hour in the order of the fused kernel, so both give the same bits.  period_of_day: -1 = day
// not in included_dates; non-negative entries non-decreasing (checked on the host).  Periods without a day keep
// the NaN the output was filled with.  grid (cs / 128, n_vars).
__global__ void __launch_bounds__(128) summarise_kernel(const double *__restrict__ hourly, const int32_t *__restrict__ period_of_day,
                                                        int64_t cs, int32_t D, int32_t P, double *__restrict__ means)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cs) return;
    const double *src = hourly + (int64_t)blockIdx.y * 24 * D * cs + cell;
    double *dst = means + (int64_t)blockIdx.y * P * cs + cell;
    double acc = 0.0;
    int n = 0, cur = -1;
    for (int day = 0; day < D; ++day) {
        const int period = period_of_day[day];
        if (period < 0) continue;                                          // filter(date %in% included_dates), :1052-1054
        if (period != cur) {
            if (cur >= 0) dst[(int64_t)cur * cs] = acc / (double)n;        // mean by period, :1055-1064 / :1070-1086
            cur = period; acc = 0.0; n = 0;
        }
        const double *row = src + (int64_t)day * 24 * cs;
#pragma unroll 8
        for (int h = 0; h < 24; ++h) acc += row[(int64_t)h * cs];
        n += 24;
    }
    if (cur >= 0) dst[(int64_t)cur * cs] = acc / (double)n;
}

// simulate_vector_lifehistory(), R/sim_pop_dynamics.R:100-135: element-wise over n values;
// out mdr, efd, das_zerodensity, ds.  Spline rows and ds axes in shared memory like the fused kernel.
__global__ void __launch_bounds__(512, 2) lifehistory_kernel(const __grid_constant__ ChainParams p, int64_t n,
                                                          const double *__restrict__ water_t,
                                                          const double *__restrict__ air_t,
                                                          const double *__restrict__ humidity,
                                                          double *__restrict__ mdr, double *__restrict__ efd,
                                                          double *__restrict__ das0, double *__restrict__ ds)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_coef = reinterpret_cast<double *>(smem_raw);              // [n][3][4]
    double *s_x = s_coef + (size_t)p.spline.n * kSplineRow;
    double *s_gx = s_x + p.spline.n;
    double *s_gy = s_gx + p.ds.nx;
    for (int k = threadIdx.x; k < p.spline.n * 12; k += blockDim.x) s_coef[(k / 12) * kSplineRow + k % 12] = p.spline.coef[k];
    for (int k = threadIdx.x; k < p.spline.n; k += blockDim.x) s_x[k] = p.spline.x[k];
    for (int k = threadIdx.x; k < p.ds.nx; k += blockDim.x) s_gx[k] = p.ds.x[k];
    for (int k = threadIdx.x; k < p.ds.ny; k += blockDim.x) s_gy[k] = p.ds.y[k];
    __syncthreads();
    const SmemTables tb{s_x, s_coef, s_gx, s_gy};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double tw = water_t[i], ta = air_t[i];
        // mdr and das take the water temperature, efd the air temperature (R/sim_pop_dynamics.R:108-114)
        const int iw = spline_interval(p.spline, tb.sx, tw), ia = spline_interval(p.spline, tb.sx, ta);
        const double dw = tw - tb.sx[iw], da = ta - tb.sx[ia];
        const double *cw = tb.coef + (size_t)iw * kSplineRow, *ca = tb.coef + (size_t)ia * kSplineRow;
        mdr[i] = ensure_positive(cw[0] + dw * (cw[1] + dw * (cw[2] + dw * cw[3])));
        efd[i] = ensure_positive(ca[4] + da * (ca[5] + da * (ca[6] + da * ca[7])));
        das0[i] = ensure_positive(cw[8] + dw * (cw[9] + dw * (cw[10] + dw * cw[11])));
        ds[i] = ds_surface(p.ds, tb, ta, humidity[i]);
    }
}

}  // namespace MMS_MODE
}  // namespace mms
