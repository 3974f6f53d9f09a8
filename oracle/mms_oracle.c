/*
 * mms_oracle.c -- CPU restatement of the mosmicrosim per-cell simulation chain.
 *
 * TEST INFRASTRUCTURE ONLY (see mms_oracle.h).  PARITY UNPINNED: no R, no
 * reference golden vectors, no fitted tables are available; this is a literal
 * restatement, cross-checked against oracle/restate_numpy.py and the
 * known-answer constants of the R source.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile) so that
 * every expression is evaluated in IEEE double in the order R evaluates it.
 * Citations are relative to /root/reference.
 */
#include "mms_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ========================================================================
 * S: monthly -> daily
 * ===================================================================== */

/* base-R stats::approx, C routine approx1() (R sources src/library/stats/src/
 * approx.c; NOT under /root/reference, restated from the published source),
 * method = "linear", rule = 1 (NA outside the knots), as called at
 * R/spline_interpolation.R:25-27 and inside fields::interp.surface. */
double mmso_approx1(double v, const double *x, const double *y, long n)
{
    if (!n) return NAN;
    long i = 0, j = n - 1;
    if (v < x[i]) return NAN;            /* rule = 1: ylow  = NA */
    if (v > x[j]) return NAN;            /*           yhigh = NA */
    while (i < j - 1) {                  /* bisection: x[i] <= v <= x[j] */
        long ij = (i + j) / 2;
        if (v < x[ij]) j = ij; else i = ij;
    }
    if (v == x[j]) return y[j];
    if (v == x[i]) return y[i];
    return y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]));
}

/* spline_seasonal(), in-range branch only: R/spline_interpolation.R:6-27.
 * Returns -1 when any(extrapolation) (R/spline_interpolation.R:30-34) -- that
 * branch fits an mgcv::gam and is not restated (third party, unpinned). */
int mmso_spline_seasonal(const double *values, const double *datenum, int M,
                         const double *datenum_predict, int D, double *out)
{
    double lo = datenum[0], hi = datenum[0];
    for (int m = 1; m < M; ++m) {
        if (datenum[m] < lo) lo = datenum[m];
        if (datenum[m] > hi) hi = datenum[m];
    }
    for (int d = 0; d < D; ++d)
        if (datenum_predict[d] > hi || datenum_predict[d] < lo) return -1;
    /* stats::approx(x, y, xout) has na.rm = TRUE: regularize.values() drops the (x, y) pairs with an NA before
     * approx1() sees them (R sources src/library/stats/R/approx.R; restated from the published source), and
     * approx() stops with "need at least two non-NA values to interpolate" below two pairs (here: -3).  Prediction
     * dates outside the knots that are left get NA (rule = 1) -- not the GAM: `extrapolation` above looks at the
     * dates only. */
    double *xk = (double *)malloc(sizeof(double) * (size_t)M), *yk = (double *)malloc(sizeof(double) * (size_t)M);
    if (!xk || !yk) { free(xk); free(yk); return -2; }
    long n = 0;
    for (int m = 0; m < M; ++m)
        if (!isnan(values[m]) && !isnan(datenum[m])) { xk[n] = datenum[m]; yk[n] = values[m]; ++n; }
    if (n < 2) { free(xk); free(yk); return -3; }
    for (int d = 0; d < D; ++d)
        out[d] = mmso_approx1(datenum_predict[d], xk, yk, n);
    free(xk);
    free(yk);
    return 0;
}

/* daily_from_monthly_climate(): R/sim_microclimate.R:7-64.  monthly is
 * [9][M] for one cell (variable order MMSO_*), daily is [9][D].  mid_day is
 * as.numeric(mid_date - min(mid_date)), day_predict likewise for
 * dates_predict (R/spline_interpolation.R:15,20). */
int mmso_daily_from_monthly(const double *monthly, const double *mid_day, int M,
                            const double *day_predict, int D, double *daily)
{
    double *tmp = (double *)malloc(sizeof(double) * (size_t)M);
    if (!tmp) return -2;
    for (int v = 0; v < MMSO_NFORCING; ++v) {
        const double *vals = monthly + (size_t)v * M;
        if (v == MMSO_RAIN) {            /* :33  rainfall_daily = log1p(rainfall_daily) */
            for (int m = 0; m < M; ++m) tmp[m] = log1p(vals[m]);
            vals = tmp;
        }
        int rc = mmso_spline_seasonal(vals, mid_day, M, day_predict, D,
                                      daily + (size_t)v * D);
        if (rc) { free(tmp); return rc; }
    }
    /* :61  rainfall_daily = pmax(0, expm1(rainfall_daily)) */
    double *r = daily + (size_t)MMSO_RAIN * D;
    for (int d = 0; d < D; ++d) {
        double e = expm1(r[d]);
        r[d] = isnan(e) ? e : ((0.0 < e) ? e : 0.0);
    }
    free(tmp);
    return 0;
}

/* ========================================================================
 * M: daily -> hourly
 * ===================================================================== */

/* interpolate_daily_air_temp() (SINEC): R/sim_microclimate.R:185-285, with
 * the default time_sunrise = 6, time_sunset = 18, time_daily_max = 13. */
void mmso_interpolate_daily_air_temp(const double *temp_max_day, const double *temp_min_day,
                                     int n_days, double *temp_hour)
{
    double time_sunrise = 6, time_sunset = 18, time_daily_max = 13;
    time_sunrise = time_sunrise * 100;                                   /* :195 */
    time_sunset = time_sunset * 100;
    time_daily_max = time_daily_max * 100;
    const double time_reference = (time_daily_max - time_sunrise) / 2 + time_sunrise; /* :204 */
    const double tau = 3 / (2400 - time_sunset + time_sunrise);                       /* :207 */
    const double Z_night = sin((360 * (time_sunset - time_reference) /
                                (2 * (time_daily_max - time_sunrise))) / 57.29577);   /* :210 */

    double *A_day = (double *)malloc(sizeof(double) * (size_t)n_days);
    double *temp_sunset_day = (double *)malloc(sizeof(double) * (size_t)n_days);
    double *temp_midnight_day = (double *)malloc(sizeof(double) * (size_t)n_days);
    for (int d = 0; d < n_days; ++d) {
        A_day[d] = (temp_max_day[d] - temp_min_day[d]) / 2;                           /* :217 */
        temp_sunset_day[d] = A_day[d] * Z_night + temp_min_day[d] + A_day[d];         /* :220 */
    }
    for (int d = 0; d < n_days; ++d) {
        const double temp_min_tomorrow = temp_min_day[d + 1 < n_days ? d + 1 : n_days - 1]; /* :213 */
        const double temp_sunset_yesterday = temp_sunset_day[d > 0 ? d - 1 : 0];            /* :223 */
        for (int h = 1; h <= 24; ++h) {
            const double times = 100.0 * h;                                           /* :201 */
            const double after_sunrise = (times > time_sunrise) ? 1.0 : 0.0;          /* :235 */
            const double temp_min_next = temp_min_day[d] * (1 - after_sunrise) +
                                         temp_min_tomorrow * after_sunrise;           /* :236-237 */
            const double before_sunset = (times < time_sunset) ? 1.0 : 0.0;           /* :242 */
            const double temp_sunset_last = temp_sunset_day[d] * (1 - before_sunset) +
                                            temp_sunset_yesterday * before_sunset;    /* :243-244 */
            const double is_light = (times >= time_sunrise && times <= time_sunset) ? 1.0 : 0.0; /* :247 */
            const double Z_hour_light = sin((360 * (times - time_reference) /
                                             (2 * (time_daily_max - time_sunrise))) / 57.29577); /* :255 */
            const double temp_hour_light = A_day[d] * Z_hour_light + temp_min_day[d] + A_day[d]; /* :256 */
            const double time_since_ss = times - time_sunset + 2400 * ((times < 1800) ? 1.0 : 0.0); /* :260 */
            const double E = tau * time_since_ss;                                     /* :261 */
            const double temp_hour_night = (temp_sunset_last - temp_min_next) / exp(E) + temp_min_next; /* :262 */
            temp_hour[(size_t)d * 24 + (h - 1)] =
                temp_hour_light * is_light + temp_hour_night * (1 - is_light);        /* :265 */
        }
        temp_midnight_day[d] = temp_hour[(size_t)d * 24 + 23];                        /* :271 */
    }
    /* :270-277 pre-dawn linear ramp from yesterday's midnight to today's minimum */
    for (int d = 0; d < n_days; ++d) {
        const double temp_midnight_yesterday = temp_midnight_day[d > 0 ? d - 1 : 0];  /* :272 */
        for (int h = 1; h <= 24; ++h) {
            const double times = 100.0 * h;
            if (times < time_sunrise) {                                               /* :276 */
                temp_hour[(size_t)d * 24 + (h - 1)] =
                    temp_min_day[d] + (temp_midnight_yesterday - temp_min_day[d]) *
                                          (1 - times / time_sunrise);                 /* :275 */
            }
        }
    }
    free(A_day); free(temp_sunset_day); free(temp_midnight_day);
}

static double pmax2(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a < b ? b : a); }
static double pmin2(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a > b ? b : a); }

/* vsine_weights(): R/sim_microclimate.R:431-454 */
void mmso_vsine_weights(const double *times, long n, double time_sunrise,
                        double time_inflection, double *weight)
{
    for (long k = 0; k < n; ++k) {
        const double t = times[k];
        const double mask_mn_to_sr = (t <= time_sunrise) ? 1.0 : 0.0;                 /* :436 */
        const double rel_mn_to_sr = (time_sunrise - t) / time_sunrise;                /* :437 */
        const double weight_mn_to_sr = (1 - 0.5 * pmax2(0, rel_mn_to_sr)) * mask_mn_to_sr; /* :438 */
        const double mask_sr_to_inf = (t > time_sunrise && t <= time_inflection) ? 1.0 : 0.0; /* :441 */
        const double rel_sr_to_inf = (t - time_sunrise) / (time_inflection - time_sunrise);   /* :442 */
        const double weight_sr_to_inf = (1 - pmin2(1, pmax2(0, rel_sr_to_inf))) * mask_sr_to_inf; /* :443 */
        const double mask_inf_to_mn = (t > time_inflection) ? 1.0 : 0.0;              /* :446 */
        const double rel_inf_to_mn = (t - time_inflection) / (24 - time_inflection);  /* :447 */
        const double weight_inf_to_mn = 0.5 * pmin2(1, pmax2(0, rel_inf_to_mn)) * mask_inf_to_mn; /* :448 */
        weight[k] = weight_mn_to_sr + weight_sr_to_inf + weight_inf_to_mn;            /* :451 */
    }
}

/* vsine_interpolate_daily() (VSINE): R/sim_microclimate.R:466-566, defaults
 * time_sunrise = 6, time_inflection = 13.  sunrise_is_max = 1 for humidity and
 * cloud cover (:316-328, :396-408), 0 for wind speed (:356-368). */
void mmso_vsine_interpolate_daily(const double *max_day, const double *min_day,
                                  int n_days, int sunrise_is_max, double *value_hour)
{
    const double time_sunrise = 6, time_inflection = 13;
    double times[24], weight[24];
    for (int h = 0; h < 24; ++h) times[h] = (double)(h + 1);                          /* :484 */
    mmso_vsine_weights(times, 24, time_sunrise, time_inflection, weight);             /* :492 */
    if (!sunrise_is_max)
        for (int h = 0; h < 24; ++h) weight[h] = 1 - weight[h];                       /* :498-500 */
    const double *sunval_day = sunrise_is_max ? max_day : min_day;                    /* :507-513 */
    const double *infval_day = sunrise_is_max ? min_day : max_day;
    for (int d = 0; d < n_days; ++d) {
        const double infval_yesterday = infval_day[d > 0 ? d - 1 : 0];                /* :522 */
        const double sunval_tomorrow = sunval_day[d + 1 < n_days ? d + 1 : n_days - 1]; /* :541 */
        for (int h = 0; h < 24; ++h) {
            const double before_sunrise = (times[h] < time_sunrise) ? 1.0 : 0.0;      /* :526 */
            const double infval_last = infval_day[d] * (1 - before_sunrise) +
                                       infval_yesterday * before_sunrise;             /* :529-530 */
            const double after_inflection = (times[h] > time_inflection) ? 1.0 : 0.0; /* :545 */
            const double sunval_next = sunval_day[d] * (1 - after_inflection) +
                                       sunval_tomorrow * after_inflection;            /* :548-549 */
            double v;
            if (!sunrise_is_max) {
                const double diff = infval_last - sunval_next;                        /* :553 */
                v = sunval_next + weight[h] * diff;                                   /* :554 */
            } else {
                const double diff = sunval_next - infval_last;                        /* :556 */
                v = infval_last + weight[h] * diff;                                   /* :557 */
            }
            value_hour[(size_t)d * 24 + h] = v;
        }
    }
}

/* hourly_from_daily_climate_ambient(): R/sim_microclimate.R:126-173.
 * daily is [9][D]; each output is [24*D]. */
void mmso_hourly_from_daily_ambient(const double *daily, int D,
                                    double *air_temperature, double *humidity,
                                    double *cloudcover, double *windspeed,
                                    double *water_temperature, double *rainfall)
{
    const double *tmin = daily + (size_t)MMSO_TMIN * D, *tmax = daily + (size_t)MMSO_TMAX * D;
    const double *rhmin = daily + (size_t)MMSO_RHMIN * D, *rhmax = daily + (size_t)MMSO_RHMAX * D;
    const double *wsmin = daily + (size_t)MMSO_WSMIN * D, *wsmax = daily + (size_t)MMSO_WSMAX * D;
    const double *ccmin = daily + (size_t)MMSO_CCMIN * D, *ccmax = daily + (size_t)MMSO_CCMAX * D;
    const double *rain = daily + (size_t)MMSO_RAIN * D;
    mmso_interpolate_daily_air_temp(tmax, tmin, D, air_temperature);                  /* :143-146 */
    mmso_vsine_interpolate_daily(rhmax, rhmin, D, 1, humidity);                       /* :147-150 */
    if (cloudcover) mmso_vsine_interpolate_daily(ccmax, ccmin, D, 1, cloudcover);     /* :151-154 */
    mmso_vsine_interpolate_daily(wsmax, wsmin, D, 0, windspeed);                      /* :155-158 */
    for (long t = 0; t < 24L * D; ++t) water_temperature[t] = air_temperature[t];     /* :163 */
    for (int d = 0; d < D; ++d)
        for (int h = 0; h < 24; ++h) rainfall[(size_t)d * 24 + h] = rain[d] / 24;     /* :132 */
}

/* ========================================================================
 * W: water body
 * ===================================================================== */

/* R's `^` operator for doubles: R sources src/main/arithmetic.c, R_pow() and
 * the y == 2 special case of R_POW (base R, restated from published source). */
double mmso_R_pow(double x, double y)
{
    if (y == 2.0) return x * x;
    if (x == 1. || y == 0.) return 1.;
    if (x == 0.) {
        if (y > 0.) return 0.;
        else if (y < 0) return INFINITY;
        else return y;
    }
    if (isfinite(x) && isfinite(y)) return pow(x, y);
    if (isnan(x) || isnan(y)) return x + y;
    if (!isfinite(x)) {
        if (x > 0) return (y < 0.) ? 0. : INFINITY;
        else {
            if (isfinite(y) && y == floor(y))
                return (y < 0.) ? 0. : (fmod(y, 2.) != 0 ? x : -x);
        }
    }
    if (!isfinite(y)) {
        if (x >= 0) {
            if (y > 0) return (x >= 1) ? INFINITY : 0.;
            else return (x < 1) ? INFINITY : 0.;
        }
    }
    return NAN;
}

/* R/sim_water.R:12-18; R's `pi` is M_PI */
double mmso_cone_depth_to_volume(double depth)
{
    const double radius = depth * sin(M_PI / 2);                 /* :13 */
    return M_PI * mmso_R_pow(radius, 2) * depth / 3;             /* :17 */
}

/* R/sim_water.R:24-31 */
double mmso_cone_volume_to_surface(double volume)
{
    return 3.046474 * mmso_R_pow(volume, 0.666667);
}

/* R/sim_water.R:36-38 (Buck equation) */
double mmso_saturation_vapour_pressure_kpa(double t)
{
    return 0.61121 * exp((18.678 - t / 234.5) * (t / (257.14 + t)));
}

/* R/sim_water.R:41-46 */
double mmso_air_pressure_kpa(double altitude_metres)
{
    const double p0 = 101.325;
    return p0 - 1.2 * altitude_metres / 100;
}

/* R/sim_water.R:50-90 */
double mmso_evaporation_rate_kg_h(double temperature_c, double relative_humidity,
                                  double altitude_m, double windspeed_m_s,
                                  double surface_area_m2)
{
    const double v = windspeed_m_s;
    const double theta = (25 + 19 * v);                                   /* :67 */
    const double A = surface_area_m2;
    const double p_ws = mmso_saturation_vapour_pressure_kpa(temperature_c);
    const double p_a = mmso_air_pressure_kpa(altitude_m);
    const double xs = 0.62198 * p_ws / (p_a - p_ws);                      /* :75 */
    const double p_w = p_ws * (relative_humidity / 100);                  /* :84 */
    const double x = 0.62198 * p_w / (p_a - p_w);                         /* :85 */
    return theta * A * (xs - x);                                          /* :88 */
}

/* cone: max_cone_volume, R/sim_water.R:257.  cylinder: EXTENSION. */
double mmso_habitat_max_volume(const mmso_habitat *hb)
{
    if (hb->shape == 0) return mmso_cone_depth_to_volume(hb->max_depth);
    return M_PI * (hb->radius * hb->radius) * hb->max_depth;
}

/* cone: cone_volume_to_surface, R/sim_water.R:24-31.  cylinder (EXTENSION):
 * vertical walls, so the free surface is pi r^2 whenever there is water. */
double mmso_habitat_surface(const mmso_habitat *hb, double volume)
{
    if (hb->shape == 0) return mmso_cone_volume_to_surface(volume);
    if (isnan(volume)) return volume;
    return (M_PI * (hb->radius * hb->radius)) * ((volume > 0) ? 1.0 : 0.0);
}

/* catchment_area, R/sim_water.R:261 (surface area of the full habitat) */
double mmso_habitat_catchment(const mmso_habitat *hb)
{
    return hb->inflow_multiplier * mmso_habitat_surface(hb, mmso_habitat_max_volume(hb));
}

/* iterate_cone_volume(): R/sim_water.R:94-132, with ensure_positive and
 * enforce_max from R/utils.R:159-168.  The leak term is an EXTENSION
 * (first-order seepage, fraction of the current volume per hour, removed with
 * evaporation before the non-negativity clamp); leak_per_hour = 0 reproduces
 * the reference bit for bit (x - 0.0 == x). */
double mmso_iterate_volume(const mmso_habitat *hb, double water_volume,
                           double rainfall_mm_h, double evaporation_rate_kg_h_m2,
                           double max_volume, double catchment_area)
{
    const double surface_area_m2 = mmso_habitat_surface(hb, water_volume);  /* :106 */
    const double loss_kg_h = evaporation_rate_kg_h_m2 * surface_area_m2;    /* :107 */
    const double loss = loss_kg_h / 1000;                                   /* :113 */
    double x = water_volume - loss;                                         /* :116 */
    x = x - hb->leak_per_hour * water_volume;                               /* EXTENSION */
    water_volume = x * ((x > 0) ? 1.0 : 0.0);                               /* ensure_positive */
    const double gain_m_h = rainfall_mm_h / 1000;                           /* :122 */
    const double gain = catchment_area * gain_m_h;                          /* :123 */
    double y = water_volume + gain;                                         /* :126 */
    if (y > max_volume) y = max_volume;                                     /* enforce_max */
    return y;
}

/* simulate_ephemeral_habitat_vectorised(): R/sim_water.R:216-301.  Matrices
 * are R time-by-pixel, column-major: element (t, c) at [c*H + t]. */
int mmso_simulate_ephemeral_habitat_vectorised(
    const double *rainfall, const double *air_temperature, const double *humidity,
    const double *windspeed, const double *altitude, long H, long C,
    const mmso_habitat *hb, int burnin_years,
    double *water_surface_area, double *water_volume, unsigned char *wet)
{
    const long n_times_year = 365 * 24;                                     /* :229 */
    if (burnin_years > 0 && H < n_times_year) return -1;    /* R: subscript out of bounds */
    const double max_volume = mmso_habitat_max_volume(hb);                  /* :257 */
    const double catchment_area = mmso_habitat_catchment(hb);               /* :261 */
    for (long c = 0; c < C; ++c) {
        const double *rn = rainfall + c * H, *ta = air_temperature + c * H;
        const double *rh = humidity + c * H, *ws = windspeed + c * H;
        double volume = hb->initial_volume;                                 /* :277 */
        for (int pass = 0; pass <= burnin_years; ++pass) {                  /* :232-233 */
            const long n = (pass < burnin_years) ? n_times_year : H;
            for (long t = 0; t < n; ++t) {
                const double rate = mmso_evaporation_rate_kg_h(ta[t], rh[t], altitude[c],
                                                               ws[t], 1);  /* :264-270 */
                volume = mmso_iterate_volume(hb, volume, rn[t], rate, max_volume,
                                             catchment_area);               /* :280-285 */
                if (pass == burnin_years) {                                 /* :290-291 */
                    if (water_surface_area)
                        water_surface_area[c * H + t] = mmso_habitat_surface(hb, volume); /* :294 */
                    if (water_volume) water_volume[c * H + t] = volume;
                    if (wet) wet[c * H + t] = (unsigned char)(volume > 0);
                }
            }
        }
    }
    return 0;
}

/* ========================================================================
 * Water temperature -- EXTENSION.  The reference only sketches this model, as
 * comments (query_terraclimate.R:370-497), stop()s when it is requested
 * (R/pipeline_functions.R:334-338) and copies the air temperature instead
 * (R/sim_microclimate.R:163).  The spec below is authored in DESIGN.md
 * section 5 from that prose; parity is unpinned by construction.
 * ===================================================================== */

/* zeta = omega (doy - 80) + 2 eps (sin(omega doy) - sin(omega 80)), delta = asin(0.39784993 sin zeta),
 * omega = 2 pi / 365, eps = 0.01675 (query_terraclimate.R:434-457) */
void mmso_solar_declination(int doy, double *sin_delta, double *cos_delta)
{
    const double omega = 2 * 3.141592653589793 / 365, eps = 0.01675;
    const double zeta = omega * (doy - 80) + 2 * eps * (sin(omega * doy) - sin(omega * 80));
    const double delta = asin(0.39784993 * sin(zeta));
    *sin_delta = sin(delta);
    *cos_delta = cos(delta);
}

/* h = 15 (t_d - 12) degrees, t_d the hour of the day (:459-466); the hourly series carries hour = 1..24 */
double mmso_cos_hour_angle(int hour)
{
    return cos(15.0 * (hour - 12) * (3.141592653589793 / 180));
}

/* The monthly clear-sky value is a daily mean; over the day it follows cos Z = cos(phi) cos(delta) cos(h) +
 * sin(phi) sin(delta) (:468-472; the notes' "cos(delta * h)" is read as the astronomical-triangle product), zero
 * while the sun is below the horizon.  (a/r)^2 (:425-432) is common to the 24 hours and cancels. */
void mmso_clear_sky_hourly(double latitude_deg, int doy, double clear_sky_mean, double *irr)
{
    double sd, cd;
    mmso_solar_declination(doy, &sd, &cd);
    const double phi = latitude_deg * (3.141592653589793 / 180);
    const double a = cos(phi) * cd, b = sin(phi) * sd;
    double w = 0;
    for (int k = 1; k <= 24; ++k) {
        const double cz = a * mmso_cos_hour_angle(k) + b;
        irr[k - 1] = (cz > 0) ? cz : 0;
        w += irr[k - 1];
    }
    const double scale = (w > 0) ? clear_sky_mean * 24 / w : 0;
    for (int k = 0; k < 24; ++k) irr[k] = scale * irr[k];
}

/* One hour of the water body with its temperature:
 *   water present at the start of the hour (volume V, surface A = S(V), mass m = 1000 V kg) exchanges
 *     Q_evap  = lambda E,  E = rate A kg evaporated (the loss of R/sim_water.R:107), lambda = 2.26e6 J/kg   (:388-393)
 *     Q_rad   = sigma eps A (Tw_K^4 - Ta_K^4) dt,  sigma = 5.67e-8, dt = 3600 s                             (:395-401)
 *     Q_solar = (1 - shade)(1 - albedo) I_clear (0.36 + 0.64 (1 - cloud / 100)) A dt   (Angstrom, :414-419)
 *   and its temperature moves by (Q_solar - Q_evap - Q_rad) / (m c + k), c = 4184 J/kg/K (:489-494), with
 *   k = 4 sigma eps Tw_K^3 A dt the derivative of Q_rad: the notes' explicit dT = Q / (m c) made linearly implicit
 *   in the radiative term, which is what keeps a film of water stable (m c -> 0); liquid water: clamped to
 *   [0, 100] C.  A dry body has the air's temperature.  Then the volume step of iterate_cone_volume(), and the rain
 *   that arrives (at air temperature) is mixed in by mass. */
void mmso_iterate_volume_temperature(const mmso_habitat *hb, const mmso_water_temperature *wt,
                                     double *volume_io, double *tw_io, double rainfall_mm_h,
                                     double evap_rate, double air_t, double cloud_pct, double i_clear,
                                     double max_volume, double catchment_area)
{
    const double volume = *volume_io, tw = *tw_io;
    const double area = mmso_habitat_surface(hb, volume);
    double tw_pre;
    if (!(volume > 0)) {
        tw_pre = air_t;
    } else {
        const double dt = 3600, se = 5.67e-8 * wt->emissivity;
        const double mass = 1000 * volume;
        const double q_evap = 2.26e6 * (evap_rate * area);
        const double twk = tw + 273.15, tak = air_t + 273.15;
        const double twk2 = twk * twk, tak2 = tak * tak;
        const double q_rad = se * area * (twk2 * twk2 - tak2 * tak2) * dt;
        const double angstrom = 0.36 + 0.64 * (1 - cloud_pct / 100);
        const double q_sol = (1 - wt->shade_proportion) * (1 - wt->albedo) * i_clear * angstrom * area * dt;
        /* optional sensible heat exchange h_c A (Ta - Tw) dt (absent from the notes; h_c = 0 by default), implicit like
         * the radiative term */
        const double q_conv = wt->convection_w_m2_k * area * (air_t - tw) * dt;
        const double k = 4 * se * (twk2 * twk) * area * dt + wt->convection_w_m2_k * area * dt;
        tw_pre = tw + (q_sol - q_evap - q_rad + q_conv) / (mass * 4184 + k);
        if (tw_pre < 0) tw_pre = 0;
        if (tw_pre > 100) tw_pre = 100;
    }
    /* iterate_cone_volume(), R/sim_water.R:94-132, as in mmso_iterate_volume() */
    const double loss = (evap_rate * area) / 1000;
    double x = volume - loss;
    x = x - hb->leak_per_hour * volume;
    x = x * ((x > 0) ? 1.0 : 0.0);
    const double gain = catchment_area * (rainfall_mm_h / 1000);
    double y = x + gain;
    *tw_io = (y > 0) ? (x * tw_pre + gain * air_t) / y : air_t;
    if (y > max_volume) y = max_volume;
    *volume_io = y;
}

/* the vectorised water simulation of one pixel with the temperature model on: same burn-in rule
 * (the first 8760 steps are repeated, R/sim_water.R:228-233), temperature carried like the volume */
static int simulate_habitat_temperature_one(const mmso_chain *ch, long c, const mmso_habitat *hb,
                                            const double *rain, const double *air_t, const double *rh,
                                            const double *wind, const double *cloud,
                                            double *area, double *volume, unsigned char *wet, double *water_t)
{
    const long H = 24L * ch->D, n_times_year = 365 * 24;
    if (ch->burnin_years_water > 0 && H < n_times_year) return -1;
    const double max_volume = mmso_habitat_max_volume(hb), catchment_area = mmso_habitat_catchment(hb);
    double vol = hb->initial_volume, tw = air_t[0];
    double irr[24];
    for (int pass = 0; pass <= ch->burnin_years_water; ++pass) {
        const long n = (pass < ch->burnin_years_water) ? n_times_year : H;
        for (long t = 0; t < n; ++t) {
            const int d = (int)(t / 24), h = (int)(t % 24);
            if (h == 0)
                mmso_clear_sky_hourly(ch->wt.latitude[c], ch->wt.day_of_year[d],
                                      ch->wt.clear_sky[(size_t)(ch->wt.month_of_day[d] - 1) * ch->C + c], irr);
            const double rate = mmso_evaporation_rate_kg_h(air_t[t], rh[t], ch->altitude[c], wind[t], 1);
            mmso_iterate_volume_temperature(hb, &ch->wt, &vol, &tw, rain[t], rate, air_t[t], cloud[t], irr[h],
                                            max_volume, catchment_area);
            if (pass == ch->burnin_years_water) {
                area[t] = mmso_habitat_surface(hb, vol);
                volume[t] = vol;
                wet[t] = (unsigned char)(vol > 0);
                water_t[t] = tw;
            }
        }
    }
    return 0;
}

/* ========================================================================
 * T: TerraClimate monthly variables -> the nine monthly forcing variables
 *    (SURVEY section 8f row N1): process_terraclimate_pixel_vars(),
 *    R/terraclimate.R:581-643, and its helpers in R/utils.R.
 * ===================================================================== */

/* NicheMapR::VAPPRS (saturation vapour pressure, Pa), called by NicheMapR::WETAIR for
 * $esat.  THIRD PARTY, NOT under /root/reference (mrke/NicheMapR, v3.3.1 named at
 * query_terraclimate.R:4-5; not declared in DESCRIPTION): restated from the published
 * Goff-Gratch (1946) formulation that VAPPRS uses -- over ice at or below the triple
 * point, over water above it.  parity unpinned. */
double mmso_vapprs(double db)
{
    const double t = db + 273.16;
    double loge;
    if (t <= 273.16) {
        loge = -9.09718 * (273.16 / t - 1.) - 3.56654 * log10(273.16 / t) +
               .876793 * (1. - t / 273.16) + log10(6.1071);
    } else {
        loge = -7.90298 * (373.16 / t - 1.) + 5.02808 * log10(373.16 / t) -
               1.3816E-07 * (mmso_R_pow(10., 11.344 * (1. - t / 373.16)) - 1.) +
               8.1328E-03 * (mmso_R_pow(10., -3.49149 * (373.16 / t - 1.)) - 1.) + log10(1013.246);
    }
    return mmso_R_pow(10., loge) * 100;
}

/* vapour_pressure(): R/utils.R:64-78.  WETAIR(db, rh = 100)$e is esat * rh / 100. */
double mmso_vapour_pressure(double tmean, double vpd_kpa)
{
    const double vapour_pressure = mmso_vapprs(tmean) * 100 / 100;       /* :67-71 */
    const double vpd_pa = vpd_kpa * 1000;                                /* :76 */
    return vapour_pressure - vpd_pa;                                     /* :77 */
}

/* relative_humidity(): R/utils.R:80-89.  pmin / pmax keep NaN. */
double mmso_relative_humidity(double temperature, double vapour_pressure)
{
    const double saturation_vapour_pressure = mmso_vapprs(temperature);  /* :82-83 */
    double rh = (vapour_pressure / saturation_vapour_pressure) * 100;    /* :85 */
    if (rh > 100) rh = 100;                                              /* :86 */
    if (rh < 0.01) rh = 0.01;                                            /* :87 */
    return rh;
}

/* adjust_wind_speed(): R/utils.R:97-134; exponent by terrain (:123-127), "level grass" = 0.15 */
double mmso_adjust_wind_speed(double wind_speed, double height_required, double height_measured,
                              double exponent)
{
    return wind_speed * mmso_R_pow(height_required / height_measured, exponent);   /* :132 */
}

/* cloud_cover(): R/utils.R:141-147 */
double mmso_cloud_cover(double solar_radiation, double clear_sky_radiation, double multiplier)
{
    double cloud = (1 - solar_radiation / clear_sky_radiation) * 100;    /* :142 */
    cloud = cloud * multiplier;                                          /* :143 */
    if (cloud > 100) cloud = 100;                                        /* :144 */
    if (cloud < 0) cloud = 0;                                            /* :145 */
    return cloud;
}

/* process_terraclimate_pixel_vars(): R/terraclimate.R:581-643, for C pixels at once.
 * terraclimate [6][M][C] in the order tmax, tmin, ppt, ws, vpd, srad (the tibble of
 * R/pipeline_functions.R:134-136); clear_sky [12][C] = clear_sky_SOLR by month_id;
 * month_id[M] = lubridate::month(start + (end - start) / 2) (:591), n_days[M] =
 * as.numeric(end - start) (:592: days in the month MINUS ONE -- replicated, not fixed).
 * monthly out [9][M][C] in the order of MMSO_F_*. */
void mmso_process_terraclimate_pixel_vars(const double *terraclimate, const double *clear_sky,
                                          const int *month_id, const double *n_days,
                                          long C, int M, double *monthly)
{
    const long plane = (long)M * C;
    for (int m = 0; m < M; ++m) {
        for (long c = 0; c < C; ++c) {
            const long at = (long)m * C + c;
            const double tmax = terraclimate[0 * plane + at], tmin = terraclimate[1 * plane + at];
            const double ppt = terraclimate[2 * plane + at], ws = terraclimate[3 * plane + at];
            const double vpd = terraclimate[4 * plane + at], srad = terraclimate[5 * plane + at];
            const double clear_sky_SOLR = clear_sky[(long)(month_id[m] - 1) * C + c];   /* :600-603 */
            const double tmean = (tmax + tmin) / 2;                      /* :607 */
            const double vp = mmso_vapour_pressure(tmean, vpd);          /* :608 */
            const double rhmax = mmso_relative_humidity(tmin, vp);       /* :609 */
            const double rhmin = mmso_relative_humidity(tmax, vp);       /* :610 */
            const double wsmax = mmso_adjust_wind_speed(ws, 2, 10, 0.15);/* :615-617 */
            const double wsmin = wsmax * 0.1;                            /* :618 */
            monthly[0 * plane + at] = tmin;
            monthly[1 * plane + at] = tmax;
            monthly[2 * plane + at] = rhmin;
            monthly[3 * plane + at] = rhmax;
            monthly[4 * plane + at] = wsmin;
            monthly[5 * plane + at] = wsmax;
            monthly[6 * plane + at] = mmso_cloud_cover(srad, clear_sky_SOLR, 0.8);   /* ccmin :625 */
            monthly[7 * plane + at] = mmso_cloud_cover(srad, clear_sky_SOLR, 1.2);   /* ccmax :624 */
            monthly[8 * plane + at] = ppt / n_days[m];                   /* :630 */
        }
    }
}

/* ========================================================================
 * L: life history
 * ===================================================================== */

/* base-R stats::splinefun(method = "fmm") coefficients: R sources
 * src/library/stats/src/splines.c fmm_spline() (Forsythe, Malcolm & Moler;
 * restated from the published algorithm).  Needed only to build synthetic
 * tables: the product receives coefficients from the caller. */
void mmso_fmm_spline(int n, const double *x0, const double *y0, double *b0, double *c0, double *d0)
{
    const double *x = x0 - 1, *y = y0 - 1;      /* 1-based as in the source */
    double *b = b0 - 1, *c = c0 - 1, *d = d0 - 1;
    double t;
    int i;
    if (n < 2) return;
    if (n < 3) {
        t = (y[2] - y[1]);
        b[1] = t / (x[2] - x[1]);
        b[2] = b[1];
        c[1] = c[2] = d[1] = d[2] = 0.0;
        return;
    }
    const int nm1 = n - 1;
    d[1] = x[2] - x[1];
    c[2] = (y[2] - y[1]) / d[1];
    for (i = 2; i < n; i++) {
        d[i] = x[i + 1] - x[i];
        b[i] = 2 * (d[i - 1] + d[i]);
        c[i + 1] = (y[i + 1] - y[i]) / d[i];
        c[i] = c[i + 1] - c[i];
    }
    b[1] = -d[1];
    b[n] = -d[nm1];
    c[1] = c[n] = 0.0;
    if (n > 3) {
        c[1] = c[3] / (x[4] - x[2]) - c[2] / (x[3] - x[1]);
        c[n] = c[nm1] / (x[n] - x[n - 2]) - c[n - 2] / (x[nm1] - x[n - 3]);
        c[1] = c[1] * d[1] * d[1] / (x[4] - x[1]);
        c[n] = -c[n] * d[nm1] * d[nm1] / (x[n] - x[n - 3]);
    }
    for (i = 2; i <= n; i++) {
        t = d[i - 1] / b[i - 1];
        b[i] = b[i] - t * d[i - 1];
        c[i] = c[i] - t * c[i - 1];
    }
    c[n] = c[n] / b[n];
    for (i = nm1; i >= 1; i--) c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
    b[n] = (y[n] - y[n - 1]) / d[nm1] + d[nm1] * (c[nm1] + 2 * c[n]);
    for (i = 1; i <= nm1; i++) {
        b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2 * c[i]);
        d[i] = (c[i + 1] - c[i]) / d[i];
        c[i] = 3 * c[i];
    }
    c[n] = 3 * c[n];
    d[n] = d[nm1];
}

/* base-R spline_eval() for method "fmm" (same source file; restated).  Note
 * the interval index i is carried from one evaluation point to the next, as
 * in R, because the reference evaluates a whole hourly vector per call
 * (R/sim_pop_dynamics.R:108-114). */
void mmso_spline_eval(const mmso_spline *s, const double *u, long nu, double *v)
{
    const int n = s->n, n_1 = n - 1;
    const double *x = s->x;
    int i = 0;
    for (long l = 0; l < nu; l++) {
        const double ul = u[l];
        if (ul < x[i] || (i < n_1 && x[i + 1] < ul)) {
            i = 0;
            int j = n;
            do {
                int k = (i + j) / 2;
                if (ul < x[k]) j = k; else i = k;
            } while (j > i + 1);
        }
        const double dx = ul - x[i];
        v[l] = s->y[i] + dx * (s->b[i] + dx * (s->c[i] + dx * s->d[i]));
    }
}

/* positive_spline(): R/lifehistory_estimation_functions.R:236-250 */
void mmso_positive_spline(const mmso_spline *s, const double *u, long nu, double *v)
{
    mmso_spline_eval(s, u, nu, v);
    for (long l = 0; l < nu; ++l) v[l] = v[l] * ((v[l] > 0) ? 1.0 : 0.0);  /* ensure_positive */
}

/* fields::interp.surface(obj, loc) (package `fields`, not under
 * /root/reference, version unpinned; restated from its published R source),
 * as called by make_temp_humid_interpolator,
 * R/lifehistory_estimation_functions.R:1144-1190. */
void mmso_interp_surface(const mmso_surface *g, const double *tx, const double *ty,
                         long n, double *out)
{
    const int nx = g->nx, ny = g->ny;
    double *ix = (double *)malloc(sizeof(double) * (size_t)(nx > ny ? nx : ny));
    for (int k = 0; k < (nx > ny ? nx : ny); ++k) ix[k] = (double)(k + 1);   /* 1:nx */
    for (long l = 0; l < n; ++l) {
        double lx = isnan(tx[l]) ? tx[l] : mmso_approx1(tx[l], g->x, ix, nx);
        double ly = isnan(ty[l]) ? ty[l] : mmso_approx1(ty[l], g->y, ix, ny);
        if (isnan(lx) || isnan(ly)) { out[l] = NAN; continue; }
        double lx1 = floor(lx), ly1 = floor(ly);
        double ex = lx - lx1, ey = ly - ly1;
        if (lx1 == nx) { ex = 1; lx1 = nx - 1; }
        if (ly1 == ny) { ey = 1; ly1 = ny - 1; }
        const long i = (long)lx1 - 1, j = (long)ly1 - 1;         /* to 0-based */
        const double z00 = g->z[i + (long)nx * j], z10 = g->z[i + 1 + (long)nx * j];
        const double z01 = g->z[i + (long)nx * (j + 1)], z11 = g->z[i + 1 + (long)nx * (j + 1)];
        out[l] = z00 * (1 - ex) * (1 - ey) + z10 * ex * (1 - ey) +
                 z01 * (1 - ex) * ey + z11 * ex * ey;
    }
    free(ix);
}

/* make_surv_densmod_function(): R/lifehistory_estimation_functions.R:525-587 */
double mmso_densmod_apply(const mmso_densmod *dm, double surv_prob, double density)
{
    const double density_dish = density * dm->surface_area_cm2;                 /* :549 / :569 */
    if (dm->type == 1) {                                                        /* cox_ph */
        const double daily_mortality_zero_density = 1 - surv_prob;              /* :551 */
        const double loghaz0 = log(-log(1 - daily_mortality_zero_density));     /* :553 */
        const double loghaz = loghaz0 + dm->dd_effect * density_dish;           /* :556-557 */
        const double daily_mortality = 1 - exp(-exp(loghaz));                   /* :560 */
        return 1 - daily_mortality;                                             /* :562 */
    }
    const double logit0 = log(surv_prob / (1 - surv_prob));                     /* :533, :571 */
    const double lg = logit0 + dm->dd_effect * density_dish;                    /* :574-575 */
    return 1 / (1 + exp(-lg));                                                  /* :537, :578 */
}

/* ========================================================================
 * P: population
 * ===================================================================== */

/* iterate_state(): R/sim_pop_dynamics.R:14-96 */
void mmso_iterate_state(double *aquatic_io, double *adult_io, const mmso_densmod *dm,
                        double larval_habitat_area, double mdr, double efd,
                        double das_zerodensity, double ds)
{
    const double timestep = 1.0 / 24;                                   /* :23 */
    const double epsilon = sqrt(2.220446049250313e-16);                 /* :26 */
    double aquatic = *aquatic_io, adult = *adult_io;
    if (isnan(aquatic)) aquatic = 0;                                    /* :33 */
    if (isnan(adult)) adult = 0;                                        /* :34 */
    const double aquatic_density = aquatic / (larval_habitat_area + epsilon);   /* :42 */
    const double das_t = mmso_densmod_apply(dm, das_zerodensity, aquatic_density);    /* :43-44 */
    const double aquatic_survival_prob = mmso_R_pow(das_t, timestep);   /* :49 */
    const double adult_survival_prob = mmso_R_pow(ds, timestep);        /* :50 */
    const double emergence_rate = timestep * mdr;                       /* :58 */
    const double emergence_fraction = 1 - exp(-emergence_rate);         /* :59 */
    const double egg_laying_rate = efd * timestep;                      /* :62 */
    const double surviving_adult = adult * adult_survival_prob;         /* :67 */
    const double surviving_aquatic = aquatic * aquatic_survival_prob;   /* :68 */
    const double new_aquatic = surviving_adult * egg_laying_rate;       /* :71 */
    const double remaining_aquatic = surviving_aquatic * (1 - emergence_fraction); /* :74 */
    const double new_adult = surviving_aquatic * emergence_fraction;    /* :77 */
    const double imported_adult = epsilon;                              /* :81 */
    *aquatic_io = remaining_aquatic + new_aquatic;                      /* :92 */
    *adult_io = surviving_adult + new_adult + imported_adult;           /* :93 */
}

/* simulate_population_vectorised(): R/sim_pop_dynamics.R:208-284; matrices
 * column-major time-by-pixel like R. */
int mmso_simulate_population_vectorised(
    const double *larval_habitat_area, const double *mdr, const double *efd,
    const double *das_zerodensity, const double *ds, long H, long C,
    const mmso_densmod *dm, double initial_adult, double initial_aquatic,
    int burnin_years, double *aquatic_out, double *adult_out)
{
    const long n_times_year = 365 * 24;                                 /* :222 */
    if (burnin_years > 0 && H < n_times_year) return -1;
    for (long c = 0; c < C; ++c) {
        double adult = initial_adult, aquatic = initial_aquatic;        /* :249-252 */
        for (int pass = 0; pass <= burnin_years; ++pass) {              /* :225-226 */
            const long n = (pass < burnin_years) ? n_times_year : H;
            for (long t = 0; t < n; ++t) {
                const long k = c * H + t;
                mmso_iterate_state(&aquatic, &adult, dm, larval_habitat_area[k], mdr[k],
                                   efd[k], das_zerodensity[k], ds[k]);  /* :259-268 */
                if (pass == burnin_years) {                             /* :276-282 */
                    aquatic_out[k] = aquatic;
                    adult_out[k] = adult;
                }
            }
        }
    }
    return 0;
}

/* simulate_ephemeral_habitat(): R/sim_water.R:136-205, the one-pixel twin.  Its burn-in
 * index is rep(seq_len(n_times), burnin_years + 1) (:151-152, which overrides :144-149):
 * the WHOLE series is repeated, and tail(..., n_times_total) keeps the last repeat (:198-199).
 * Returns cone_volume_to_surface(volumes) (:203). */
int mmso_simulate_ephemeral_habitat(
    const double *rainfall, const double *air_temperature, const double *humidity,
    const double *windspeed, double altitude, long n_times,
    const mmso_habitat *hb, int burnin_years, double *surface_areas)
{
    const double max_volume = mmso_habitat_max_volume(hb);                  /* :164 */
    const double catchment_area = mmso_habitat_catchment(hb);               /* :168 */
    double volume = hb->initial_volume;                                     /* :182 */
    for (int rep = 0; rep <= burnin_years; ++rep) {                         /* :152 */
        for (long t = 0; t < n_times; ++t) {                                /* :184 */
            const double rate = mmso_evaporation_rate_kg_h(air_temperature[t], humidity[t], altitude,
                                                           windspeed[t], 1); /* :171-177 */
            volume = mmso_iterate_volume(hb, volume, rainfall[t], rate, max_volume, catchment_area); /* :185-194 */
            if (rep == burnin_years) surface_areas[t] = mmso_habitat_surface(hb, volume); /* :198-203 */
        }
    }
    return 0;
}

/* simulate_population(): R/sim_pop_dynamics.R:138-196, the one-pixel twin; burn-in index
 * rep(seq_len(n_times), burnin_years + 1) (:147-148), keep tail(..., n_times) (:187). */
int mmso_simulate_population(
    const double *larval_habitat_area, const double *mdr, const double *efd,
    const double *das_zerodensity, const double *ds, long n_times,
    const mmso_densmod *dm, double initial_adult, double initial_aquatic,
    int burnin_years, double *aquatic_out, double *adult_out)
{
    double adult = initial_adult, aquatic = initial_aquatic;               /* :161-164 */
    for (int rep = 0; rep <= burnin_years; ++rep) {
        for (long t = 0; t < n_times; ++t) {                                /* :169 */
            mmso_iterate_state(&aquatic, &adult, dm, larval_habitat_area[t], mdr[t], efd[t],
                               das_zerodensity[t], ds[t]);                  /* :171-180 */
            if (rep == burnin_years) {                                      /* :187-193 */
                aquatic_out[t] = aquatic;
                adult_out[t] = adult;
            }
        }
    }
    return 0;
}

/* base-R mean() for doubles: R sources src/main/summary.c real_mean()
 * (long double accumulation plus one correction pass; restated).  Used by
 * aggregate_vectors(), R/pipeline_functions.R:1055-1086. */
double mmso_R_mean(const double *x, long n)
{
    long double s = 0.0L;
    for (long i = 0; i < n; ++i) s += x[i];
    s /= n;
    long double t = 0.0L;
    for (long i = 0; i < n; ++i) t += (x[i] - s);
    s += t / n;
    return (double)s;
}

/* ========================================================================
 * fused chain: tc_to_vectors() from interpolate_daily_climate() onwards,
 * R/pipeline_functions.R:1099-1130, for one cell at a time.
 * ===================================================================== */

int mmso_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int chain_one_cell(const mmso_chain *ch, long c, double *work)
{
    const int M = ch->M, D = ch->D, S = ch->S, P = ch->P;
    const long C = ch->C, H = 24L * D;
    /* workspace layout */
    double *monthly = work;                 work += (size_t)MMSO_NFORCING * M;
    double *daily = work;                   work += (size_t)MMSO_NFORCING * D;
    double *air_t = work;                   work += H;
    double *rh = work;                      work += H;
    double *cloud = work;                   work += H;
    double *wind = work;                    work += H;
    double *water_t = work;                 work += H;
    double *rain = work;                    work += H;
    double *mdr = work;                     work += H;
    double *efd = work;                     work += H;
    double *das0 = work;                    work += H;
    double *ds = work;                      work += H;
    double *area = work;                    work += H;
    double *volume = work;                  work += H;
    double *aquatic = work;                 work += H;
    double *adult = work;                   work += H;
    double *gather = work;                  work += H;
    unsigned char *wet = (unsigned char *)work;

    for (int v = 0; v < MMSO_NFORCING; ++v)
        for (int m = 0; m < M; ++m)
            monthly[(size_t)v * M + m] = ch->monthly[((size_t)v * M + m) * C + c];

    /* interpolate_daily_climate(), R/pipeline_functions.R:258-275 */
    int rc = mmso_daily_from_monthly(monthly, ch->mid_day, M, ch->day_predict, D, daily);
    if (rc) return rc;
    /* simulate_hourly_microclimate(), :284-301 */
    mmso_hourly_from_daily_ambient(daily, D, air_t, rh, cloud, wind, water_t, rain);
    /* simulate_hourly_lifehistory(), :634-683 -> simulate_vector_lifehistory(),
     * R/sim_pop_dynamics.R:100-135 (state independent, so hoisted above the
     * settings loop; the reference order is conditions -> lifehistory) */
    mmso_positive_spline(&ch->lh.mdr, water_t, H, mdr);        /* :108 */
    mmso_positive_spline(&ch->lh.efd, air_t, H, efd);          /* :110 */
    mmso_positive_spline(&ch->lh.das, water_t, H, das0);       /* :114 */
    mmso_interp_surface(&ch->lh.ds, air_t, rh, H, ds);         /* :116-119 */

    const double altitude = ch->altitude[c];
    for (int s = 0; s < S; ++s) {
        /* simulate_hourly_conditions(), R/pipeline_functions.R:319-538 */
        if (ch->wt.enabled) {
            /* EXTENSION: model_water_temperature = TRUE.  The water temperature is a state of the water body, so
             * the two rates that take it (R/sim_pop_dynamics.R:108,114) are evaluated per setting */
            rc = simulate_habitat_temperature_one(ch, c, &ch->habitats[s], rain, air_t, rh, wind, cloud,
                                                  area, volume, wet, water_t);
            if (rc) return rc;
            mmso_positive_spline(&ch->lh.mdr, water_t, H, mdr);
            mmso_positive_spline(&ch->lh.das, water_t, H, das0);
        } else {
        rc = mmso_simulate_ephemeral_habitat_vectorised(rain, air_t, rh, wind, &altitude, H, 1,
                                                        &ch->habitats[s], ch->burnin_years_water,
                                                        area, volume, wet);
        if (rc) return rc;
        }
        /* simulate_hourly_vectors(), :691-906; larval_habitat_area =
         * water_surface_area (R/sim_pop_dynamics.R:123) */
        rc = mmso_simulate_population_vectorised(area, mdr, efd, das0, ds, H, 1, &ch->densmod,
                                                 ch->initial_adult, ch->initial_aquatic,
                                                 ch->burnin_years_pop, aquatic, adult);
        if (rc) return rc;

        const double *series[MMSO_NVARS] = { adult, aquatic, mdr, efd, das0, ds,
                                             area, volume, air_t, rh, wind, rain, water_t };
        if (ch->out_hourly) {
            for (int v = 0; v < MMSO_NVARS; ++v)
                for (long t = 0; t < H; ++t)
                    ch->out_hourly[(((size_t)s * MMSO_NVARS + v) * H + t) * C + c] = series[v][t];
        }
        if (ch->out_wet)
            for (long t = 0; t < H; ++t) ch->out_wet[((size_t)s * H + t) * C + c] = wet[t];
        /* summarise_vectors() -> aggregate_vectors(), :995-1091 */
        if (ch->out_mean || ch->out_wet_hours) {
            for (int p = 0; p < P; ++p) {
                for (int v = 0; v < MMSO_NVARS; ++v) {
                    long n = 0;
                    int wet_hours = 0;
                    for (int d = 0; d < D; ++d) {
                        if (ch->period_of_day[d] != p) continue;
                        for (int h = 0; h < 24; ++h) {
                            gather[n++] = series[v][(size_t)d * 24 + h];
                            wet_hours += wet[(size_t)d * 24 + h];
                        }
                    }
                    if (ch->out_mean)
                        ch->out_mean[(((size_t)s * MMSO_NVARS + v) * P + p) * C + c] =
                            n ? mmso_R_mean(gather, n) : NAN;
                    if (ch->out_wet_hours && v == 0)
                        ch->out_wet_hours[((size_t)s * P + p) * C + c] = wet_hours;
                }
            }
        }
    }
    return 0;
}

int mmso_run_chain(const mmso_chain *ch)
{
    const long H = 24L * ch->D;
    const size_t work_doubles = (size_t)MMSO_NFORCING * ch->M + (size_t)MMSO_NFORCING * ch->D +
                                (size_t)15 * H + (size_t)(H / 8 + 2);
    int status = 0;
    int nt = ch->n_threads > 0 ? ch->n_threads : mmso_max_threads();
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        double *work = (double *)malloc(sizeof(double) * work_doubles);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (long c = 0; c < ch->C; ++c) {
            if (!work) { status = -2; continue; }
            int rc = chain_one_cell(ch, c, work);
            if (rc) status = rc;
        }
        free(work);
    }
    return status;
}
