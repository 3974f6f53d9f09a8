// mms_math.cuh -- per-step device arithmetic of the mosmicrosim chain.
//
// Compiled twice (see build.py): MMS_STRICT=1 with -fmad=false keeps the
// reference's literal operation order and CUDA's IEEE-accurate libm, so that the
// only differences from the CPU oracle are libm last-bit differences;
// MMS_STRICT=0 ("fast", the default at run time) allows FMA contraction, turns
// pow() into exp(y*log(x)) compositions and divisions by literals into
// multiplications, and uses the branch-free exp/log of mms_fastmath.cuh.
// Cheap, badly conditioned operations (1 - (1 - p), 1 - exp(-x), lx - floor(lx),
// the clamps) are kept literal in both modes: the target is the reference's
// floating-point arithmetic, not the exact mathematics.
//
// Citations are relative to the reference root.
#pragma once

#include "mms_common.cuh"
#include "mms_fastmath.cuh"

#ifndef MMS_STRICT
#define MMS_STRICT 0
#endif

#if MMS_STRICT
#define MMS_MODE strict
#else
#define MMS_MODE fast
#endif

namespace mms {
namespace MMS_MODE {

constexpr double kEps = 1.4901161193847656e-08;   // sqrt(.Machine$double.eps), R/sim_pop_dynamics.R:26
constexpr double kTimestep = 1.0 / 24.0;          // timestep = 1 / 24, R/sim_pop_dynamics.R:23

// x / c for a literal c of the R source; exp / log / division flavour
#if MMS_STRICT
#define MMS_DIVC(x, c) ((x) / (c))
#else
#define MMS_DIVC(x, c) ((x) * (1.0 / (c)))
#endif

// R's `^` (arithmetic.c R_pow) for a positive, non-integer literal exponent y.
__device__ __forceinline__ double r_pow_pos(double x, double y)
{
#if MMS_STRICT
    if (x == 1.0) return 1.0;
    if (x == 0.0) return 0.0;           // also -0.0 -> +0.0, like R_pow
    return pow(x, y);
#else
    return ::mms::fm::pow_unit(x, y);    // callers: water volume, survival probabilities (0 <= x, or NaN)
#endif
}

// saturation_vapour_pressure_kpa(), R/sim_water.R:36-38 (Buck)
#if !MMS_STRICT
static __constant__ double kEvapC[5] = {0.61121, 18.678, 1.0 / 234.5, 257.14, 0.62198};
// 3.046474, 0.666667 (R/sim_water.R:30), 1 / 1000 (:107), timestep and -timestep (R/sim_pop_dynamics.R:23)
static __constant__ double kLitC[5] = {3.046474, 0.666667, 1.0 / 1000.0, 1.0 / 24.0, -(1.0 / 24.0)};
#define MMS_LIT(i, literal) kLitC[i]
#else
#define MMS_LIT(i, literal) (literal)
#endif

__device__ __forceinline__ double svp_kpa(double t)
{
#if MMS_STRICT
    return 0.61121 * exp((18.678 - t / 234.5) * (t / (257.14 + t)));
#else
    // the exponent is within +-20 for any temperature between -100 and +1000 C: no range guard
    // literals through constant memory: as immediates each costs two UMOVs per use in the rolled hour loop
    return kEvapC[0] * ::mms::fm::exp_core((kEvapC[1] - t * kEvapC[2]) * ::mms::fm::div_fast(t, kEvapC[3] + t));
#endif
}

// evaporation_rate_kg_h() with surface_area_m2 = 1, R/sim_water.R:50-90,264-270.
// p_a = air_pressure_kpa(altitude) (R/sim_water.R:41-46) is per cell.
__device__ __forceinline__ double evaporation_rate(double t, double rh, double p_a, double wind)
{
    const double theta = 25.0 + 19.0 * wind;                   // :67
    const double p_ws = svp_kpa(t);                            // :73
#if MMS_STRICT
    const double xs = 0.62198 * p_ws / (p_a - p_ws);           // :75
    const double p_w = p_ws * (rh / 100.0);                    // :84
    const double x = 0.62198 * p_w / (p_a - p_w);              // :85
    return theta * (xs - x);                                   // :88 (A = 1: theta * 1 == theta)
#else
    // xs - x over one denominator (one division instead of two; no cancellation of two
    // rounded quotients): 0.62198 p_a (p_ws - p_w) / ((p_a - p_ws)(p_a - p_w))
    const double p_w = p_ws * (rh * (1.0 / 100.0));
    const double num = (kEvapC[4] * p_a) * (p_ws - p_w);
    const double den = (p_a - p_ws) * (p_a - p_w);
    return theta * ::mms::fm::div_fast(num, den);
#endif
}

__device__ __forceinline__ double air_pressure_kpa(double altitude)
{
    return 101.325 - MMS_DIVC(1.2 * altitude, 100.0);          // R/sim_water.R:44-45
}

// ensure_positive(), R/utils.R:159-161: x * as.numeric(x > 0).  x > 0 -> x; +-0 -> +-0;
// x < 0 -> -0.0; NaN -> NaN.  Also used by positive_spline(),
// R/lifehistory_estimation_functions.R:236-250.
__device__ __forceinline__ double ensure_positive(double v)
{
#if MMS_STRICT
    return (v > 0.0) ? v : ((v != v) ? v : ((v == 0.0) ? v : -0.0));
#else
    // same values (a zero of either sign behaves identically downstream), one compare
    return (v < 0.0) ? 0.0 : v;
#endif
}

// cone_volume_to_surface(), R/sim_water.R:24-31; cylinder is an EXTENSION
__device__ __forceinline__ double habitat_surface(const HabitatDev &hb, double volume, int shape)
{
    if (shape == 0) return MMS_LIT(0, 3.046474) * r_pow_pos(volume, MMS_LIT(1, 0.666667));
    return (volume != volume) ? volume : ((volume > 0.0) ? hb.cyl_area : 0.0);
}
__device__ __forceinline__ double habitat_surface(const HabitatDev &hb, double volume)
{
    return habitat_surface(hb, volume, hb.shape);
}

// iterate_cone_volume(), R/sim_water.R:94-132; `surface` is
// habitat_surface(volume) (already known from the previous step's output).
// gain = catchment_area * (rainfall / 1000), :122-123, is constant over a day.
__device__ __forceinline__ double iterate_volume(const HabitatDev &hb, double volume, double surface,
                                                 double evap_rate, double gain)
{
#if MMS_STRICT
    const double loss = (evap_rate * surface) / 1000.0;        // :107,:113
#else
    const double loss = (evap_rate * surface) * kLitC[2];
#endif
    double x = volume - loss;                                  // :116
    x = x - hb.leak_per_hour * volume;                         // EXTENSION (leak = 0: x - 0 == x)
    x = ensure_positive(x);                                    // R/utils.R:159-161
    double y = x + gain;                                       // :126
    y = (y > hb.max_volume) ? hb.max_volume : y;               // enforce_max(), R/utils.R:164-168
    return y;
}

#if !MMS_STRICT
// The water step of the N settings a thread integrates, stage by stage (mms_fastmath.cuh: why).  hab = the first of
// the N consecutive settings; shape[i]: 0 cone, 1 cylinder.  Same operations as iterate_volume() / habitat_surface().
template <int N>
__device__ __forceinline__ void water_step_n(const HabitatDev *hab, double (&volume)[N], double (&surface)[N],
                                             double evap_rate, const double (&gain)[N], const int (&shape)[N])
{
    double x[N], y[N];
    MMS_EACH(i, N) x[i] = volume[i] - (evap_rate * surface[i]) * kLitC[2];         // R/sim_water.R:107-116
    MMS_EACH(i, N) x[i] = x[i] - hab[i].leak_per_hour * volume[i];                 // EXTENSION (leak = 0: x - 0 == x)
    MMS_EACH(i, N) x[i] = ensure_positive(x[i]);                                   // R/utils.R:159-161
    MMS_EACH(i, N) y[i] = x[i] + gain[i];                                          // :126
    MMS_EACH(i, N) volume[i] = (y[i] > hab[i].max_volume) ? hab[i].max_volume : y[i];   // enforce_max(), R/utils.R:164-168
    bool all_cones = true;
    MMS_EACH(i, N) all_cones = all_cones && shape[i] == 0;
    if (all_cones) {                                                               // cone_volume_to_surface(), :24-31
        double pw[N];
        ::mms::fm::pow_unit_n<N>(volume, kLitC[1], pw);
        MMS_EACH(i, N) surface[i] = kLitC[0] * pw[i];
    } else {
        MMS_EACH(i, N) surface[i] = habitat_surface(hab[i], volume[i], shape[i]);
    }
}
#endif

// State-independent part of iterate_state() for one hour.
struct PopHour {
    double surv_ad;     // ds ^ timestep                          R/sim_pop_dynamics.R:50
    double emf;         // 1 - exp(-timestep * mdr)               :58-59
    double one_m_emf;   // 1 - emergence_fraction                 :74
    double egg;         // efd * timestep                         :62
    double link0;       // logit(das0) / cox_ph link of das0: the state independent part of densmod
};

// the part that takes the AIR temperature / humidity rates: adult survival, eggs
__device__ __forceinline__ void pop_hour_air(PopHour &ph, double efd, double ds)
{
    ph.surv_ad = r_pow_pos(ds, MMS_LIT(3, kTimestep));
    ph.egg = efd * MMS_LIT(3, kTimestep);
}

// exp(-timestep * mdr): 1 - emergence_fraction, R/sim_pop_dynamics.R:58-59,74
__device__ __forceinline__ double pop_emergence_complement(double mdr)
{
#if MMS_STRICT
    return exp(-(kTimestep * mdr));
#else
    return ::mms::fm::exp_core(kLitC[4] * mdr);   // mdr is a clamped spline value: 0 <= mdr << 708 * 24
#endif
}

// the density-free part of the density dependence: logit(das0), or the cox_ph link of das0
// (R/lifehistory_estimation_functions.R:551-553 / :533,:571)
__device__ __forceinline__ double pop_link0(int densmod_type, double das0)
{
    if (densmod_type == 1) {
        const double daily_mortality_zero_density = 1.0 - das0;
#if MMS_STRICT
        return log(-log(1.0 - daily_mortality_zero_density));
#else
        // log p'; exp(loghaz) = -log p' * exp(dd*D).  A spline that overshoots 1 (positive_spline clamps at 0 only):
        // p' > 1 -> the reference takes log(-log p') of a negative number: NaN; p' == 1 -> log(-0) = -Inf, das_t = 1:
        // log p' must be an exact 0 for that (the table-driven log(1) is ~1e-17, and dd*D may be huge)
        const double pp = 1.0 - daily_mortality_zero_density;
        double lp = ::mms::fm::log_nonneg(pp);
        lp = ::mms::fm::sel(pp == 1.0, 0.0, lp);
        return ::mms::fm::sel(pp > 1.0, ::mms::fm::make_nan(), lp);
#endif
    }
#if MMS_STRICT
    return log(das0 / (1.0 - das0));                             // logit(), :533,:571
#else
    const double om = 1.0 - das0;                                // das0 == 1 -> p / 0 = +Inf like R
    const double odds = ::mms::fm::sel(om == 0.0, ::mms::fm::make_inf(),
                                       ::mms::fm::div_fast(das0, ::mms::fm::sel(om == 0.0, 1.0, om)));
    return ::mms::fm::log_nonneg(odds);
#endif
}

// the part that takes the WATER temperature rates (R/sim_pop_dynamics.R:108,114): emergence, density-free link
__device__ __forceinline__ void pop_hour_water(PopHour &ph, int densmod_type, double mdr, double das0)
{
    const double e = pop_emergence_complement(mdr);
    ph.emf = 1.0 - e;
#if MMS_STRICT
    ph.one_m_emf = 1.0 - ph.emf;
#else
    ph.one_m_emf = e;
#endif
    ph.link0 = pop_link0(densmod_type, das0);
}

__device__ __forceinline__ PopHour pop_hour(int densmod_type, double mdr, double efd, double das0, double ds)
{
    PopHour ph;
    pop_hour_air(ph, efd, ds);
    pop_hour_water(ph, densmod_type, mdr, das0);
    return ph;
}

// ---- water temperature (EXTENSION; oracle: mmso_iterate_volume_temperature) ------------------
// One hour of the water body WITH its temperature: the water present at the start of the hour (volume, surface
// `surface` = S(volume), temperature tw) exchanges evaporative, long-wave, solar [and sensible] heat with a linearly
// implicit step, then the volume step of iterate_volume() with the same arithmetic, then the rain that arrives (at
// air temperature) is mixed in by mass.  Plain IEEE arithmetic in both math modes (an optional model, off the
// benchmarked path); returns the new volume, updates tw.
__device__ __forceinline__ double iterate_volume_temperature(const HabitatDev &hb, const WaterTempDev &wt, double volume,
                                                             double surface, double &tw, double evap_rate, double gain,
                                                             double air_t, double cloud_pct, double i_clear)
{
    double tw_pre;
    if (!(volume > 0.0)) {
        tw_pre = air_t;
    } else {
        const double dt = 3600.0, se = wt.sigma_eps;
        const double mass = 1000.0 * volume;
        const double q_evap = 2.26e6 * (evap_rate * surface);
        const double twk = tw + 273.15, tak = air_t + 273.15;
        const double twk2 = twk * twk, tak2 = tak * tak;
        const double q_rad = se * surface * (twk2 * twk2 - tak2 * tak2) * dt;
        const double angstrom = 0.36 + 0.64 * (1.0 - cloud_pct / 100.0);
        const double q_sol = wt.solar_gain * i_clear * angstrom * surface * dt;
        const double q_conv = wt.convection * surface * (air_t - tw) * dt;
        const double k = 4.0 * se * (twk2 * twk) * surface * dt + wt.convection * surface * dt;
        tw_pre = tw + (q_sol - q_evap - q_rad + q_conv) / (mass * 4184.0 + k);
        if (tw_pre < 0.0) tw_pre = 0.0;
        if (tw_pre > 100.0) tw_pre = 100.0;
    }
#if MMS_STRICT
    const double loss = (evap_rate * surface) / 1000.0;
#else
    const double loss = (evap_rate * surface) * kLitC[2];
#endif
    double x = volume - loss;
    x = x - hb.leak_per_hour * volume;
    x = ensure_positive(x);
    double y = x + gain;
    tw = (y > 0.0) ? (x * tw_pre + gain * air_t) / y : air_t;
    y = (y > hb.max_volume) ? hb.max_volume : y;
    return y;
}

// iterate_state(), R/sim_pop_dynamics.R:14-96 (state dependent part)
#if MMS_STRICT
__device__ __forceinline__ void iterate_state(double &aquatic, double &adult, const PopHour &ph,
                                              int densmod_type, double dd_effect, double area_cm2,
                                              double larval_habitat_area)
{
    const double aq = (aquatic != aquatic) ? 0.0 : aquatic;     // :33
    const double ad = (adult != adult) ? 0.0 : adult;           // :34
    const double density = aq / (larval_habitat_area + kEps);   // :42
    const double density_dish = density * area_cm2;             // densmod :549 / :569
    double surv_aq;
    double das_t;
    if (densmod_type == 1) {
        const double loghaz = ph.link0 + dd_effect * density_dish;          // :556-557
        const double daily_mortality = 1.0 - exp(-exp(loghaz));             // :560
        das_t = 1.0 - daily_mortality;                                      // :562
    } else {
        const double lg = ph.link0 + dd_effect * density_dish;              // :574-575
        das_t = 1.0 / (1.0 + exp(-lg));                                     // :537,:578
    }
    surv_aq = r_pow_pos(das_t, kTimestep);                                  // :49
    const double surviving_adult = ad * ph.surv_ad;             // :67
    const double surviving_aquatic = aq * surv_aq;              // :68
    const double new_aquatic = surviving_adult * ph.egg;        // :71
    const double remaining_aquatic = surviving_aquatic * ph.one_m_emf;   // :74
    const double new_adult = surviving_aquatic * ph.emf;        // :77
    aquatic = remaining_aquatic + new_aquatic;                  // :92
    adult = surviving_adult + new_adult + kEps;                 // :93
}
#else
// N settings of one cell at once, stage by stage (mms_fastmath.cuh: why); the scalar form is N = 1
template <int N>
__device__ __forceinline__ void iterate_state_n(double (&aquatic)[N], double (&adult)[N], const PopHour &ph,
                                                int densmod_type, double dd_effect, double area_cm2,
                                                const double (&larval_habitat_area)[N])
{
    double aq[N], ad[N], den[N], density[N], surv_aq[N];
    MMS_EACH(i, N) aq[i] = (aquatic[i] != aquatic[i]) ? 0.0 : aquatic[i];     // :33
    MMS_EACH(i, N) ad[i] = (adult[i] != adult[i]) ? 0.0 : adult[i];           // :34
    MMS_EACH(i, N) den[i] = larval_habitat_area[i] + kEps;
    ::mms::fm::div_fast_n<N>(aq, den, density);                                // :42
    if (densmod_type == 1) {
        // exp(-exp(log(-log p') + dd*D)) = exp(log p' * exp(dd*D)); the reference's
        // 1 - (1 - z) rounding (z < 1.1e-16 -> 0) is kept literally.  dd*D is clamped to
        // +-700 before exp (beyond that z is 0 or 1 either way) so that 0 * Inf cannot arise.
        double u[N], uc[N], e1[N], z[N], das_t[N];
        MMS_EACH(i, N) {
            const double density_dish = density[i] * area_cm2;                 // densmod :549
            u[i] = dd_effect * density_dish;
            const int uhi = __double2hiint(u[i]);
            const bool ubig = ((unsigned)uhi & 0x7fffffffu) >= 0x4085e000u;     // |u| >= 700, Inf, NaN
            uc[i] = ::mms::fm::sel(ubig, (uhi < 0) ? -700.0 : 700.0, u[i]);
        }
        ::mms::fm::exp_core_n<N>(uc, e1);
        MMS_EACH(i, N) e1[i] = ph.link0 * e1[i];
        ::mms::fm::exp_sat_n<N>(e1, z);
        MMS_EACH(i, N) {
            z[i] = ::mms::fm::sel(::mms::fm::is_nan_bits(u[i]), u[i], z[i]);
            das_t[i] = 1.0 - (1.0 - z[i]);
        }
        ::mms::fm::pow_pos_n<N>(das_t, kLitC[3], surv_aq);
    } else {
        // das_t = 1 / (1 + e), e = exp(a), a = -lg;  das_t ^ (1/24) = exp(-log(1 + e) / 24).
        // a is clamped to [-50, 40] for e: below, 1 + e == 1 exactly; above, log(1 + e) == a
        // to double precision, and e = Inf (a > 709.78...) makes the reference's das_t 0.
        double a[N], ac[N], e[N], L[N], arg[N];
        bool hi_side[N], overflow[N];
        MMS_EACH(i, N) {
            const double density_dish = density[i] * area_cm2;                 // densmod :569
            a[i] = -(ph.link0 + dd_effect * density_dish);
        }
        MMS_EACH(i, N) {
            const int ahi = __double2hiint(a[i]);
            const unsigned aabs = (unsigned)ahi & 0x7fffffffu;
            hi_side[i] = ahi >= 0x40440000;                                     // a >= 40 (or +Inf, +NaN)
            const bool lo_side = ahi < 0 && aabs >= 0x40490000u;                // a <= -50 (or -Inf, -NaN)
            ac[i] = ::mms::fm::sel(hi_side[i] || lo_side, hi_side[i] ? 40.0 : -50.0, a[i]);
        }
        ::mms::fm::exp_core_n<N>(ac, e);
        MMS_EACH(i, N) e[i] = 1.0 + e[i];
        ::mms::fm::log_core_n<N>(e, L);
        MMS_EACH(i, N) {
            L[i] = ::mms::fm::sel(hi_side[i], a[i], L[i]);
            // -L / 24 is within [-29.6, 0] whenever the result is used: beyond a = 709.78 the select
            // below wins (exp(a) = Inf -> das_t = 0); a NaN a propagates through exp_core
            const int ahi = __double2hiint(a[i]);
            overflow[i] = ahi > 0x40862e42 || (ahi == 0x40862e42 && (unsigned)__double2loint(a[i]) > 0xfefa39efu);
            arg[i] = kLitC[4] * ::mms::fm::sel(overflow[i], 0.0, L[i]);
        }
        ::mms::fm::exp_core_n<N>(arg, surv_aq);
        MMS_EACH(i, N) {
            surv_aq[i] = ::mms::fm::sel(overflow[i], 0.0, surv_aq[i]);
            surv_aq[i] = ::mms::fm::sel(::mms::fm::is_nan_bits(a[i]), a[i], surv_aq[i]);
        }
    }
    MMS_EACH(i, N) {
        const double surviving_adult = ad[i] * ph.surv_ad;             // :67
        const double surviving_aquatic = aq[i] * surv_aq[i];           // :68
        const double new_aquatic = surviving_adult * ph.egg;           // :71
        const double remaining_aquatic = surviving_aquatic * ph.one_m_emf;   // :74
        const double new_adult = surviving_aquatic * ph.emf;           // :77
        aquatic[i] = remaining_aquatic + new_aquatic;                  // :92
        adult[i] = surviving_adult + new_adult + kEps;                 // :93
    }
}
__device__ __forceinline__ void iterate_state(double &aquatic, double &adult, const PopHour &ph,
                                              int densmod_type, double dd_effect, double area_cm2,
                                              double larval_habitat_area)
{
    double aqv[1] = {aquatic}, adv[1] = {adult};
    const double area[1] = {larval_habitat_area};
    iterate_state_n<1>(aqv, adv, ph, densmod_type, dd_effect, area_cm2, area);
    aquatic = aqv[0];
    adult = adv[0];
}
#endif

}  // namespace MMS_MODE
}  // namespace mms
