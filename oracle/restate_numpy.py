"""Second, independently written restatement of the reference chain, in the
reference's own *vectorised* style (numpy standing in for R vectors).

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no R available).  Its job is
double-entry book-keeping: oracle/mms_oracle.c is a scalar, cell-at-a-time
restatement; this file follows the R code's whole-vector formulation (mask
multiplies, rep(each = 24), time-by-pixel matrices, `for (t in ...)` over rows).
tests/ compare the two and tests/golden/ holds vectors generated from this
file (tests/golden/make_golden.py).

Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# base-R helpers
# ---------------------------------------------------------------------------


def r_approx(x, y, xout):
    """stats::approx(x, y, xout)$y, method linear, rule 1 (NA outside)."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    xout = np.atleast_1d(np.asarray(xout, float))
    n = len(x)
    out = np.full(xout.shape, np.nan)
    ok = (xout >= x[0]) & (xout <= x[-1])
    v = xout[ok]
    i = np.searchsorted(x, v, side="right") - 1      # largest i with x[i] <= v
    i = np.minimum(i, n - 2)
    j = i + 1
    with np.errstate(invalid="ignore", divide="ignore"):
        lin = y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]))
    res = np.where(v == x[j], y[j], np.where(v == x[i], y[i], lin))
    out[ok] = res
    return out


def r_pow(x, y):
    """R's `^` for doubles (y == 2 -> x*x; x == 0, y > 0 -> 0; else libm pow)."""
    x = np.asarray(x, float)
    if y == 2:
        return x * x
    with np.errstate(invalid="ignore", divide="ignore"):
        p = np.power(x, y)
    if y > 0:
        p = np.where(x == 0, 0.0, p)
    return p


def r_mean(x):
    """base-R mean(): long double sum, then one correction pass."""
    x = np.asarray(x, float)
    n = len(x)
    xl = x.astype(np.longdouble)
    s = np.longdouble(0)
    for v in xl:                      # sequential like R (no pairwise summation)
        s += v
    s /= n
    t = np.longdouble(0)
    for v in xl:
        t += (v - s)
    s += t / n
    return float(s)


def ensure_positive(x):               # R/utils.R:159-161
    with np.errstate(invalid="ignore"):
        return x * (x > 0).astype(float)


def enforce_max(x, mx):               # R/utils.R:164-168
    x = np.array(x, float, copy=True)
    x[x > mx] = mx
    return x


# ---------------------------------------------------------------------------
# S: R/spline_interpolation.R, R/sim_microclimate.R:7-64
# ---------------------------------------------------------------------------

FORCING = ("tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax", "rainfall_daily")


def spline_seasonal(values, datenum, datenum_predict):
    datenum = np.asarray(datenum, float)
    datenum_predict = np.asarray(datenum_predict, float)
    if np.any((datenum_predict > datenum.max()) | (datenum_predict < datenum.min())):
        raise ValueError("extrapolation: GAM branch (R/spline_interpolation.R:30-112) not restated")
    return r_approx(datenum, values, datenum_predict)          # :25-27


def daily_from_monthly_climate(monthly: dict, mid_day, day_predict) -> dict:
    m = dict(monthly)
    m["rainfall_daily"] = np.log1p(np.asarray(m["rainfall_daily"], float))   # :33
    daily = {k: spline_seasonal(m[k], mid_day, day_predict) for k in FORCING}   # :46-48
    daily["rainfall_daily"] = np.maximum(0, np.expm1(daily["rainfall_daily"]))  # :61
    return daily


# ---------------------------------------------------------------------------
# M: R/sim_microclimate.R:126-566
# ---------------------------------------------------------------------------


def interpolate_daily_air_temp(daily_max, daily_min, time_sunrise=6, time_sunset=18, time_daily_max=13):
    temp_max_day = np.asarray(daily_max, float)
    temp_min_day = np.asarray(daily_min, float)
    time_sunrise = time_sunrise * 100
    time_sunset = time_sunset * 100
    time_daily_max = time_daily_max * 100
    n_days = len(temp_max_day)
    times = np.tile(100.0 * np.arange(1, 25), n_days)                              # :201
    time_reference = (time_daily_max - time_sunrise) / 2 + time_sunrise            # :204
    tau = 3 / (2400 - time_sunset + time_sunrise)                                  # :207
    Z_night = np.sin((360 * (time_sunset - time_reference) / (2 * (time_daily_max - time_sunrise))) / 57.29577)
    temp_min_tomorrow_day = np.concatenate([temp_min_day[1:], temp_min_day[-1:]])  # :213
    A_day = (temp_max_day - temp_min_day) / 2                                      # :217
    temp_sunset_day = A_day * Z_night + temp_min_day + A_day                       # :220
    temp_sunset_yesterday_day = np.concatenate([temp_sunset_day[:1], temp_sunset_day[:-1]])  # :223
    rep = lambda a: np.repeat(a, 24)
    temp_min_hour = rep(temp_min_day)
    A_hour = rep(A_day)
    temp_sunset_hour = rep(temp_sunset_day)
    temp_min_tomorrow_hour = rep(temp_min_tomorrow_day)
    temp_sunset_yesterday_hour = rep(temp_sunset_yesterday_day)
    after_sunrise = (times > time_sunrise).astype(float)                           # :235
    temp_min_next_hour = temp_min_hour * (1 - after_sunrise) + temp_min_tomorrow_hour * after_sunrise
    before_sunset = (times < time_sunset).astype(float)                            # :242
    temp_sunset_last_hour = temp_sunset_hour * (1 - before_sunset) + temp_sunset_yesterday_hour * before_sunset
    is_light = ((times >= time_sunrise) & (times <= time_sunset)).astype(float)    # :247
    Z_hour_light = np.sin((360 * (times - time_reference) / (2 * (time_daily_max - time_sunrise))) / 57.29577)
    temp_hour_light = A_hour * Z_hour_light + temp_min_hour + A_hour               # :256
    time_since_ss = times - time_sunset + 2400 * (times < 1800)                    # :260
    E = tau * time_since_ss
    temp_hour_night = (temp_sunset_last_hour - temp_min_next_hour) / np.exp(E) + temp_min_next_hour
    temp_hour = temp_hour_light * is_light + temp_hour_night * (1 - is_light)      # :265
    temp_midnight_day = temp_hour[times == 2400]                                   # :271
    temp_midnight_yesterday_day = np.concatenate([temp_midnight_day[:1], temp_midnight_day[:-1]])
    temp_midnight_yesterday_hour = rep(temp_midnight_yesterday_day)
    temp_hour_predawn = temp_min_hour + (temp_midnight_yesterday_hour - temp_min_hour) * (1 - times / time_sunrise)
    is_predawn = times < time_sunrise
    temp_hour[is_predawn] = temp_hour_predawn[is_predawn]                          # :277
    return temp_hour


def vsine_weights(times, time_sunrise=6, time_inflection=13):
    times = np.asarray(times, float)
    mask_mn_to_sr = (times <= time_sunrise).astype(float)
    rel_mn_to_sr = (time_sunrise - times) / time_sunrise
    weight_mn_to_sr = (1 - 0.5 * np.maximum(0, rel_mn_to_sr)) * mask_mn_to_sr
    mask_sr_to_inf = ((times > time_sunrise) & (times <= time_inflection)).astype(float)
    rel_sr_to_inf = (times - time_sunrise) / (time_inflection - time_sunrise)
    weight_sr_to_inf = (1 - np.minimum(1, np.maximum(0, rel_sr_to_inf))) * mask_sr_to_inf
    mask_inf_to_mn = (times > time_inflection).astype(float)
    rel_inf_to_mn = (times - time_inflection) / (24 - time_inflection)
    weight_inf_to_mn = 0.5 * np.minimum(1, np.maximum(0, rel_inf_to_mn)) * mask_inf_to_mn
    return weight_mn_to_sr + weight_sr_to_inf + weight_inf_to_mn


def vsine_interpolate_daily(daily_max, daily_min, time_sunrise=6, time_inflection=13, sunrise_value="max"):
    max_day = np.asarray(daily_max, float)
    min_day = np.asarray(daily_min, float)
    n_days = len(max_day)
    times = np.tile(np.arange(1.0, 25.0), n_days)
    weight = vsine_weights(times, time_sunrise, time_inflection)
    if sunrise_value == "min":
        weight = 1 - weight
        sunval_day, infval_day = min_day, max_day
    else:
        sunval_day, infval_day = max_day, min_day
    infval_hour = np.repeat(infval_day, 24)
    infval_yesterday_hour = np.repeat(np.concatenate([infval_day[:1], infval_day[:-1]]), 24)
    before_sunrise = (times < time_sunrise).astype(float)
    infval_last_hour = infval_hour * (1 - before_sunrise) + infval_yesterday_hour * before_sunrise
    sunval_hour = np.repeat(sunval_day, 24)
    sunval_tomorrow_hour = np.repeat(np.concatenate([sunval_day[1:], sunval_day[-1:]]), 24)
    after_inflection = (times > time_inflection).astype(float)
    sunval_next_hour = sunval_hour * (1 - after_inflection) + sunval_tomorrow_hour * after_inflection
    if sunrise_value == "min":
        diff_hour = infval_last_hour - sunval_next_hour
        return sunval_next_hour + weight * diff_hour
    diff_hour = sunval_next_hour - infval_last_hour
    return infval_last_hour + weight * diff_hour


def hourly_from_daily_climate_ambient(daily: dict) -> dict:
    air = interpolate_daily_air_temp(daily["tmax"], daily["tmin"])
    return {
        "air_temperature": air,
        "humidity": vsine_interpolate_daily(daily["rhmax"], daily["rhmin"], sunrise_value="max"),
        "cloudcover": vsine_interpolate_daily(daily["ccmax"], daily["ccmin"], sunrise_value="max"),
        "windspeed": vsine_interpolate_daily(daily["wsmax"], daily["wsmin"], sunrise_value="min"),
        "water_temperature": air.copy(),                                           # :163
        "rainfall": np.repeat(np.asarray(daily["rainfall_daily"], float) / 24, 24),  # :132
    }


# ---------------------------------------------------------------------------
# W: R/sim_water.R
# ---------------------------------------------------------------------------


def cone_depth_to_volume(depth):
    radius = depth * np.sin(np.pi / 2)
    return np.pi * radius * radius * depth / 3


def cone_volume_to_surface(volume):
    return 3.046474 * r_pow(volume, 0.666667)


def saturation_vapour_pressure_kpa(t):
    return 0.61121 * np.exp((18.678 - t / 234.5) * (t / (257.14 + t)))


def air_pressure_kpa(alt):
    return 101.325 - 1.2 * alt / 100


def evaporation_rate_kg_h(temperature_c, relative_humidity, altitude_m, windspeed_m_s, surface_area_m2):
    theta = (25 + 19 * windspeed_m_s)
    p_ws = saturation_vapour_pressure_kpa(temperature_c)
    p_a = air_pressure_kpa(altitude_m)
    xs = 0.62198 * p_ws / (p_a - p_ws)
    p_w = p_ws * (relative_humidity / 100)
    x = 0.62198 * p_w / (p_a - p_w)
    return theta * surface_area_m2 * (xs - x)


def habitat_surface(hb: dict, volume):
    """cone = reference; cylinder = EXTENSION (constant free surface while wet)."""
    if hb.get("shape", 0) == 0:
        return cone_volume_to_surface(volume)
    r = hb["radius"]
    return (np.pi * (r * r)) * (np.asarray(volume) > 0).astype(float)


def habitat_max_volume(hb: dict):
    if hb.get("shape", 0) == 0:
        return cone_depth_to_volume(hb.get("max_depth", 1.0))
    r = hb["radius"]
    return np.pi * (r * r) * hb["max_depth"]


def simulate_ephemeral_habitat_vectorised(rainfall_matrix, air_temperature_matrix, humidity_matrix,
                                          windspeed_matrix, altitude_vector, hb: dict | None = None,
                                          burnin_years=0):
    """R/sim_water.R:216-301 (matrices [n_times, n_pixels]).  hb carries the
    cone parameters (initial_volume, max_depth, inflow_multiplier) plus the
    EXTENSION fields (shape, radius, leak_per_hour)."""
    hb = dict(shape=0, max_depth=1.0, radius=1.0, inflow_multiplier=1.0, leak_per_hour=0.0,
              initial_volume=0.0) | (hb or {})
    n_times_obs = rainfall_matrix.shape[0]
    first_year_index = np.arange(365 * 24)
    index = np.concatenate([np.tile(first_year_index, burnin_years), np.arange(n_times_obs)])
    n_pixels = rainfall_matrix.shape[1]
    n_times_sim = len(index)
    rainfall = rainfall_matrix[index, :]
    air_temperature = air_temperature_matrix[index, :]
    relative_humidity = humidity_matrix[index, :]
    windspeed = windspeed_matrix[index, :]
    altitude = np.zeros((n_times_sim, n_pixels)) + np.asarray(altitude_vector, float)[None, :]
    max_volume = habitat_max_volume(hb)
    catchment_area = hb["inflow_multiplier"] * habitat_surface(hb, max_volume)
    loss_kg_h_m2 = evaporation_rate_kg_h(air_temperature, relative_humidity, altitude, windspeed, 1)
    volumes = np.full((n_times_sim, n_pixels), np.nan)
    volume = np.full(n_pixels, float(hb["initial_volume"]))
    leak = hb["leak_per_hour"]
    for t in range(n_times_sim):
        surface_area_m2 = habitat_surface(hb, volume)
        loss = loss_kg_h_m2[t, :] * surface_area_m2 / 1000
        x = volume - loss
        x = x - leak * volume                      # EXTENSION; leak = 0 -> reference
        volume = ensure_positive(x)
        gain = catchment_area * (rainfall[t, :] / 1000)
        volume = enforce_max(volume + gain, max_volume)
        volumes[t, :] = volume
    volumes = volumes[-n_times_obs:, :]
    return {"water_surface_area": habitat_surface(hb, volumes), "water_volume": volumes,
            "wet": (volumes > 0).astype(np.uint8)}


# ---------------------------------------------------------------------------
# L: R/lifehistory_estimation_functions.R runtime closures
# ---------------------------------------------------------------------------


def spline_eval(tab: dict, u):
    """splinefun(method='fmm') evaluation, vectorised, fresh interval search per
    point (R carries the interval between points; the two agree to rounding
    because the spline is continuous)."""
    x = tab["x"]
    u = np.asarray(u, float)
    i = np.clip(np.searchsorted(x, u, side="right") - 1, 0, len(x) - 1)
    dx = u - x[i]
    return tab["y"][i] + dx * (tab["b"][i] + dx * (tab["c"][i] + dx * tab["d"][i]))


def positive_spline_eval(tab: dict, u):
    return ensure_positive(spline_eval(tab, u))


def interp_surface(tab: dict, tx, ty):
    """fields::interp.surface (restated)."""
    x, y, z = tab["x"], tab["y"], tab["z"]            # z[ix, iy]
    nx, ny = len(x), len(y)
    lx = r_approx(x, np.arange(1.0, nx + 1), tx)
    ly = r_approx(y, np.arange(1.0, ny + 1), ty)
    bad = np.isnan(lx) | np.isnan(ly)
    lx = np.where(bad, 1.0, lx)
    ly = np.where(bad, 1.0, ly)
    lx1 = np.floor(lx)
    ly1 = np.floor(ly)
    ex = lx - lx1
    ey = ly - ly1
    ex[lx1 == nx] = 1
    ey[ly1 == ny] = 1
    lx1[lx1 == nx] = nx - 1
    ly1[ly1 == ny] = ny - 1
    i = lx1.astype(int) - 1
    j = ly1.astype(int) - 1
    out = (z[i, j] * (1 - ex) * (1 - ey) + z[i + 1, j] * ex * (1 - ey) +
           z[i, j + 1] * (1 - ex) * ey + z[i + 1, j + 1] * ex * ey)
    return np.where(bad, np.nan, out)


def densmod_apply(dm: dict, surv_prob, density):
    density_dish = density * dm["surface_area_cm2"]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if dm["type"] == "cox_ph":
            daily_mortality_zero_density = 1 - surv_prob
            loghaz0 = np.log(-np.log(1 - daily_mortality_zero_density))
            loghaz = loghaz0 + dm["dd_effect"] * density_dish
            daily_mortality = 1 - np.exp(-np.exp(loghaz))
            return 1 - daily_mortality
        logit0 = np.log(surv_prob / (1 - surv_prob))
        lg = logit0 + dm["dd_effect"] * density_dish
        return 1 / (1 + np.exp(-lg))


# ---------------------------------------------------------------------------
# P: R/sim_pop_dynamics.R
# ---------------------------------------------------------------------------


def iterate_state(aquatic, adult, dm, larval_habitat_area, mdr, efd, das_zerodensity, ds, timestep=1 / 24):
    epsilon = np.sqrt(np.finfo(float).eps)
    aquatic = np.where(np.isnan(aquatic), 0.0, aquatic)
    adult = np.where(np.isnan(adult), 0.0, adult)
    aquatic_density = aquatic / (larval_habitat_area + epsilon)
    das_t = densmod_apply(dm, das_zerodensity, aquatic_density)
    aquatic_survival_prob = r_pow(das_t, timestep)
    adult_survival_prob = r_pow(ds, timestep)
    emergence_rate = timestep * mdr
    emergence_fraction = 1 - np.exp(-emergence_rate)
    egg_laying_rate = efd * timestep
    surviving_adult = adult * adult_survival_prob
    surviving_aquatic = aquatic * aquatic_survival_prob
    new_aquatic = surviving_adult * egg_laying_rate
    remaining_aquatic = surviving_aquatic * (1 - emergence_fraction)
    new_adult = surviving_aquatic * emergence_fraction
    return remaining_aquatic + new_aquatic, surviving_adult + new_adult + epsilon


def simulate_population_vectorised(larval_habitat_area_matrix, mdr_matrix, efd_matrix,
                                   das_zerodensity_matrix, ds_matrix, dm, initial_adult=100,
                                   initial_aquatic=100, burnin_years=1):
    n_times_obs = larval_habitat_area_matrix.shape[0]
    index = np.concatenate([np.tile(np.arange(365 * 24), burnin_years), np.arange(n_times_obs)])
    n_pixels = larval_habitat_area_matrix.shape[1]
    n_times_sim = len(index)
    lha = larval_habitat_area_matrix[index, :]
    mdr = mdr_matrix[index, :]
    efd = efd_matrix[index, :]
    das0 = das_zerodensity_matrix[index, :]
    ds = ds_matrix[index, :]
    adult_states = np.full((n_times_sim, n_pixels), np.nan)
    aquatic_states = np.full((n_times_sim, n_pixels), np.nan)
    adult = np.full(n_pixels, float(initial_adult))
    aquatic = np.full(n_pixels, float(initial_aquatic))
    for t in range(n_times_sim):
        aquatic, adult = iterate_state(aquatic, adult, dm, lha[t, :], mdr[t, :], efd[t, :], das0[t, :], ds[t, :])
        aquatic_states[t, :] = aquatic
        adult_states[t, :] = adult
    return {"aquatic": aquatic_states[-n_times_obs:, :], "adult": adult_states[-n_times_obs:, :]}


# ---------------------------------------------------------------------------
# fused chain for a few cells (tc_to_vectors from interpolate_daily_climate on)
# ---------------------------------------------------------------------------

VAR_NAMES = ("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds",
             "larval_habitat_area", "water_volume", "air_temperature", "humidity",
             "windspeed", "rainfall")


def run_chain(monthly, mid_day, day_predict, altitude, habitats, dm, tables, period_of_day, n_periods,
              burnin_years_water=0, burnin_years_pop=1, initial_adult=100, initial_aquatic=100):
    """monthly [9][M][C]; returns hourly [S][12][H][C], mean [S][12][P][C],
    wet [S][H][C], wet_hours [S][P][C]."""
    monthly = np.asarray(monthly, float)
    _, M, C = monthly.shape
    D = len(day_predict)
    H = 24 * D
    hourly_cells = []
    for c in range(C):
        daily = daily_from_monthly_climate({k: monthly[v, :, c] for v, k in enumerate(FORCING)},
                                           mid_day, day_predict)
        hourly_cells.append(hourly_from_daily_climate_ambient(daily))
    mat = lambda k: np.stack([hc[k] for hc in hourly_cells], axis=1)   # [H, C]
    air, rh, wind, rain, wt = (mat(k) for k in ("air_temperature", "humidity", "windspeed",
                                                "rainfall", "water_temperature"))
    ds_tab = dict(tables["ds"])
    ds_tab["z"] = np.asarray(tables["ds"]["z"]).reshape(len(ds_tab["y"]), len(ds_tab["x"])).T  # z[ix, iy]
    mdr = np.stack([positive_spline_eval(tables["mdr"], wt[:, c]) for c in range(C)], axis=1)
    efd = np.stack([positive_spline_eval(tables["efd"], air[:, c]) for c in range(C)], axis=1)
    das0 = np.stack([positive_spline_eval(tables["das"], wt[:, c]) for c in range(C)], axis=1)
    ds = np.stack([interp_surface(ds_tab, air[:, c], rh[:, c]) for c in range(C)], axis=1)
    S = len(habitats)
    hourly = np.empty((S, len(VAR_NAMES), H, C))
    wet = np.empty((S, H, C), dtype=np.uint8)
    for s, hb in enumerate(habitats):
        w = simulate_ephemeral_habitat_vectorised(rain, air, rh, wind, altitude, hb, burnin_years_water)
        pop = simulate_population_vectorised(w["water_surface_area"], mdr, efd, das0, ds, dm,
                                             initial_adult, initial_aquatic, burnin_years_pop)
        for v, arr in enumerate((pop["adult"], pop["aquatic"], mdr, efd, das0, ds,
                                 w["water_surface_area"], w["water_volume"], air, rh, wind, rain)):
            hourly[s, v] = arr
        wet[s] = w["wet"]
    P = int(n_periods)
    mean = np.full((S, len(VAR_NAMES), P, C), np.nan)
    wet_hours = np.zeros((S, P, C), dtype=np.int32)
    hour_period = np.repeat(np.asarray(period_of_day), 24)
    for p in range(P):
        sel = hour_period == p
        if not sel.any():
            continue
        for s in range(S):
            wet_hours[s, p] = wet[s][sel].sum(axis=0)
            for v in range(len(VAR_NAMES)):
                for c in range(C):
                    mean[s, v, p, c] = r_mean(hourly[s, v, sel, c])
    return {"hourly": hourly, "wet": wet, "mean": mean, "wet_hours": wet_hours}


# ---- T: TerraClimate variables -> monthly forcing (R/terraclimate.R:581-643), vectorised like the R code ----

def vapprs(db):
    """NicheMapR::VAPPRS (third party, restated: Goff-Gratch; over ice at or below the triple point)."""
    t = np.asarray(db, dtype=np.float64) + 273.16
    with np.errstate(all="ignore"):
        ice = -9.09718 * (273.16 / t - 1.) - 3.56654 * np.log10(273.16 / t) + .876793 * (1. - t / 273.16) + np.log10(6.1071)
        water = (-7.90298 * (373.16 / t - 1.) + 5.02808 * np.log10(373.16 / t)
                 - 1.3816E-07 * (np.power(10., 11.344 * (1. - t / 373.16)) - 1.)
                 + 8.1328E-03 * (np.power(10., -3.49149 * (373.16 / t - 1.)) - 1.) + np.log10(1013.246))
    loge = np.where(t <= 273.16, ice, water)
    return np.power(10., loge) * 100                 # R_pow(10, y) is libm pow for these y


def process_terraclimate_pixel_vars(terraclimate, clear_sky, month_id, n_days):
    """terraclimate [6][M][C] (tmax, tmin, ppt, ws, vpd, srad), clear_sky [12][C] -> monthly [9][M][C]."""
    tmax, tmin, ppt, ws, vpd, srad = (np.asarray(a, dtype=np.float64) for a in terraclimate)
    clear = np.asarray(clear_sky, dtype=np.float64)[np.asarray(month_id) - 1]          # :600-603
    tmean = (tmax + tmin) / 2                                                          # :607
    vp = vapprs(tmean) * 100 / 100 - vpd * 1000                                        # R/utils.R:64-78

    def relative_humidity(temperature):                                                # R/utils.R:80-89
        rh = (vp / vapprs(temperature)) * 100
        return np.maximum(np.minimum(rh, 100), 0.01)

    def cloud_cover(multiplier):                                                       # R/utils.R:141-147
        cloud = (1 - srad / clear) * 100 * multiplier
        return np.maximum(np.minimum(cloud, 100), 0)

    wsmax = ws * r_pow(2 / 10, 0.15)                                                   # R/utils.R:132
    n = np.asarray(n_days, dtype=np.float64)[:, None]
    return np.stack([tmin, tmax, relative_humidity(tmax), relative_humidity(tmin), wsmax * 0.1, wsmax,
                     cloud_cover(0.8), cloud_cover(1.2), ppt / n])
