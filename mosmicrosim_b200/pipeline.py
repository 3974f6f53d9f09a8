"""Host-side mirror of the reference's interface for the hot path: the same
function names, argument names, defaults, error behaviour and output schema as
R/sim_water.R, R/sim_pop_dynamics.R and R/pipeline_functions.R, with the
arithmetic done by libmmsgpu on the GPU.  (R is not installed in this image,
so this Python layer stands where the R functions of r-binding/R/gpu_backend.R
stand in a real installation; both sit on the same C ABI.)

R tibbles become plain dicts of numpy arrays; R matrices are numpy [n_times,
n_pixels] arrays (stored column-major before they cross the ABI, like R).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import backend, dates
from .lifehistory import Lifehistory, SurvDensmod

SPECIES = ("An. gambiae", "An. stephensi")
FORCING_VARS = ("tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax", "rainfall_daily")
VECTOR_VARIABLES = ("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds")   # R/pipeline_functions.R:1041-1048


def _match_arg(value, choices, name):
    if isinstance(value, (tuple, list)):
        value = value[0]
    if value not in choices:
        raise ValueError(f"'{name}' should be one of " + ", ".join(repr(c) for c in choices))
    return value


# ---------------------------------------------------------------------------
# numeric cores (matrix in / matrix out)
# ---------------------------------------------------------------------------

def simulate_ephemeral_habitat_vectorised(rainfall_matrix, air_temperature_matrix, humidity_matrix,
                                          windspeed_matrix, altitude_vector=None, initial_volume=0,
                                          burnin_years=0, max_cone_depth=1, inflow_multiplier=1, *,
                                          shape="cone", radius=1.0, leak_per_hour=0.0,
                                          math_mode="fast", device=0):
    """simulate_ephemeral_habitat_vectorised(), R/sim_water.R:216-301.  Returns
    {"water_surface_area": [n_times, n_pixels]}.  The keyword-only arguments are
    extensions (cylinder / leakage; math mode; device)."""
    rainfall_matrix = np.asarray(rainfall_matrix, dtype=np.float64)
    if altitude_vector is None:
        altitude_vector = np.zeros(rainfall_matrix.shape[1])
    hb = backend.make_habitat(shape={"cone": 0, "cylinder": 1}[shape], max_depth=max_cone_depth, radius=radius,
                              inflow_multiplier=inflow_multiplier, leak_per_hour=leak_per_hour,
                              initial_volume=initial_volume)
    res = backend.water_vectorised(rainfall_matrix, air_temperature_matrix, humidity_matrix, windspeed_matrix,
                                   altitude_vector, hb, burnin_years=burnin_years, math_mode=math_mode,
                                   device=device)
    return {"water_surface_area": res["water_surface_area"]}


def simulate_population_vectorised(larval_habitat_area_matrix, mdr_matrix, efd_matrix, das_zerodensity_matrix,
                                   ds_matrix, das_densmod_function: SurvDensmod, initial_adult=100,
                                   initial_aquatic=100, burnin_years=1, *, math_mode="fast", device=0):
    """simulate_population_vectorised(), R/sim_pop_dynamics.R:208-284.  Returns
    {"aquatic": ..., "adult": ...} as [n_times, n_pixels] matrices."""
    dm = backend.make_densmod(das_densmod_function.type, das_densmod_function.dd_effect,
                              das_densmod_function.surface_area_cm2)
    return backend.population_vectorised(larval_habitat_area_matrix, mdr_matrix, efd_matrix,
                                         das_zerodensity_matrix, ds_matrix, dm, initial_adult=initial_adult,
                                         initial_aquatic=initial_aquatic, burnin_years=burnin_years,
                                         math_mode=math_mode, device=device)


def simulate_ephemeral_habitat(hourly_climate, altitude, initial_volume=0, burnin_years=0, max_cone_depth=1,
                               inflow_multiplier=1, *, shape="cone", radius=1.0, leak_per_hour=0.0,
                               math_mode="fast", device=0):
    """simulate_ephemeral_habitat(), R/sim_water.R:136-205: one pixel; hourly_climate is a mapping with
    the columns rainfall, air_temperature, humidity, windspeed (vectors of n_times).  Returns the vector
    of surface areas.  Unlike the vectorised function the burn-in repeats the whole series (:151-152)."""
    cols = [np.asarray(hourly_climate[k], dtype=np.float64).reshape(-1, 1)
            for k in ("rainfall", "air_temperature", "humidity", "windspeed")]
    hb = backend.make_habitat(shape={"cone": 0, "cylinder": 1}[shape], max_depth=max_cone_depth, radius=radius,
                              inflow_multiplier=inflow_multiplier, leak_per_hour=leak_per_hour,
                              initial_volume=initial_volume)
    res = backend.water_vectorised(*cols, np.array([float(altitude)]), hb, burnin_years=burnin_years,
                                   math_mode=math_mode, device=device, whole_series_burnin=True)
    return res["water_surface_area"][:, 0]


def simulate_population(hourly_lifehistory, das_densmod_function: SurvDensmod, initial_adult=100,
                        initial_aquatic=100, burnin_years=1, *, math_mode="fast", device=0):
    """simulate_population(), R/sim_pop_dynamics.R:138-196: one pixel; hourly_lifehistory is a mapping with
    the columns larval_habitat_area, mdr, efd, das_zerodensity, ds.  Returns {"aquatic", "adult"} vectors.
    The burn-in repeats the whole series (:147-148)."""
    cols = [np.asarray(hourly_lifehistory[k], dtype=np.float64).reshape(-1, 1)
            for k in ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")]
    dm = backend.make_densmod(das_densmod_function.type, das_densmod_function.dd_effect,
                              das_densmod_function.surface_area_cm2)
    res = backend.population_vectorised(*cols, dm, initial_adult=initial_adult, initial_aquatic=initial_aquatic,
                                        burnin_years=burnin_years, math_mode=math_mode, device=device,
                                        whole_series_burnin=True)
    return {"aquatic": res["aquatic"][:, 0], "adult": res["adult"][:, 0]}


# ---------------------------------------------------------------------------
# the pipeline (tc_to_vectors() after terraclimate_to_monthly_climate())
# ---------------------------------------------------------------------------

def _monthly_array(pixel_monthly_climate):
    m = pixel_monthly_climate
    start, end = np.asarray(m["start"], "datetime64[D]"), np.asarray(m["end"], "datetime64[D]")
    M = len(start)
    cols = []
    for v in FORCING_VARS:
        a = np.asarray(m[v], dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        if a.shape[0] != M:
            raise ValueError(f"monthly variable {v} must have one row per month")
        cols.append(a)
    monthly = np.ascontiguousarray(np.stack(cols, axis=0))            # [9][M][C]
    altitude = np.atleast_1d(np.asarray(m["altitude"], dtype=np.float64))
    return start, end, monthly, altitude


def fill_na_knots(monthly, start, end):
    """What stats::approx() does with NA knots, done to the knots: approx(x, y, xout) drops the (x, y) pairs with an NA
    (na.rm = TRUE, R/spline_interpolation.R:25-27), i.e. it interpolates linearly between the neighbouring knots that
    are left -- the same line as through an NA knot REPLACED by the linear interpolation of its two non-NA neighbours
    (approx1()'s formula at the knot's own date; rainfall in the log1p space it is interpolated in,
    R/sim_microclimate.R:33,61).  The device library refuses NA knots (MMS_ERR_NONFINITE: its kernels read the two
    knots of a day without a search); this host-side pass makes such tiles acceptable with the reference's results
    to rounding (<= 1e-12 relative; tests/test_host.py).  NA runs at the start or the end of a series cannot be filled
    (the reference gives NA days there, not the GAM) and stay NA: the library still refuses those pixels.
    monthly: [9][M][C] (FORCING_VARS order); returns (filled copy, number of knots filled)."""
    a = np.array(monthly, dtype=np.float64, copy=True)
    x = dates.month_mid_days(start, end).astype(np.float64)
    M = a.shape[1]
    rain = FORCING_VARS.index("rainfall_daily")
    a[rain] = np.log1p(a[rain])
    missing = np.isnan(a)
    n_filled = 0
    if missing.any():
        idx = np.arange(M)[None, :, None]
        prev = np.maximum.accumulate(np.where(missing, -1, idx), axis=1)                      # last non-NA knot at or before m
        nxt = np.minimum.accumulate(np.where(missing, M, idx)[:, ::-1], axis=1)[:, ::-1]       # first non-NA knot at or after m
        fill = missing & (prev >= 0) & (nxt < M)
        pi, ni = np.clip(prev, 0, M - 1), np.clip(nxt, 0, M - 1)
        yi, yj = np.take_along_axis(a, pi, axis=1), np.take_along_axis(a, ni, axis=1)
        xi, xj = x[pi], x[ni]
        with np.errstate(invalid="ignore", divide="ignore"):
            y = yi + (yj - yi) * ((x[None, :, None] - xi) / (xj - xi))                        # approx1()
        a[fill] = y[fill]
        n_filled = int(fill.sum())
    back = np.expm1(a[rain])
    a[rain] = np.where(np.isnan(np.asarray(monthly[rain], dtype=np.float64)), back, monthly[rain])   # untouched knots keep their bits
    return a, n_filled


def interpolate_daily_climate(pixel_monthly_climate, dates_predict=None, *, on_extrapolation="error", math_mode="fast",
                              device=0):
    """interpolate_daily_climate() (R/pipeline_functions.R:258-275): monthly -> daily for every
    pixel of the tile.  Returns the tile's `daily_climate` tibbles side by side: a dict with
    `date` [D], the pixel columns carried through, and one [D, C] array per daily variable."""
    start, end, monthly, altitude = _monthly_array(pixel_monthly_climate)
    on_extrapolation = _match_arg(on_extrapolation, ("error", "clip"), "on_extrapolation")
    if dates_predict is None:
        dates_predict = dates.default_dates_predict(start, end)
    if on_extrapolation == "clip":
        # a half-day overhang of the default dates over the knots (30-day first month, non-leap February last month)
        # would send the reference into its mgcv::gam branch, which the library refuses (MMS_ERR_EXTRAPOLATION)
        dates_predict, dropped = dates.clip_dates_predict(start, end, dates_predict)
        if dropped:
            warnings.warn(f"{dropped} prediction day(s) outside the monthly mid-dates dropped "
                          "(the reference would extrapolate them with an mgcv::gam)")
    mid_day, day_predict = dates.datenums(start, end, dates_predict)
    daily = backend.daily_from_monthly(monthly, mid_day, day_predict, math_mode=math_mode, device=device)
    out = {"date": np.asarray(dates_predict, "datetime64[D]"), "altitude": altitude,
           "longitude": pixel_monthly_climate.get("longitude"), "latitude": pixel_monthly_climate.get("latitude")}
    for v, name in enumerate(FORCING_VARS):
        out[name] = daily[v]
    return out


def simulate_hourly_microclimate(pixel_daily_climate, *, math_mode="fast", device=0):
    """simulate_hourly_microclimate() (R/pipeline_functions.R:284-301): daily -> hourly ambient
    microclimate.  Returns `date` [H], `hour` [H] (1..24) and one [H, C] array per hourly variable
    (air_temperature, humidity, cloudcover, windspeed, water_temperature, rainfall)."""
    daily = np.stack([np.asarray(pixel_daily_climate[name], dtype=np.float64) for name in FORCING_VARS], axis=0)
    hourly = backend.hourly_from_daily(daily, math_mode=math_mode, device=device)
    date = np.asarray(pixel_daily_climate["date"], "datetime64[D]")
    out = {"date": np.repeat(date, 24), "hour": np.tile(np.arange(1, 25), len(date)),
           "altitude": pixel_daily_climate.get("altitude"), "longitude": pixel_daily_climate.get("longitude"),
           "latitude": pixel_daily_climate.get("latitude")}
    out.update(hourly)
    return out


def simulate_hourly_lifehistory(pixel_hourly_conditions, lifehistory_functions, species=SPECIES,
                                fast_ds_temp_humid=True, *, math_mode="fast", device=0):
    """simulate_hourly_lifehistory() (R/pipeline_functions.R:634-683): mdr, efd, das_zerodensity, ds
    from the hourly conditions (water_temperature, air_temperature, humidity, water_surface_area).
    Only the bilinear ds emulator (fast_ds_temp_humid = TRUE, the reference default) exists here."""
    species = _match_arg(species, SPECIES, "species")
    if not fast_ds_temp_humid:
        raise NotImplementedError("the GAM-predict form of ds_temp_humid is not available on the GPU; "
                                  "fast_ds_temp_humid must be TRUE (the reference default)")
    lh: Lifehistory = lifehistory_functions[{"An. gambiae": "An_gambiae", "An. stephensi": "An_stephensi"}[species]]
    c = pixel_hourly_conditions
    res = backend.lifehistory_vectorised(c["water_temperature"], c["air_temperature"], c["humidity"], lh.tables(),
                                         math_mode=math_mode, device=device)
    out = {k: c.get(k) for k in ("date", "hour", "longitude", "latitude")}
    out.update(res)
    out["larval_habitat_area"] = c.get("water_surface_area")      # R/sim_pop_dynamics.R:123
    out["species"] = species
    out["das_densmod_function"] = lh.das_densmod
    return out


def simulate_hourly_conditions(pixel_hourly_microclimate, model_water_temperature=False,
                               water_shade_proportion=1, *, math_mode="fast", device=0):
    """simulate_hourly_conditions(), R/pipeline_functions.R:319-538: adds water_surface_area to the hourly
    microclimate of every pixel (the reference's cone, depth 1, no burn-in: :431-441).  pixel_hourly_microclimate
    is the dict simulate_hourly_microclimate() returns ([H, C] arrays and the pixels' altitude)."""
    if water_shade_proportion != 1 and not model_water_temperature:           # :329-331
        warnings.warn("water temperature is not being modelled, so water_shade_proportion argument is ignored")
    if model_water_temperature:
        # the reference stop()s here (:335-337).  The GPU backend does model the water temperature,
        # but inside the fused chain, where the water body, its temperature and the rates that depend on it advance
        # together: monthly_climate_to_vectors() / tc_to_vectors(model_water_temperature = TRUE)
        raise NotImplementedError("model_water_temperature = TRUE is available through monthly_climate_to_vectors() / "
                                  "tc_to_vectors() (the fused chain), not through this stage-level verb")
    m = pixel_hourly_microclimate
    area = simulate_ephemeral_habitat_vectorised(m["rainfall"], m["air_temperature"], m["humidity"], m["windspeed"],
                                                 m["altitude"], math_mode=math_mode, device=device)["water_surface_area"]
    out = dict(m)
    out["water_surface_area"] = area
    return out


def simulate_hourly_vectors(pixel_hourly_lifehistory, *, burnin_years=1, math_mode="fast", device=0):
    """simulate_hourly_vectors(), R/pipeline_functions.R:691-906: the population model on the hourly life-history
    parameters of every pixel; adds adult and aquatic ([H, C])."""
    lhp = pixel_hourly_lifehistory
    pop = simulate_population_vectorised(lhp["larval_habitat_area"], lhp["mdr"], lhp["efd"], lhp["das_zerodensity"],
                                         lhp["ds"], lhp["das_densmod_function"], burnin_years=burnin_years,
                                         math_mode=math_mode, device=device)
    out = dict(lhp)
    out.update(pop)
    return out


def summarise_vectors(pixel_hourly_vectors, aggregate_by="month", included_dates=None, *, device=0):
    """summarise_vectors() -> aggregate_vectors(), R/pipeline_functions.R:995-1091: means of the six vector
    variables by date or by month over included_dates, on the device (mms_summarise_vectors).
    pixel_hourly_vectors: dict of [H, C] arrays; returns start[P], end[P] and one [P, C] array per variable."""
    aggregate_by = _match_arg(aggregate_by, ("month", "day"), "aggregate_by")
    dates_predict = np.asarray(pixel_hourly_vectors["date"], "datetime64[D]")[::24]      # one entry per simulated day
    period_of_day, p_start, p_end = dates.period_index(dates_predict, aggregate_by, included_dates)
    hourly = np.stack([np.asarray(pixel_hourly_vectors[v], dtype=np.float64) for v in VECTOR_VARIABLES], axis=0)
    means = backend.summarise_vectors(hourly, period_of_day, len(p_start), device=device)
    out = {"start": p_start, "end": p_end}
    for k, v in enumerate(VECTOR_VARIABLES):
        out[v] = means[k]
    return out


def _terraclimate_arrays(pixel_terraclimate_data):
    """(start, end, terraclimate[6][M][C], clear_sky[12][C], altitude[C]) from the dict form of
    extract_terraclimate_tile()'s per-pixel tibbles placed side by side."""
    d = pixel_terraclimate_data
    start = np.asarray(d["start"], dtype="datetime64[D]")
    end = np.asarray(d["end"], dtype="datetime64[D]")
    tc = np.stack([np.asarray(d[name], dtype=np.float64) for name in backend.TERRACLIMATE_VARS], axis=0)
    clear_sky = np.asarray(d["clear_sky_SOLR"], dtype=np.float64)
    altitude = np.asarray(d["altitude"], dtype=np.float64)
    if tc.shape[1] != len(start) or clear_sky.shape != (12, tc.shape[2]) or altitude.shape != (tc.shape[2],):
        raise ValueError("terraclimate variables must be [M, C], clear_sky_SOLR [12, C] and altitude [C]")
    return start, end, tc, clear_sky, altitude


def terraclimate_to_monthly_climate(pixel_terraclimate_data, *, pad_last=False, math_mode="fast", device=0):
    """terraclimate_to_monthly_climate() (R/pipeline_functions.R:179-252) for a whole tile.

    pixel_terraclimate_data: dict with start[M], end[M], the six TerraClimate variables tmax, tmin, ppt,
    ws, vpd, srad as [M, C] arrays, clear_sky_SOLR [12, C] (the pixel's clear-sky radiation by calendar
    month, R/pipeline_functions.R:194-222), altitude [C] (the reference reads it from its elevation raster,
    :228-233 -- raster I/O is outside this library), longitude / latitude (carried through).
    pad_last: apply pad_last_month() first (R/terraclimate.R:519-569; the reference does so when the series
    ends at the last published month, R/pipeline_functions.R:148-166).
    Returns the pixel_monthly_climate dict that monthly_climate_to_vectors() takes."""
    start, end, tc, clear_sky, altitude = _terraclimate_arrays(pixel_terraclimate_data)
    if pad_last:
        start, end, tc = dates.pad_last_month(start, end, tc)
    month_id, n_days = dates.terraclimate_month_info(start, end)
    monthly = backend.terraclimate_to_monthly(tc, clear_sky, month_id, n_days, math_mode=math_mode, device=device)
    out = {"start": start, "end": end, "altitude": altitude,
           "longitude": pixel_terraclimate_data.get("longitude"), "latitude": pixel_terraclimate_data.get("latitude")}
    for v, name in enumerate(FORCING_VARS):
        out[name] = monthly[v]
    return out


def _run_vectors(run, n_cells, start, end, carried, lifehistory_functions, species, included_dates,
                 aggregate_by, dates_predict, habitats, model_water_temperature, water_shade_proportion,
                 burnin_years_pop, extra_variables, math_mode, devices, water_temperature_options=None,
                 cell_idx=None, n_raster_cells=None, on_extrapolation="error"):
    """interpolate_daily_climate() ... summarise_vectors() in one pipelined call (mms_chain_run[_terraclimate]);
    `run(plan, out)` hands the monthly knots (either form) and the output buffers to the plan."""
    species = _match_arg(species, SPECIES, "species")
    aggregate_by = _match_arg(aggregate_by, ("month", "day"), "aggregate_by")
    # simulate_hourly_conditions() argument handling, R/pipeline_functions.R:324-338
    if water_shade_proportion != 1 and not model_water_temperature:
        warnings.warn("water temperature is not being modelled, so water_shade_proportion argument is ignored")
    lh: Lifehistory = lifehistory_functions[{"An. gambiae": "An_gambiae", "An. stephensi": "An_stephensi"}[species]]
    on_extrapolation = _match_arg(on_extrapolation, ("error", "clip"), "on_extrapolation")
    if dates_predict is None:
        dates_predict = dates.default_dates_predict(start, end)
    if on_extrapolation == "clip":
        # a half-day overhang of the default dates over the knots (30-day first month, non-leap February last month)
        # would send the reference into its mgcv::gam branch, which the library refuses (MMS_ERR_EXTRAPOLATION)
        dates_predict, dropped = dates.clip_dates_predict(start, end, dates_predict)
        if dropped:
            warnings.warn(f"{dropped} prediction day(s) outside the monthly mid-dates dropped "
                          "(the reference would extrapolate them with an mgcv::gam)")
    mid_day, day_predict = dates.datenums(start, end, dates_predict)
    period_of_day, p_start, p_end = dates.period_index(dates_predict, aggregate_by, included_dates)
    if habitats is None:      # R/pipeline_functions.R:431-441
        habitats = [dict(shape=0, max_depth=1.0, radius=1.0, inflow_multiplier=1.0, leak_per_hour=0.0,
                         initial_volume=0.0)]
    names = list(VECTOR_VARIABLES) + [v for v in extra_variables if v not in VECTOR_VARIABLES]
    wt = None
    if model_water_temperature:
        # EXTENSION (the reference stop()s, R/pipeline_functions.R:334-338): the model sketched in
        # query_terraclimate.R:370-497, DESIGN.md section 5.  Needs the pixels' latitude and clear-sky radiation.
        if carried.get("latitude") is None or carried.get("clear_sky_SOLR") is None:
            raise ValueError("model_water_temperature = TRUE needs the pixels' latitude and clear_sky_SOLR [12, C]")
        dp = np.asarray(dates_predict, dtype="datetime64[D]")
        wt = dict(water_temperature_options or {})
        wt.update(shade_proportion=float(water_shade_proportion), day_of_year=dates.day_of_year(dp),
                  month_of_day=dates.month_of_year(dp), latitude=np.asarray(carried["latitude"], dtype=np.float64),
                  clear_sky=np.asarray(carried["clear_sky_SOLR"], dtype=np.float64))
        if "water_temperature" not in names:
            names.append("water_temperature")
    with backend.ChainPlan(
            n_cells, mid_day=mid_day, day_predict=day_predict, habitats=habitats,
            densmod=backend.make_densmod(lh.das_densmod.type, lh.das_densmod.dd_effect, lh.das_densmod.surface_area_cm2),
            tables=lh.tables(), period_of_day=period_of_day, n_periods=len(p_start), burnin_years_water=0,
            burnin_years_pop=burnin_years_pop, mean_vars=names, want_wet_hours=True, math_mode=math_mode,
            devices=devices, water_temperature=wt, raster_cell_idx=cell_idx, n_raster_cells=n_raster_cells) as plan:
        out = run(plan, plan.allocate_outputs())
    res = {"longitude": carried.get("longitude"), "latitude": carried.get("latitude"),
           "species": species, "start": p_start, "end": p_end, "wet_hours": out["wet_hours"]}
    for k, v in enumerate(out["mean_state_vars"]):
        res[v] = out["mean_state"][:, k] if len(habitats) > 1 else out["mean_state"][0, k]
    for k, v in enumerate(out["mean_forcing_vars"]):
        res[v] = out["mean_forcing"][k]
    return res


def monthly_climate_to_vectors(pixel_monthly_climate, lifehistory_functions, species=SPECIES,
                               included_dates=None, aggregate_by="month", dates_predict=None, *,
                               habitats=None, model_water_temperature=False, water_shade_proportion=1,
                               burnin_years_pop=1, extra_variables=(), math_mode="fast", devices=None,
                               water_temperature_options=None, cell_idx=None, n_raster_cells=None,
                               on_extrapolation="error", na_knots="error"):
    """tc_to_vectors() (R/pipeline_functions.R:1099-1130) from
    interpolate_daily_climate() to summarise_vectors(), for a whole tile in one
    fused kernel launch per device.

    pixel_monthly_climate: dict with start[M], end[M] (dates), altitude[C],
      longitude[C], latitude[C] (carried through) and the nine monthly
      variables as [M, C] arrays (the `monthly_climate` tibbles of
      terraclimate_to_monthly_climate(), side by side).
    lifehistory_functions: {"An_gambiae": Lifehistory, "An_stephensi": Lifehistory}
      (the reference's sysdata object of the same name).
    Returns a dict: longitude, latitude, species, start[P], end[P] and one
    [P, C] array per vector variable -- the `pixel_vectors` tibbles side by side.
    on_extrapolation: "error" (default) or "clip": what to do when prediction days lie outside the monthly
      mid-dates -- the reference's default dates overhang them by half a day when the series starts in a 30-day
      month or ends in a non-leap February (dates.clip_dates_predict) -- where the reference fits an mgcv::gam.
    cell_idx / n_raster_cells: the pixels' 0-based raster cell numbers (strictly increasing) and the raster's size;
      every row then comes back as a whole raster slice [P, n_raster_cells] with NaN where the raster has no pixel
      of the tile -- the assignment raster[cell_idx] <- values of create_vector_raster_slice()
      (R/pipeline_functions.R:1316-1322), done by the kernel's stores on the device.
    na_knots: "error" (default: the library refuses NA in the monthly forcing) or "interpolate": NA knots between two
      non-NA ones are replaced by what stats::approx(), which drops them, interpolates there (fill_na_knots()).
    """
    start, end, monthly, altitude = _monthly_array(pixel_monthly_climate)
    if na_knots == "interpolate":
        monthly, _ = fill_na_knots(monthly, start, end)
    elif na_knots != "error":
        raise ValueError("na_knots must be 'error' or 'interpolate'")
    return _run_vectors(lambda plan, out: plan.run(monthly, altitude, out), monthly.shape[2], start, end,
                        pixel_monthly_climate, lifehistory_functions, species, included_dates, aggregate_by,
                        dates_predict, habitats, model_water_temperature, water_shade_proportion,
                        burnin_years_pop, extra_variables, math_mode, devices, water_temperature_options,
                        cell_idx, n_raster_cells, on_extrapolation)


def tc_to_vectors(pixel_terraclimate_data, lifehistory_functions, species=SPECIES, included_dates=None, *,
                  aggregate_by="month", dates_predict=None, pad_last=False, habitats=None,
                  model_water_temperature=False, water_shade_proportion=1, burnin_years_pop=1,
                  extra_variables=(), math_mode="fast", devices=None, water_temperature_options=None,
                  cell_idx=None, n_raster_cells=None, on_extrapolation="error"):
    """tc_to_vectors(pixel_terraclimate_data, species, included_dates), R/pipeline_functions.R:1099-1130,
    whole: the six raw TerraClimate variables go to the device and process_terraclimate_pixel_vars()
    (R/terraclimate.R:581-643) runs there, in front of the fused chain (mms_chain_upload_terraclimate).
    Input as for terraclimate_to_monthly_climate(); output as for monthly_climate_to_vectors()."""
    start, end, tc, clear_sky, altitude = _terraclimate_arrays(pixel_terraclimate_data)
    if pad_last:
        start, end, tc = dates.pad_last_month(start, end, tc)
    month_id, n_days = dates.terraclimate_month_info(start, end)
    return _run_vectors(lambda plan, out: plan.run_terraclimate(tc, clear_sky, month_id, n_days, altitude, out),
                        tc.shape[2], start, end, pixel_terraclimate_data, lifehistory_functions, species,
                        included_dates, aggregate_by, dates_predict, habitats, model_water_temperature,
                        water_shade_proportion, burnin_years_pop, extra_variables, math_mode, devices,
                        water_temperature_options, cell_idx, n_raster_cells, on_extrapolation)


def raster_slice(vectors, variable, year, month):
    """The values create_vector_raster_slice() pulls out of a tile for one (variable, year, month) slice
    (R/pipeline_functions.R:1262-1303): the row of the period whose start falls in that month, over
    the tile's pixels in their original order -- one contiguous [C] row of the library's output."""
    starts = np.asarray(vectors["start"], dtype="datetime64[D]")
    want = np.datetime64(f"{int(year):04d}-{int(month):02d}", "M")
    rows = np.flatnonzero(starts.astype("datetime64[M]") == want)          # matches_year & matches_month
    return vectors[variable][..., rows, :]
