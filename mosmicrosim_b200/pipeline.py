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


def interpolate_daily_climate(pixel_monthly_climate, dates_predict=None, *, math_mode="fast", device=0):
    """interpolate_daily_climate() (R/pipeline_functions.R:258-275): monthly -> daily for every
    pixel of the tile.  Returns the tile's `daily_climate` tibbles side by side: a dict with
    `date` [D], the pixel columns carried through, and one [D, C] array per daily variable."""
    start, end, monthly, altitude = _monthly_array(pixel_monthly_climate)
    if dates_predict is None:
        dates_predict = dates.default_dates_predict(start, end)
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


def monthly_climate_to_vectors(pixel_monthly_climate, lifehistory_functions, species=SPECIES,
                               included_dates=None, aggregate_by="month", dates_predict=None, *,
                               habitats=None, model_water_temperature=False, water_shade_proportion=1,
                               burnin_years_pop=1, extra_variables=(), math_mode="fast", devices=None):
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
    """
    species = _match_arg(species, SPECIES, "species")
    aggregate_by = _match_arg(aggregate_by, ("month", "day"), "aggregate_by")
    # simulate_hourly_conditions() argument handling, R/pipeline_functions.R:324-338
    if water_shade_proportion != 1 and not model_water_temperature:
        warnings.warn("water temperature is not being modelled, so water_shade_proportion argument is ignored")
    if model_water_temperature:
        raise NotImplementedError("model of water temperature not yet implemented, "
                                  "model_water_temperature must be set to FALSE")
    lh: Lifehistory = lifehistory_functions[{"An. gambiae": "An_gambiae", "An. stephensi": "An_stephensi"}[species]]
    start, end, monthly, altitude = _monthly_array(pixel_monthly_climate)
    if dates_predict is None:
        dates_predict = dates.default_dates_predict(start, end)
    mid_day, day_predict = dates.datenums(start, end, dates_predict)
    period_of_day, p_start, p_end = dates.period_index(dates_predict, aggregate_by, included_dates)
    if habitats is None:      # R/pipeline_functions.R:431-441
        habitats = [dict(shape=0, max_depth=1.0, radius=1.0, inflow_multiplier=1.0, leak_per_hour=0.0,
                         initial_volume=0.0)]
    names = list(VECTOR_VARIABLES) + [v for v in extra_variables if v not in VECTOR_VARIABLES]
    out = backend.run_chain(
        monthly, altitude, mid_day=mid_day, day_predict=day_predict, habitats=habitats,
        densmod=backend.make_densmod(lh.das_densmod.type, lh.das_densmod.dd_effect, lh.das_densmod.surface_area_cm2),
        tables=lh.tables(), period_of_day=period_of_day, n_periods=len(p_start), burnin_years_water=0,
        burnin_years_pop=burnin_years_pop, mean_vars=names, want_wet_hours=True, math_mode=math_mode,
        devices=devices)
    res = {"longitude": pixel_monthly_climate.get("longitude"), "latitude": pixel_monthly_climate.get("latitude"),
           "species": species, "start": p_start, "end": p_end, "wet_hours": out["wet_hours"]}
    for k, v in enumerate(out["mean_state_vars"]):
        res[v] = out["mean_state"][:, k] if len(habitats) > 1 else out["mean_state"][0, k]
    for k, v in enumerate(out["mean_forcing_vars"]):
        res[v] = out["mean_forcing"][k]
    return res
