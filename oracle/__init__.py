"""ctypes front end of the CPU oracle (oracle/mms_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product
package mosmicrosim_b200.  PARITY UNPINNED (see mms_oracle.h): no R, no
reference golden vectors; the oracle is a literal restatement cross-checked
against oracle/restate_numpy.py and the known-answer constants.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmms_oracle.so")

NFORCING = 9
NVARS = 13
VAR_NAMES = ("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds",
             "larval_habitat_area", "water_volume", "air_temperature", "humidity",
             "windspeed", "rainfall", "water_temperature")

_dp = ct.POINTER(ct.c_double)
_ip = ct.POINTER(ct.c_int)
_bp = ct.POINTER(ct.c_ubyte)


class Habitat(ct.Structure):
    _fields_ = [("shape", ct.c_int), ("max_depth", ct.c_double), ("radius", ct.c_double),
                ("inflow_multiplier", ct.c_double), ("leak_per_hour", ct.c_double),
                ("initial_volume", ct.c_double)]


class Densmod(ct.Structure):
    _fields_ = [("type", ct.c_int), ("dd_effect", ct.c_double), ("surface_area_cm2", ct.c_double)]


class Spline(ct.Structure):
    _fields_ = [("n", ct.c_int), ("x", _dp), ("y", _dp), ("b", _dp), ("c", _dp), ("d", _dp)]


class Surface(ct.Structure):
    _fields_ = [("nx", ct.c_int), ("ny", ct.c_int), ("x", _dp), ("y", _dp), ("z", _dp)]


class Lifehistory(ct.Structure):
    _fields_ = [("mdr", Spline), ("efd", Spline), ("das", Spline), ("ds", Surface)]


class WaterTemperature(ct.Structure):
    """EXTENSION (mms_oracle.h mmso_water_temperature): the model sketched in query_terraclimate.R:370-497."""
    _fields_ = [("enabled", ct.c_int), ("shade_proportion", ct.c_double), ("albedo", ct.c_double),
                ("emissivity", ct.c_double), ("convection_w_m2_k", ct.c_double), ("day_of_year", _ip), ("month_of_day", _ip),
                ("latitude", _dp), ("clear_sky", _dp)]


class Chain(ct.Structure):
    _fields_ = [("M", ct.c_int), ("D", ct.c_int), ("C", ct.c_int), ("S", ct.c_int), ("P", ct.c_int),
                ("monthly", _dp), ("mid_day", _dp), ("day_predict", _dp), ("altitude", _dp),
                ("habitats", ct.POINTER(Habitat)),
                ("burnin_years_water", ct.c_int), ("burnin_years_pop", ct.c_int),
                ("initial_adult", ct.c_double), ("initial_aquatic", ct.c_double),
                ("densmod", Densmod), ("lh", Lifehistory),
                ("period_of_day", _ip), ("wt", WaterTemperature),
                ("out_mean", _dp), ("out_wet_hours", _ip), ("out_hourly", _dp), ("out_wet", _bp),
                ("n_threads", ct.c_int)]


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "mms_oracle.c")
    hdr = os.path.join(_HERE, "mms_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.exists(p) and os.path.getmtime(_LIB_PATH) < os.path.getmtime(p)
                    for p in (src, hdr)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmms_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ct.CDLL(_LIB_PATH)
        L.mmso_approx1.restype = ct.c_double
        L.mmso_approx1.argtypes = [ct.c_double, _dp, _dp, ct.c_long]
        L.mmso_R_pow.restype = ct.c_double
        L.mmso_R_pow.argtypes = [ct.c_double, ct.c_double]
        for name in ("mmso_cone_depth_to_volume", "mmso_cone_volume_to_surface",
                     "mmso_saturation_vapour_pressure_kpa", "mmso_air_pressure_kpa"):
            getattr(L, name).restype = ct.c_double
            getattr(L, name).argtypes = [ct.c_double]
        L.mmso_evaporation_rate_kg_h.restype = ct.c_double
        L.mmso_evaporation_rate_kg_h.argtypes = [ct.c_double] * 5
        for name in ("mmso_habitat_max_volume", "mmso_habitat_catchment"):
            getattr(L, name).restype = ct.c_double
            getattr(L, name).argtypes = [ct.POINTER(Habitat)]
        L.mmso_habitat_surface.restype = ct.c_double
        L.mmso_habitat_surface.argtypes = [ct.POINTER(Habitat), ct.c_double]
        L.mmso_densmod_apply.restype = ct.c_double
        L.mmso_densmod_apply.argtypes = [ct.POINTER(Densmod), ct.c_double, ct.c_double]
        L.mmso_R_mean.restype = ct.c_double
        L.mmso_R_mean.argtypes = [_dp, ct.c_long]
        L.mmso_run_chain.restype = ct.c_int
        L.mmso_run_chain.argtypes = [ct.POINTER(Chain)]
        L.mmso_max_threads.restype = ct.c_int
        L.mmso_solar_declination.restype = None
        L.mmso_solar_declination.argtypes = [ct.c_int, _dp, _dp]
        L.mmso_cos_hour_angle.restype = ct.c_double
        L.mmso_cos_hour_angle.argtypes = [ct.c_int]
        L.mmso_clear_sky_hourly.restype = None
        L.mmso_clear_sky_hourly.argtypes = [ct.c_double, ct.c_int, ct.c_double, _dp]
        for name, n in (("mmso_vapprs", 1), ("mmso_vapour_pressure", 2), ("mmso_relative_humidity", 2),
                        ("mmso_adjust_wind_speed", 4), ("mmso_cloud_cover", 3)):
            getattr(L, name).restype = ct.c_double
            getattr(L, name).argtypes = [ct.c_double] * n
        L.mmso_process_terraclimate_pixel_vars.restype = None
        L.mmso_process_terraclimate_pixel_vars.argtypes = [_dp, _dp, ct.POINTER(ct.c_int), _dp, ct.c_long,
                                                           ct.c_int, _dp]
        _lib = L
    return _lib


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


# ---------------------------------------------------------------------------
# stage-level wrappers (numpy in / numpy out)
# ---------------------------------------------------------------------------

def approx1(v, x, y):
    x, y = _d(x), _d(y)
    return lib().mmso_approx1(float(v), _p(x), _p(y), len(x))


def spline_seasonal(values, datenum, datenum_predict):
    values, datenum, dp = _d(values), _d(datenum), _d(datenum_predict)
    out = np.empty(len(dp))
    rc = lib().mmso_spline_seasonal(_p(values), _p(datenum), len(values), _p(dp), len(dp), _p(out))
    if rc == -3:
        raise ValueError("need at least two non-NA values to interpolate (stats::approx)")
    if rc:
        raise ValueError("extrapolation needed (GAM branch, R/spline_interpolation.R:30-112, not restated)")
    return out


def daily_from_monthly(monthly, mid_day, day_predict):
    """monthly [9][M] (one cell) -> daily [9][D]"""
    monthly, mid_day, dp = _d(monthly), _d(mid_day), _d(day_predict)
    M, D = monthly.shape[1], len(dp)
    daily = np.empty((NFORCING, D))
    rc = lib().mmso_daily_from_monthly(_p(monthly), _p(mid_day), M, _p(dp), D, _p(daily))
    if rc == -3:
        raise ValueError("need at least two non-NA values to interpolate (stats::approx)")
    if rc:
        raise ValueError("extrapolation needed (GAM branch not restated)")
    return daily


def interpolate_daily_air_temp(daily_max, daily_min):
    mx, mn = _d(daily_max), _d(daily_min)
    out = np.empty(24 * len(mx))
    lib().mmso_interpolate_daily_air_temp(_p(mx), _p(mn), len(mx), _p(out))
    return out


def vsine_weights(times, time_sunrise=6.0, time_inflection=13.0):
    t = _d(times)
    out = np.empty(len(t))
    lib().mmso_vsine_weights(_p(t), ct.c_long(len(t)), ct.c_double(time_sunrise),
                             ct.c_double(time_inflection), _p(out))
    return out


def vsine_interpolate_daily(daily_max, daily_min, sunrise_value="max"):
    mx, mn = _d(daily_max), _d(daily_min)
    out = np.empty(24 * len(mx))
    lib().mmso_vsine_interpolate_daily(_p(mx), _p(mn), len(mx), int(sunrise_value == "max"), _p(out))
    return out


def hourly_from_daily_ambient(daily):
    daily = _d(daily)
    D = daily.shape[1]
    outs = [np.empty(24 * D) for _ in range(6)]
    lib().mmso_hourly_from_daily_ambient(_p(daily), D, *[_p(o) for o in outs])
    names = ("air_temperature", "humidity", "cloudcover", "windspeed", "water_temperature", "rainfall")
    return dict(zip(names, outs))


def habitat(shape=0, max_depth=1.0, radius=1.0, inflow_multiplier=1.0, leak_per_hour=0.0,
            initial_volume=0.0) -> Habitat:
    return Habitat(int(shape), float(max_depth), float(radius), float(inflow_multiplier),
                   float(leak_per_hour), float(initial_volume))


def simulate_ephemeral_habitat_vectorised(rainfall, air_temperature, humidity, windspeed,
                                          altitude, hb: Habitat, burnin_years=0):
    """Matrices are R-style time-by-pixel; pass them as numpy [H, C] arrays (any order)."""
    mats = [np.asfortranarray(m, dtype=np.float64) for m in (rainfall, air_temperature, humidity, windspeed)]
    H, C = mats[0].shape
    alt = _d(altitude)
    area = np.empty((H, C), order="F")
    vol = np.empty((H, C), order="F")
    wet = np.empty((H, C), order="F", dtype=np.uint8)
    rc = lib().mmso_simulate_ephemeral_habitat_vectorised(
        *[m.ctypes.data_as(_dp) for m in mats], _p(alt), ct.c_long(H), ct.c_long(C),
        ct.byref(hb), int(burnin_years), area.ctypes.data_as(_dp), vol.ctypes.data_as(_dp),
        wet.ctypes.data_as(_bp))
    if rc:
        raise IndexError("subscript out of bounds (burn-in needs >= 8760 rows)")
    return {"water_surface_area": area, "water_volume": vol, "wet": wet}


def fmm_spline(x, y):
    x, y = _d(x), _d(y)
    n = len(x)
    b, c, d = np.empty(n), np.empty(n), np.empty(n)
    lib().mmso_fmm_spline(n, _p(x), _p(y), _p(b), _p(c), _p(d))
    return b, c, d


def make_spline(tab) -> Spline:
    """tab: dict with x, y, b, c, d (kept alive by the caller)."""
    return Spline(len(tab["x"]), *[_p(tab[k]) for k in ("x", "y", "b", "c", "d")])


def make_surface(tab) -> Surface:
    return Surface(len(tab["x"]), len(tab["y"]), _p(tab["x"]), _p(tab["y"]), _p(tab["z"]))


def spline_eval(tab, u, positive=False):
    u = _d(u)
    out = np.empty(len(u))
    s = make_spline(tab)
    fn = lib().mmso_positive_spline if positive else lib().mmso_spline_eval
    fn(ct.byref(s), _p(u), ct.c_long(len(u)), _p(out))
    return out


def interp_surface(tab, tx, ty):
    tx, ty = _d(tx), _d(ty)
    out = np.empty(len(tx))
    g = make_surface(tab)
    lib().mmso_interp_surface(ct.byref(g), _p(tx), _p(ty), ct.c_long(len(tx)), _p(out))
    return out


def densmod(type_, dd_effect, surface_area_cm2) -> Densmod:
    t = {"logit": 0, "cox_ph": 1}.get(type_, type_)
    return Densmod(int(t), float(dd_effect), float(surface_area_cm2))


def simulate_population_vectorised(larval_habitat_area, mdr, efd, das_zerodensity, ds, dm: Densmod,
                                   initial_adult=100.0, initial_aquatic=100.0, burnin_years=1):
    mats = [np.asfortranarray(m, dtype=np.float64)
            for m in (larval_habitat_area, mdr, efd, das_zerodensity, ds)]
    H, C = mats[0].shape
    aquatic = np.empty((H, C), order="F")
    adult = np.empty((H, C), order="F")
    rc = lib().mmso_simulate_population_vectorised(
        *[m.ctypes.data_as(_dp) for m in mats], ct.c_long(H), ct.c_long(C), ct.byref(dm),
        ct.c_double(initial_adult), ct.c_double(initial_aquatic), int(burnin_years),
        aquatic.ctypes.data_as(_dp), adult.ctypes.data_as(_dp))
    if rc:
        raise IndexError("subscript out of bounds (burn-in needs >= 8760 rows)")
    return {"aquatic": aquatic, "adult": adult}


def simulate_ephemeral_habitat(rainfall, air_temperature, humidity, windspeed, altitude, hb: Habitat,
                               burnin_years=0):
    """One-pixel twin (R/sim_water.R:136-205): vectors in, surface-area vector out."""
    vecs = [_d(v) for v in (rainfall, air_temperature, humidity, windspeed)]
    out = np.empty(len(vecs[0]))
    lib().mmso_simulate_ephemeral_habitat(*[_p(v) for v in vecs], ct.c_double(altitude), ct.c_long(len(out)),
                                          ct.byref(hb), int(burnin_years), _p(out))
    return out


def simulate_population(larval_habitat_area, mdr, efd, das_zerodensity, ds, dm: Densmod,
                        initial_adult=100.0, initial_aquatic=100.0, burnin_years=1):
    """One-pixel twin (R/sim_pop_dynamics.R:138-196): vectors in, {"aquatic", "adult"} vectors out."""
    vecs = [_d(v) for v in (larval_habitat_area, mdr, efd, das_zerodensity, ds)]
    aquatic, adult = np.empty(len(vecs[0])), np.empty(len(vecs[0]))
    lib().mmso_simulate_population(*[_p(v) for v in vecs], ct.c_long(len(aquatic)), ct.byref(dm),
                                   ct.c_double(initial_adult), ct.c_double(initial_aquatic), int(burnin_years),
                                   _p(aquatic), _p(adult))
    return {"aquatic": aquatic, "adult": adult}


TERRACLIMATE_VARS = ("tmax", "tmin", "ppt", "ws", "vpd", "srad")      # R/pipeline_functions.R:134-136


def process_terraclimate_pixel_vars(terraclimate, clear_sky, month_id, n_days):
    """process_terraclimate_pixel_vars() (R/terraclimate.R:581-643) for C pixels: terraclimate [6][M][C]
    (TERRACLIMATE_VARS order), clear_sky [12][C], month_id [M] (1..12), n_days [M] -> monthly [9][M][C]."""
    tc, cs, nd = _d(terraclimate), _d(clear_sky), _d(n_days)
    mid = np.ascontiguousarray(month_id, dtype=np.int32)
    _, M, C = tc.shape
    assert cs.shape == (12, C) and len(mid) == M and len(nd) == M
    out = np.empty((9, M, C))
    lib().mmso_process_terraclimate_pixel_vars(_p(tc), _p(cs), mid.ctypes.data_as(ct.POINTER(ct.c_int)), _p(nd),
                                               ct.c_long(C), M, _p(out))
    return out


def solar_declination(doy):
    """(sin delta, cos delta) of a day of the year (EXTENSION: water-temperature model)"""
    sd, cd = ct.c_double(), ct.c_double()
    lib().mmso_solar_declination(int(doy), ct.byref(sd), ct.byref(cd))
    return sd.value, cd.value


def cos_hour_angle(hour):
    return lib().mmso_cos_hour_angle(int(hour))


def clear_sky_hourly(latitude, doy, clear_sky_mean):
    """Clear-sky irradiance of the 24 hours of a day whose daily mean is clear_sky_mean."""
    out = np.empty(24)
    lib().mmso_clear_sky_hourly(float(latitude), int(doy), float(clear_sky_mean), _p(out))
    return out


def R_mean(x):
    x = _d(x)
    return lib().mmso_R_mean(_p(x), len(x))


def max_threads() -> int:
    return lib().mmso_max_threads()


def run_chain(monthly, mid_day, day_predict, altitude, habitats, densmod_, tables,
              period_of_day=None, n_periods=0, burnin_years_water=0, burnin_years_pop=1,
              initial_adult=100.0, initial_aquatic=100.0, want_mean=True, want_hourly=False,
              n_threads=0, water_temperature=None):
    """Fused chain for C cells.  monthly is [9][M][C]; tables is the dict made
    by mosmicrosim_b200.synthetic.synthetic_lifehistory (or any dict with the
    same keys).  Returns a dict of numpy arrays laid out [S][var][period|t][C]."""
    monthly = _d(monthly)
    _, M, C = monthly.shape
    mid_day, dp, alt = _d(mid_day), _d(day_predict), _d(altitude)
    D, H = len(dp), 24 * len(dp)
    S = len(habitats)
    hab_arr = (Habitat * S)(*habitats)
    if period_of_day is None:
        period_of_day = np.full(D, -1, dtype=np.int32)
    pod = np.ascontiguousarray(period_of_day, dtype=np.int32)
    P = int(n_periods)
    ch = Chain()
    ch.M, ch.D, ch.C, ch.S, ch.P = M, D, C, S, P
    ch.monthly, ch.mid_day, ch.day_predict, ch.altitude = _p(monthly), _p(mid_day), _p(dp), _p(alt)
    ch.habitats = hab_arr
    ch.burnin_years_water, ch.burnin_years_pop = int(burnin_years_water), int(burnin_years_pop)
    ch.initial_adult, ch.initial_aquatic = float(initial_adult), float(initial_aquatic)
    ch.densmod = densmod_
    ch.lh = Lifehistory(make_spline(tables["mdr"]), make_spline(tables["efd"]),
                        make_spline(tables["das"]), make_surface(tables["ds"]))
    ch.period_of_day = pod.ctypes.data_as(_ip)
    if water_temperature is not None:
        # dict(shade_proportion, albedo, emissivity, convection_w_m2_k, day_of_year[D], month_of_day[D], latitude[C], clear_sky[12][C])
        w = water_temperature
        doy = np.ascontiguousarray(w["day_of_year"], dtype=np.int32)
        mon = np.ascontiguousarray(w["month_of_day"], dtype=np.int32)
        lat, cs = _d(w["latitude"]), _d(w["clear_sky"])
        assert doy.shape == (D,) and mon.shape == (D,) and lat.shape == (C,) and cs.shape == (12, C)
        ch.wt = WaterTemperature(1, float(w.get("shade_proportion", 1.0)), float(w.get("albedo", 0.0)),
                                 float(w.get("emissivity", 0.97)), float(w.get("convection_w_m2_k", 0.0)),
                                 doy.ctypes.data_as(_ip), mon.ctypes.data_as(_ip),
                                 _p(lat), _p(cs))
    out = {}
    if want_mean and P > 0:
        out["mean"] = np.empty((S, NVARS, P, C))
        out["wet_hours"] = np.empty((S, P, C), dtype=np.int32)
        ch.out_mean = _p(out["mean"])
        ch.out_wet_hours = out["wet_hours"].ctypes.data_as(_ip)
    if want_hourly:
        out["hourly"] = np.empty((S, NVARS, H, C))
        out["wet"] = np.empty((S, H, C), dtype=np.uint8)
        ch.out_hourly = _p(out["hourly"])
        ch.out_wet = out["wet"].ctypes.data_as(_bp)
    ch.n_threads = int(n_threads)
    rc = lib().mmso_run_chain(ct.byref(ch))
    if rc == -1:
        raise ValueError("extrapolation needed or burn-in longer than the series")
    if rc == -3:
        raise ValueError("need at least two non-NA values to interpolate (stats::approx)")
    if rc:
        raise RuntimeError(f"oracle chain failed: {rc}")
    return out
