"""ctypes binding of libmmsgpu (include/mmsgpu.h) -- the same C ABI the R
`.Call` shim binds (r-binding/src/mmsgpu_shim.c).  numpy arrays in, numpy
arrays out; nothing here computes: every numeric result comes from the CUDA
kernels, and every entry point raises if the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMS_LIB") or os.path.join(_HERE, "libmmsgpu.so")   # MMS_LIB: another build of the same library

NFORCING = 9
NTERRACLIMATE = 6
TERRACLIMATE_VARS = ("tmax", "tmin", "ppt", "ws", "vpd", "srad")   # R/pipeline_functions.R:134-136
NVARS = 14
VAR_NAMES = ("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds",
             "larval_habitat_area", "water_volume",
             "air_temperature", "humidity", "windspeed", "rainfall", "cloudcover", "water_temperature")
VAR_INDEX = {n: i for i, n in enumerate(VAR_NAMES)}
STATE_VAR_MASK = (1 << 0) | (1 << 1) | (1 << 6) | (1 << 7) | (1 << 13)
PIPELINE_VAR_MASK = 0x3F
MAX_SETTINGS = 64

STATUS = {0: "MMS_OK", 1: "MMS_ERR_INVALID_ARG", 2: "MMS_ERR_EXTRAPOLATION", 3: "MMS_ERR_BURNIN",
          4: "MMS_ERR_NONFINITE", 5: "MMS_ERR_CUDA", 6: "MMS_ERR_NO_DEVICE", 7: "MMS_ERR_ALLOC",
          8: "MMS_ERR_UNSUPPORTED"}

_dp = ct.POINTER(ct.c_double)
_ip = ct.POINTER(ct.c_int32)
_bp = ct.POINTER(ct.c_uint8)


class MmsError(RuntimeError):
    """A non-zero mms_status; mirrors the reference's stop(..., call. = FALSE)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status
        self.code = STATUS.get(status, str(status))


class Habitat(ct.Structure):
    _fields_ = [("shape", ct.c_int32), ("reserved", ct.c_int32), ("max_depth", ct.c_double),
                ("radius", ct.c_double), ("inflow_multiplier", ct.c_double),
                ("leak_per_hour", ct.c_double), ("initial_volume", ct.c_double)]


class Densmod(ct.Structure):
    _fields_ = [("type", ct.c_int32), ("reserved", ct.c_int32), ("dd_effect", ct.c_double),
                ("surface_area_cm2", ct.c_double)]


class Spline(ct.Structure):
    _fields_ = [("n", ct.c_int32), ("reserved", ct.c_int32), ("x", _dp), ("y", _dp), ("b", _dp),
                ("c", _dp), ("d", _dp)]


class Surface(ct.Structure):
    _fields_ = [("nx", ct.c_int32), ("ny", ct.c_int32), ("x", _dp), ("y", _dp), ("z", _dp)]


class LifehistoryC(ct.Structure):
    _fields_ = [("mdr_temp", Spline), ("efd_temp", Spline), ("das_temp", Spline), ("ds_temp_humid", Surface)]


class WaterTemperature(ct.Structure):
    """mms_water_temperature: the water-temperature EXTENSION (include/mmsgpu.h)."""
    _fields_ = [("enabled", ct.c_int32), ("reserved", ct.c_int32), ("shade_proportion", ct.c_double),
                ("albedo", ct.c_double), ("emissivity", ct.c_double), ("convection_w_m2_k", ct.c_double),
                ("day_of_year", _ip), ("month_of_day", _ip)]


class ChainConfig(ct.Structure):
    _fields_ = [("n_cells", ct.c_int64), ("n_months", ct.c_int32), ("n_days", ct.c_int32),
                ("n_settings", ct.c_int32), ("n_periods", ct.c_int32),
                ("mid_day", _dp), ("day_predict", _dp), ("period_of_day", _ip),
                ("habitats", ct.POINTER(Habitat)),
                ("burnin_years_water", ct.c_int32), ("burnin_years_pop", ct.c_int32),
                ("initial_adult", ct.c_double), ("initial_aquatic", ct.c_double),
                ("densmod", Densmod), ("lifehistory", LifehistoryC),
                ("mean_var_mask", ct.c_uint32), ("hourly_var_mask", ct.c_uint32),
                ("want_wet_hours", ct.c_int32), ("want_wet_flags", ct.c_int32),
                ("math_mode", ct.c_int32), ("n_devices", ct.c_int32), ("devices", _ip),
                ("water_temperature", WaterTemperature),
                ("raster_cell_idx", ct.POINTER(ct.c_int64)), ("n_raster_cells", ct.c_int64)]


class ChainOutputs(ct.Structure):
    _fields_ = [("mean_state", _dp), ("mean_forcing", _dp), ("hourly_state", _dp),
                ("hourly_forcing", _dp), ("wet_hours", _ip), ("wet_flags", _bp)]


EXPORTS = ("mms_abi_version", "mms_device_count", "mms_last_error", "mms_run_chain",
           "mms_chain_create", "mms_chain_upload", "mms_chain_launch", "mms_chain_sync",
           "mms_chain_run", "mms_chain_run_terraclimate", "mms_chain_plan_info", "mms_chain_set_location",
           "mms_chain_download", "mms_chain_kernel_ms", "mms_chain_launch_count", "mms_chain_destroy",
           "mms_shard_range",
           "mms_water_vectorised", "mms_population_vectorised", "mms_water_series", "mms_population_series",
           "mms_stage_kernel_ms",
           "mms_chain_upload_terraclimate", "mms_run_chain_terraclimate", "mms_terraclimate_to_monthly",
           "mms_summarise_vectors",
           "mms_daily_from_monthly", "mms_hourly_from_daily", "mms_lifehistory_vectorised", "mms_host_alloc", "mms_host_free",
           "mms_measure_fp64_peak", "mms_measure_hbm_copy", "mms_flush_l2", "mms_device_info")

_lib = None


def lib() -> ct.CDLL:
    """Load libmmsgpu.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m mosmicrosim_b200.build` "
                               "(there is no CPU fallback)")
        L = ct.CDLL(LIB_PATH)
        L.mms_last_error.restype = ct.c_char_p
        vp = ct.c_void_p
        L.mms_run_chain.argtypes = [ct.POINTER(ChainConfig), _dp, _dp, ct.POINTER(ChainOutputs)]
        L.mms_chain_create.argtypes = [ct.POINTER(ChainConfig), ct.POINTER(vp)]
        L.mms_chain_upload.argtypes = [vp, _dp, _dp]
        L.mms_chain_upload_terraclimate.argtypes = [vp, _dp, _dp, _ip, _dp, _dp]
        L.mms_run_chain_terraclimate.argtypes = [ct.POINTER(ChainConfig), _dp, _dp, _ip, _dp, _dp,
                                                 ct.POINTER(ChainOutputs)]
        L.mms_summarise_vectors.argtypes = [_dp, ct.c_int64, ct.c_int32, ct.c_int32, _ip, ct.c_int32, ct.c_int32, _dp]
        L.mms_terraclimate_to_monthly.argtypes = [_dp, _dp, _ip, _dp, ct.c_int64, ct.c_int32, ct.c_int32,
                                                  ct.c_int32, _dp]
        L.mms_chain_launch.argtypes = [vp]
        L.mms_chain_run.argtypes = [vp, _dp, _dp, ct.POINTER(ChainOutputs)]
        L.mms_chain_run_terraclimate.argtypes = [vp, _dp, _dp, _ip, _dp, _dp, ct.POINTER(ChainOutputs)]
        L.mms_chain_plan_info.argtypes = [vp, _ip, _ip, _ip]
        L.mms_chain_set_location.argtypes = [vp, _dp, _dp]
        L.mms_chain_sync.argtypes = [vp]
        L.mms_chain_download.argtypes = [vp, ct.POINTER(ChainOutputs)]
        L.mms_chain_kernel_ms.argtypes = [vp, ct.POINTER(ct.c_float)]
        L.mms_chain_launch_count.argtypes = [vp, ct.POINTER(ct.c_int64)]
        L.mms_chain_destroy.argtypes = [vp]
        L.mms_chain_destroy.restype = None
        L.mms_shard_range.argtypes = [ct.c_int64, ct.c_int32, ct.c_int32, ct.POINTER(ct.c_int64),
                                      ct.POINTER(ct.c_int64)]
        L.mms_water_vectorised.argtypes = [_dp, _dp, _dp, _dp, _dp, ct.c_int64, ct.c_int64,
                                           ct.POINTER(Habitat), ct.c_int32, ct.c_int32, ct.c_int32,
                                           _dp, _dp, _bp]
        L.mms_population_vectorised.argtypes = [_dp, _dp, _dp, _dp, _dp, ct.c_int64, ct.c_int64,
                                                ct.POINTER(Densmod), ct.c_double, ct.c_double,
                                                ct.c_int32, ct.c_int32, ct.c_int32, _dp, _dp]
        L.mms_water_series.argtypes = L.mms_water_vectorised.argtypes
        L.mms_population_series.argtypes = L.mms_population_vectorised.argtypes
        L.mms_stage_kernel_ms.argtypes = [ct.POINTER(ct.c_float)]
        L.mms_daily_from_monthly.argtypes = [_dp, _dp, _dp, ct.c_int64, ct.c_int32, ct.c_int32, ct.c_int32,
                                             ct.c_int32, _dp]
        L.mms_hourly_from_daily.argtypes = [_dp, ct.c_int64, ct.c_int32, ct.c_int32, ct.c_int32, _dp]
        L.mms_lifehistory_vectorised.argtypes = [_dp, _dp, _dp, ct.c_int64, ct.POINTER(LifehistoryC), ct.c_int32,
                                                 ct.c_int32, _dp, _dp, _dp, _dp]
        L.mms_host_alloc.argtypes = [ct.POINTER(vp), ct.c_uint64]
        L.mms_host_free.argtypes = [vp]
        L.mms_measure_fp64_peak.argtypes = [ct.c_int32, ct.POINTER(ct.c_double)]
        L.mms_measure_hbm_copy.argtypes = [ct.c_int32, ct.POINTER(ct.c_double)]
        L.mms_flush_l2.argtypes = [ct.c_int32]
        L.mms_device_info.argtypes = [ct.c_int32, _ip, _ip, ct.POINTER(ct.c_uint64)]
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise MmsError(status, lib().mms_last_error().decode("utf-8", "replace"))


def shard_range(n_cells: int, n_shards: int, index: int):
    """(cell0, count) of shard `index` of `n_shards` (mms_shard_range; pure host logic)."""
    c0, n = ct.c_int64(), ct.c_int64()
    check(lib().mms_shard_range(int(n_cells), int(n_shards), int(index), ct.byref(c0), ct.byref(n)))
    return int(c0.value), int(n.value)


def device_count() -> int:
    return int(lib().mms_device_count())


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def make_habitat(shape=0, max_depth=1.0, radius=1.0, inflow_multiplier=1.0, leak_per_hour=0.0,
                 initial_volume=0.0) -> Habitat:
    return Habitat(int(shape), 0, float(max_depth), float(radius), float(inflow_multiplier),
                   float(leak_per_hour), float(initial_volume))


def make_densmod(type_, dd_effect, surface_area_cm2) -> Densmod:
    t = {"logit": 0, "cox_ph": 1}.get(type_, type_)
    return Densmod(int(t), 0, float(dd_effect), float(surface_area_cm2))


def var_mask(names) -> int:
    m = 0
    for n in names or ():
        m |= 1 << VAR_INDEX[n]
    return m


def split_mask(mask: int):
    state = [v for v in range(NVARS) if mask >> v & 1 and STATE_VAR_MASK >> v & 1]
    forcing = [v for v in range(NVARS) if mask >> v & 1 and not STATE_VAR_MASK >> v & 1]
    return state, forcing


class PinnedArray:
    """A numpy view of page-locked host memory from mms_host_alloc()."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = ct.c_void_p()
        check(lib().mms_host_alloc(ct.byref(self._ptr), max(self.nbytes, 1)))
        buf = (ct.c_uint8 * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().mms_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ChainPlan:
    """Resident form of the fused chain (mms_chain_create ... mms_chain_destroy)."""

    def __init__(self, n_cells, mid_day, day_predict, habitats, densmod, tables, period_of_day=None,
                 n_periods=0, burnin_years_water=0, burnin_years_pop=1, initial_adult=100.0,
                 initial_aquatic=100.0, mean_vars=(), hourly_vars=(), want_wet_hours=False,
                 want_wet_flags=False, math_mode="fast", devices=None, water_temperature=None,
                 raster_cell_idx=None, n_raster_cells=None, _create=True):
        self._keep = []
        cfg = ChainConfig()
        self.C = int(n_cells)
        mid_day, day_predict = _d(mid_day), _d(day_predict)
        self.M, self.D = len(mid_day), len(day_predict)
        self.H = 24 * self.D
        self.S = len(habitats)
        self.P = int(n_periods)
        hab = (Habitat * self.S)(*[h if isinstance(h, Habitat) else make_habitat(**h) for h in habitats])
        self._keep += [mid_day, day_predict, hab]
        cfg.n_cells, cfg.n_months, cfg.n_days = self.C, self.M, self.D
        cfg.n_settings, cfg.n_periods = self.S, self.P
        cfg.mid_day, cfg.day_predict = _p(mid_day), _p(day_predict)
        if period_of_day is not None:
            pod = np.ascontiguousarray(period_of_day, dtype=np.int32)
            if len(pod) != self.D:
                raise ValueError("period_of_day must have one entry per simulated day")
            self._keep.append(pod)
            cfg.period_of_day = pod.ctypes.data_as(_ip)
        cfg.habitats = hab
        cfg.burnin_years_water, cfg.burnin_years_pop = int(burnin_years_water), int(burnin_years_pop)
        cfg.initial_adult, cfg.initial_aquatic = float(initial_adult), float(initial_aquatic)
        cfg.densmod = densmod if isinstance(densmod, Densmod) else make_densmod(**densmod)
        tabs = {k: {kk: _d(vv) for kk, vv in v.items()} for k, v in tables.items()}
        self._keep.append(tabs)

        def sp(t):
            return Spline(len(t["x"]), 0, _p(t["x"]), _p(t["y"]), _p(t["b"]), _p(t["c"]), _p(t["d"]))

        g = tabs["ds"]
        cfg.lifehistory = LifehistoryC(sp(tabs["mdr"]), sp(tabs["efd"]), sp(tabs["das"]),
                                       Surface(len(g["x"]), len(g["y"]), _p(g["x"]), _p(g["y"]), _p(g["z"])))
        self.mean_mask = mean_vars if isinstance(mean_vars, int) else var_mask(mean_vars)
        self.hourly_mask = hourly_vars if isinstance(hourly_vars, int) else var_mask(hourly_vars)
        cfg.mean_var_mask, cfg.hourly_var_mask = self.mean_mask, self.hourly_mask
        self.want_wet_hours, self.want_wet_flags = bool(want_wet_hours), bool(want_wet_flags)
        cfg.want_wet_hours, cfg.want_wet_flags = int(self.want_wet_hours), int(self.want_wet_flags)
        cfg.math_mode = {"fast": 0, "strict": 1}.get(math_mode, math_mode)
        if devices is not None:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            self._keep.append(dv)
            cfg.n_devices, cfg.devices = len(dv), dv.ctypes.data_as(_ip)
        self.R = self.C                  # columns of an output row
        if raster_cell_idx is not None:
            idx = np.ascontiguousarray(raster_cell_idx, dtype=np.int64)
            if idx.shape != (self.C,):
                raise ValueError("raster_cell_idx must have one entry per cell")
            self._keep.append(idx)
            self.R = int(n_raster_cells)
            cfg.raster_cell_idx, cfg.n_raster_cells = idx.ctypes.data_as(ct.POINTER(ct.c_int64)), self.R
        self._location = None
        if water_temperature is not None:
            # EXTENSION: dict(shade_proportion, albedo, emissivity, convection_w_m2_k, day_of_year[D], month_of_day[D],
            #                 latitude[C], clear_sky[12][C])
            w = water_temperature
            doy = np.ascontiguousarray(w["day_of_year"], dtype=np.int32)
            mon = np.ascontiguousarray(w["month_of_day"], dtype=np.int32)
            if doy.shape != (self.D,) or mon.shape != (self.D,):
                raise ValueError("day_of_year and month_of_day must have one entry per simulated day")
            self._keep += [doy, mon]
            cfg.water_temperature = WaterTemperature(1, 0, float(w.get("shade_proportion", 1.0)), float(w.get("albedo", 0.0)),
                                                     float(w.get("emissivity", 0.97)), float(w.get("convection_w_m2_k", 0.0)),
                                                     doy.ctypes.data_as(_ip), mon.ctypes.data_as(_ip))
            if "latitude" in w:
                self._location = (_d(w["latitude"]), _d(w["clear_sky"]))
        self._cfg = cfg
        self._plan = ct.c_void_p()
        if _create:
            check(lib().mms_chain_create(ct.byref(cfg), ct.byref(self._plan)))
            if self._location is not None:
                self.set_location(*self._location)

    def set_location(self, latitude, clear_sky):
        """mms_chain_set_location(): latitude [C] and clear-sky radiation by calendar month [12][C]."""
        lat, cs = _d(latitude), _d(clear_sky)
        if lat.shape != (self.C,) or cs.shape != (12, self.C):
            raise ValueError(f"latitude must be [{self.C}] and clear_sky [12][{self.C}]")
        check(lib().mms_chain_set_location(self._plan, _p(lat), _p(cs)))

    # -- output buffers -------------------------------------------------------
    def allocate_outputs(self, pinned=False):
        ms, mf = split_mask(self.mean_mask)
        hs, hf = split_mask(self.hourly_mask)
        shapes = {
            "mean_state": ((self.S, len(ms), self.P, self.R), np.float64) if ms and self.P else None,
            "mean_forcing": ((len(mf), self.P, self.R), np.float64) if mf and self.P else None,
            "hourly_state": ((self.S, len(hs), self.H, self.R), np.float64) if hs else None,
            "hourly_forcing": ((len(hf), self.H, self.R), np.float64) if hf else None,
            "wet_hours": ((self.S, self.P, self.R), np.int32) if self.want_wet_hours and self.P else None,
            "wet_flags": ((self.S, self.H, self.R), np.uint8) if self.want_wet_flags else None,
        }
        out, keep = {}, []
        for k, spec in shapes.items():
            if spec is None:
                continue
            if pinned:
                pa = PinnedArray(*spec)
                keep.append(pa)
                out[k] = pa.array
            else:
                out[k] = np.empty(spec[0], dtype=spec[1])
        out["_pinned"] = keep
        out["mean_state_vars"] = [VAR_NAMES[v] for v in ms]
        out["mean_forcing_vars"] = [VAR_NAMES[v] for v in mf]
        out["hourly_state_vars"] = [VAR_NAMES[v] for v in hs]
        out["hourly_forcing_vars"] = [VAR_NAMES[v] for v in hf]
        return out

    @staticmethod
    def _outputs_struct(out) -> ChainOutputs:
        o = ChainOutputs()
        for k in ("mean_state", "mean_forcing", "hourly_state", "hourly_forcing"):
            if k in out:
                setattr(o, k, out[k].ctypes.data_as(_dp))
        if "wet_hours" in out:
            o.wet_hours = out["wet_hours"].ctypes.data_as(_ip)
        if "wet_flags" in out:
            o.wet_flags = out["wet_flags"].ctypes.data_as(_bp)
        return o

    # -- the ABI calls --------------------------------------------------------
    def upload(self, monthly, altitude):
        monthly, altitude = _d(monthly), _d(altitude)
        if monthly.shape != (NFORCING, self.M, self.C) or altitude.shape != (self.C,):
            raise ValueError(f"monthly must be [9][{self.M}][{self.C}] and altitude [{self.C}]")
        self._inputs = (monthly, altitude)
        check(lib().mms_chain_upload(self._plan, _p(monthly), _p(altitude)))

    def upload_terraclimate(self, terraclimate, clear_sky, month_id, n_days, altitude):
        """mms_chain_upload_terraclimate(): the six raw TerraClimate variables [6][M][C] and the pixel's
        clear-sky radiation by calendar month [12][C]; the nine forcing rows are derived on the device."""
        tc, cs, nd, altitude = _d(terraclimate), _d(clear_sky), _d(n_days), _d(altitude)
        mid = np.ascontiguousarray(month_id, dtype=np.int32)
        if tc.shape != (NTERRACLIMATE, self.M, self.C) or cs.shape != (12, self.C) or altitude.shape != (self.C,):
            raise ValueError(f"terraclimate must be [6][{self.M}][{self.C}], clear_sky [12][{self.C}], altitude [{self.C}]")
        if mid.shape != (self.M,) or nd.shape != (self.M,):
            raise ValueError("month_id and n_days must have one entry per monthly knot")
        self._inputs = (tc, cs, mid, nd, altitude)
        check(lib().mms_chain_upload_terraclimate(self._plan, _p(tc), _p(cs), mid.ctypes.data_as(_ip), _p(nd),
                                                  _p(altitude)))

    def launch(self):
        check(lib().mms_chain_launch(self._plan))

    def sync(self):
        check(lib().mms_chain_sync(self._plan))

    def download(self, out):
        o = self._outputs_struct(out)
        check(lib().mms_chain_download(self._plan, ct.byref(o)))
        return out

    def kernel_ms(self) -> float:
        ms = ct.c_float()
        check(lib().mms_chain_kernel_ms(self._plan, ct.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = ct.c_int64()
        check(lib().mms_chain_launch_count(self._plan, ct.byref(n)))
        return int(n.value)

    def run(self, monthly, altitude, out=None):
        """mms_chain_run(): upload, kernels and download pipelined over the chunks of every shard."""
        monthly, altitude = _d(monthly), _d(altitude)
        if monthly.shape != (NFORCING, self.M, self.C) or altitude.shape != (self.C,):
            raise ValueError(f"monthly must be [9][{self.M}][{self.C}] and altitude [{self.C}]")
        out = out if out is not None else self.allocate_outputs()
        o = self._outputs_struct(out)
        check(lib().mms_chain_run(self._plan, _p(monthly), _p(altitude), ct.byref(o)))
        return out

    def run_terraclimate(self, terraclimate, clear_sky, month_id, n_days, altitude, out=None):
        """mms_chain_run_terraclimate(): the same from the six raw TerraClimate variables."""
        tc, cs, nd, altitude = _d(terraclimate), _d(clear_sky), _d(n_days), _d(altitude)
        mid = np.ascontiguousarray(month_id, dtype=np.int32)
        if tc.shape != (NTERRACLIMATE, self.M, self.C) or cs.shape != (12, self.C) or altitude.shape != (self.C,):
            raise ValueError(f"terraclimate must be [6][{self.M}][{self.C}], clear_sky [12][{self.C}], altitude [{self.C}]")
        out = out if out is not None else self.allocate_outputs()
        o = self._outputs_struct(out)
        check(lib().mms_chain_run_terraclimate(self._plan, _p(tc), _p(cs), mid.ctypes.data_as(_ip), _p(nd),
                                               _p(altitude), ct.byref(o)))
        return out

    def info(self):
        """(shards, chunks of the first shard, host threads per shard)"""
        a, b, c = ct.c_int32(), ct.c_int32(), ct.c_int32()
        check(lib().mms_chain_plan_info(self._plan, ct.byref(a), ct.byref(b), ct.byref(c)))
        return {"shards": a.value, "chunks_per_shard": b.value, "workers_per_shard": c.value}

    def close(self):
        if getattr(self, "_plan", None):
            lib().mms_chain_destroy(self._plan)
            self._plan = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_chain(monthly, altitude, **kwargs):
    """One-shot mms_run_chain(): host buffers in, host buffers out."""
    monthly = _d(monthly)
    if kwargs.get("water_temperature") is not None:
        # the water-temperature EXTENSION takes per-cell location inputs: create, set_location, run, destroy
        with ChainPlan(monthly.shape[2], **kwargs) as plan:
            return plan.run(monthly, altitude)
    with ChainPlan(monthly.shape[2], _create=False, **kwargs) as plan:
        out = plan.allocate_outputs()
        o = plan._outputs_struct(out)
        altitude = _d(altitude)
        check(lib().mms_run_chain(ct.byref(plan._cfg), _p(monthly), _p(altitude), ct.byref(o)))
        return out


def run_chain_terraclimate(terraclimate, clear_sky, month_id, n_days, altitude, **kwargs):
    """One-shot mms_run_chain_terraclimate(): the raw TerraClimate variables in, summaries out."""
    tc, cs, nd, altitude = _d(terraclimate), _d(clear_sky), _d(n_days), _d(altitude)
    mid = np.ascontiguousarray(month_id, dtype=np.int32)
    with ChainPlan(tc.shape[2], _create=False, **kwargs) as plan:
        if tc.shape != (NTERRACLIMATE, plan.M, plan.C) or cs.shape != (12, plan.C) or mid.shape != (plan.M,):
            raise ValueError("terraclimate must be [6][M][C], clear_sky [12][C], month_id [M]")
        out = plan.allocate_outputs()
        o = plan._outputs_struct(out)
        check(lib().mms_run_chain_terraclimate(ct.byref(plan._cfg), _p(tc), _p(cs), mid.ctypes.data_as(_ip), _p(nd),
                                               _p(altitude), ct.byref(o)))
        return out


def water_vectorised(rainfall, air_temperature, humidity, windspeed, altitude, habitat, burnin_years=0,
                     math_mode="fast", device=0, want_volume=False, want_wet=False, whole_series_burnin=False):
    """mms_water_vectorised(), or mms_water_series() (the scalar twin's burn-in rule: the whole
    series is repeated) when whole_series_burnin is set."""
    mats = [np.asfortranarray(m, dtype=np.float64) for m in (rainfall, air_temperature, humidity, windspeed)]
    H, C = mats[0].shape
    if any(m.shape != (H, C) for m in mats):
        raise ValueError("all matrices must have the same dimensions")
    alt = _d(altitude)
    if alt.shape != (C,):
        raise ValueError("altitude_vector must have one element per pixel")
    hb = habitat if isinstance(habitat, Habitat) else make_habitat(**habitat)
    area = np.empty((H, C), order="F")
    vol = np.empty((H, C), order="F") if want_volume else None
    wet = np.empty((H, C), order="F", dtype=np.uint8) if want_wet else None
    fn = lib().mms_water_series if whole_series_burnin else lib().mms_water_vectorised
    check(fn(*[m.ctypes.data_as(_dp) for m in mats], _p(alt), H, C, ct.byref(hb),
             int(burnin_years), {"fast": 0, "strict": 1}.get(math_mode, math_mode),
             int(device), area.ctypes.data_as(_dp),
             None if vol is None else vol.ctypes.data_as(_dp),
             None if wet is None else wet.ctypes.data_as(_bp)))
    return {"water_surface_area": area, "water_volume": vol, "wet": wet}


def population_vectorised(larval_habitat_area, mdr, efd, das_zerodensity, ds, densmod, initial_adult=100.0,
                          initial_aquatic=100.0, burnin_years=1, math_mode="fast", device=0,
                          whole_series_burnin=False):
    """mms_population_vectorised(), or mms_population_series() when whole_series_burnin is set."""
    mats = [np.asfortranarray(m, dtype=np.float64) for m in (larval_habitat_area, mdr, efd, das_zerodensity, ds)]
    H, C = mats[0].shape
    if any(m.shape != (H, C) for m in mats):
        raise ValueError("all matrices must have the same dimensions")
    dm = densmod if isinstance(densmod, Densmod) else make_densmod(**densmod)
    aquatic = np.empty((H, C), order="F")
    adult = np.empty((H, C), order="F")
    fn = lib().mms_population_series if whole_series_burnin else lib().mms_population_vectorised
    check(fn(*[m.ctypes.data_as(_dp) for m in mats], H, C, ct.byref(dm),
             float(initial_adult), float(initial_aquatic), int(burnin_years),
             {"fast": 0, "strict": 1}.get(math_mode, math_mode), int(device),
             aquatic.ctypes.data_as(_dp), adult.ctypes.data_as(_dp)))
    return {"aquatic": aquatic, "adult": adult}


def _mode(math_mode) -> int:
    return {"fast": 0, "strict": 1}.get(math_mode, math_mode)


def terraclimate_to_monthly(terraclimate, clear_sky, month_id, n_days, math_mode="fast", device=0):
    """mms_terraclimate_to_monthly(): terraclimate [6][M][C], clear_sky [12][C] -> monthly [9][M][C]."""
    tc, cs, nd = _d(terraclimate), _d(clear_sky), _d(n_days)
    mid = np.ascontiguousarray(month_id, dtype=np.int32)
    _, M, C = tc.shape
    if tc.shape[0] != NTERRACLIMATE or cs.shape != (12, C) or mid.shape != (M,) or nd.shape != (M,):
        raise ValueError("terraclimate must be [6][M][C], clear_sky [12][C], month_id and n_days [M]")
    monthly = np.empty((NFORCING, M, C))
    check(lib().mms_terraclimate_to_monthly(_p(tc), _p(cs), mid.ctypes.data_as(_ip), _p(nd), C, M, _mode(math_mode),
                                            int(device), _p(monthly)))
    return monthly


def summarise_vectors(hourly, period_of_day, n_periods, device=0):
    """mms_summarise_vectors(): hourly [n_vars][24 D][C] -> period means [n_vars][P][C]."""
    hourly = _d(hourly)
    nv, H, C = hourly.shape
    pod = np.ascontiguousarray(period_of_day, dtype=np.int32)
    if H != 24 * len(pod):
        raise ValueError("hourly must have 24 steps per entry of period_of_day")
    means = np.empty((nv, int(n_periods), C))
    check(lib().mms_summarise_vectors(_p(hourly), C, len(pod), nv, pod.ctypes.data_as(_ip), int(n_periods), int(device),
                                      _p(means)))
    return means


def daily_from_monthly(monthly, mid_day, day_predict, math_mode="fast", device=0):
    """mms_daily_from_monthly(): monthly [9][M][C] -> daily [9][D][C]."""
    monthly, mid_day, day_predict = _d(monthly), _d(mid_day), _d(day_predict)
    _, M, C = monthly.shape
    D = len(day_predict)
    if len(mid_day) != M:
        raise ValueError("mid_day must have one entry per monthly knot")
    daily = np.empty((NFORCING, D, C))
    check(lib().mms_daily_from_monthly(_p(monthly), _p(mid_day), _p(day_predict), C, M, D, _mode(math_mode),
                                       int(device), _p(daily)))
    return daily


def hourly_from_daily(daily, math_mode="fast", device=0):
    """mms_hourly_from_daily(): daily [9][D][C] -> dict of [24 D][C] arrays."""
    daily = _d(daily)
    _, D, C = daily.shape
    hourly = np.empty((6, 24 * D, C))
    check(lib().mms_hourly_from_daily(_p(daily), C, D, _mode(math_mode), int(device), _p(hourly)))
    names = ("air_temperature", "humidity", "cloudcover", "windspeed", "water_temperature", "rainfall")
    return {n: hourly[k] for k, n in enumerate(names)}


def lifehistory_vectorised(water_temperature, air_temperature, humidity, tables, math_mode="fast", device=0):
    """mms_lifehistory_vectorised(): element-wise mdr, efd, das_zerodensity, ds (any array shape)."""
    shape = np.shape(water_temperature)
    tw, ta, rh = (_d(a).ravel() for a in (water_temperature, air_temperature, humidity))
    if not (tw.size == ta.size == rh.size):
        raise ValueError("inputs must have the same number of elements")
    tabs = {k: {kk: _d(vv) for kk, vv in v.items()} for k, v in tables.items()}

    def sp(t):
        return Spline(len(t["x"]), 0, _p(t["x"]), _p(t["y"]), _p(t["b"]), _p(t["c"]), _p(t["d"]))

    g = tabs["ds"]
    lh = LifehistoryC(sp(tabs["mdr"]), sp(tabs["efd"]), sp(tabs["das"]),
                      Surface(len(g["x"]), len(g["y"]), _p(g["x"]), _p(g["y"]), _p(g["z"])))
    outs = [np.empty(tw.size) for _ in range(4)]
    check(lib().mms_lifehistory_vectorised(_p(tw), _p(ta), _p(rh), tw.size, ct.byref(lh), _mode(math_mode),
                                           int(device), *[_p(o) for o in outs]))
    names = ("mdr", "efd", "das_zerodensity", "ds")
    return {n: o.reshape(shape) for n, o in zip(names, outs)}


def stage_kernel_ms() -> float:
    """Device time of the kernel of the last stage-level twin call on this thread."""
    ms = ct.c_float()
    check(lib().mms_stage_kernel_ms(ct.byref(ms)))
    return float(ms.value)


def measure_fp64_peak(device=0) -> float:
    v = ct.c_double()
    check(lib().mms_measure_fp64_peak(int(device), ct.byref(v)))
    return float(v.value)


def measure_hbm_copy(device=0) -> float:
    v = ct.c_double()
    check(lib().mms_measure_hbm_copy(int(device), ct.byref(v)))
    return float(v.value)


def device_info(device=0) -> dict:
    n_sm, khz, mem = ct.c_int32(), ct.c_int32(), ct.c_uint64()
    check(lib().mms_device_info(int(device), ct.byref(n_sm), ct.byref(khz), ct.byref(mem)))
    return {"n_sm": n_sm.value, "sm_clock_mhz_max": khz.value / 1000.0, "memory_bytes": mem.value}


def flush_l2(device=0) -> None:
    check(lib().mms_flush_l2(int(device)))
