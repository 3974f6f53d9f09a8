"""Host-side mirrors of the reference's *runtime closures* for life-history
parameters (R/lifehistory_estimation_functions.R).

In the reference these are R closures evaluated on the CPU for every
cell-hour.  Here they are plain table carriers: construction happens once on
the host (exactly where the reference constructs them -- at package build /
call set-up time), evaluation happens inside the CUDA kernels.  None of these
objects is callable: there is no CPU evaluation path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np


@dataclass(frozen=True)
class SplineTable:
    """Coefficients of stats::splinefun(x, y) (default method "fmm"): on
    [x[i], x[i+1]] the value is y[i] + dx*(b[i] + dx*(c[i] + dx*d[i])).
    This is the list `environment(f)$z` of an R splinefun closure."""
    x: np.ndarray
    y: np.ndarray
    b: np.ndarray
    c: np.ndarray
    d: np.ndarray

    def as_dict(self):
        return {"x": self.x, "y": self.y, "b": self.b, "c": self.c, "d": self.d}


def splinefun_fmm(x, y) -> SplineTable:
    """Forsythe-Malcolm-Moler cubic spline coefficients, as computed by base R's
    splinefun(x, y) (method "fmm"; call site
    R/lifehistory_estimation_functions.R:243).  Set-up code: runs once per
    table on the host, like the R call it mirrors."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = len(x)
    if n < 2 or len(y) != n:
        raise ValueError("splinefun_fmm needs at least two (x, y) pairs of equal length")
    if np.any(np.diff(x) <= 0):
        raise ValueError("x must be strictly increasing")
    b = np.zeros(n)
    c = np.zeros(n)
    d = np.zeros(n)
    if n < 3:
        b[:] = (y[1] - y[0]) / (x[1] - x[0])
        return SplineTable(x, y, b, c, d)
    # tridiagonal system: b = diagonal, d = off-diagonal, c = right-hand side
    d[:-1] = np.diff(x)
    slope = np.diff(y) / d[:-1]
    b[1:-1] = 2 * (d[:-2] + d[1:-1])
    c[1:-1] = slope[1:] - slope[:-1]
    b[0] = -d[0]
    b[-1] = -d[-2]
    if n > 3:
        # third derivatives at the ends from divided differences
        c0 = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
        cn = c[-2] / (x[-1] - x[-3]) - c[-3] / (x[-2] - x[-4])
        c[0] = c0 * d[0] * d[0] / (x[3] - x[0])
        c[-1] = -cn * d[-2] * d[-2] / (x[-1] - x[-4])
    for i in range(1, n):                       # forward elimination
        t = d[i - 1] / b[i - 1]
        b[i] = b[i] - t * d[i - 1]
        c[i] = c[i] - t * c[i - 1]
    c[-1] = c[-1] / b[-1]                       # back substitution
    for i in range(n - 2, -1, -1):
        c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
    h = np.diff(x)
    b[-1] = (y[-1] - y[-2]) / h[-1] + h[-1] * (c[-2] + 2 * c[-1])
    b[:-1] = (y[1:] - y[:-1]) / h - h * (c[1:] + 2 * c[:-1])
    dd = (c[1:] - c[:-1]) / h
    c = 3 * c
    d[:-1] = dd
    d[-1] = d[-2]
    return SplineTable(x, y, b, c, d)


@dataclass(frozen=True)
class PositiveSpline:
    """positive_spline(pred, temps_out), R/lifehistory_estimation_functions.R:236-250:
    an fmm spline through (temps_out, pred), clamped with ensure_positive()."""
    table: SplineTable


def positive_spline(pred, temps_out) -> PositiveSpline:
    return PositiveSpline(splinefun_fmm(temps_out, pred))


@dataclass(frozen=True)
class SurvDensmod:
    """make_surv_densmod_function(dd_effect, surface_area_cm2, type),
    R/lifehistory_estimation_functions.R:525-587.  `type` is the closure's
    attr(fun, "type")."""
    dd_effect: float
    surface_area_cm2: float
    type: str = "cox_ph"

    @property
    def type_code(self) -> int:
        return {"logit": 0, "cox_ph": 1}[self.type]


def make_surv_densmod_function(dd_effect, surface_area_cm2, type="cox_ph") -> SurvDensmod:
    if type not in ("cox_ph", "logit"):          # match.arg
        raise ValueError("'arg' should be one of 'cox_ph', 'logit'")
    return SurvDensmod(float(dd_effect), float(surface_area_cm2), type)


@dataclass(frozen=True)
class TempHumidSurface:
    """The `interpolation_info` list of make_temp_humid_interpolator(),
    R/lifehistory_estimation_functions.R:1144-1190: x = temperature steps,
    y = humidity steps, z[ix + nx*iy] (R column-major matrix)."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray

    def as_dict(self):
        return {"x": self.x, "y": self.y, "z": self.z}


def r_seq_length_out(start: float, stop: float, n: int) -> np.ndarray:
    """base-R seq(from, to, length.out = n) for doubles: exact end points,
    from + i*by in between."""
    out = start + np.arange(n, dtype=np.float64) * ((stop - start) / (n - 1))
    out[0] = start
    out[-1] = stop
    return out


def make_temp_humid_interpolator(original_function: Callable[[np.ndarray, np.ndarray], np.ndarray],
                                 temperature_limits=(0.0, 45.0), humidity_limits=(0.0, 100.0),
                                 grid_resolution=250) -> TempHumidSurface:
    """Tabulate original_function(temperature, humidity) on the reference's
    grid (set-up time, once); the bilinear evaluation
    (fields::interp.surface) happens on the GPU."""
    tx = r_seq_length_out(temperature_limits[0], temperature_limits[1], grid_resolution)
    hy = r_seq_length_out(humidity_limits[0], humidity_limits[1], grid_resolution)
    # expand.grid(temperature, humidity): temperature varies fastest
    tt = np.tile(tx, grid_resolution)
    hh = np.repeat(hy, grid_resolution)
    z = np.ascontiguousarray(original_function(tt, hh), dtype=np.float64)
    if z.shape != (grid_resolution * grid_resolution,):
        raise ValueError("original_function must return one value per grid point")
    return TempHumidSurface(tx, hy, z)


@dataclass(frozen=True)
class Lifehistory:
    """One element of the reference's `lifehistory_functions` list (e.g.
    lifehistory_functions$An_gambiae; data-raw/estimate_lifehistory_parameters.R:813-840)
    with ds_temp_humid already replaced by its bilinear emulator
    (fast_ds_temp_humid = TRUE, R/pipeline_functions.R:649-653)."""
    mdr_temp: PositiveSpline
    efd_temp: PositiveSpline
    das_temp: PositiveSpline
    ds_temp_humid: TempHumidSurface
    das_densmod: SurvDensmod

    def tables(self):
        return {"mdr": self.mdr_temp.table.as_dict(), "efd": self.efd_temp.table.as_dict(),
                "das": self.das_temp.table.as_dict(), "ds": self.ds_temp_humid.as_dict()}
