"""Shared builders and the parity metric for the test-suite."""
from __future__ import annotations

import numpy as np

import oracle
from mosmicrosim_b200 import dates, synthetic as syn

REL_TOL_FAST = 1e-9      # north_star: fp64 state within 1e-9 relative error
REL_TOL_STRICT = 1e-11   # strict mode differs from the oracle only by libm last bits
ABS_FLOOR = 1e-15        # absolute floor (about 1e-7 * sqrt(eps)): cancellation noise


def rel_err(got, ref, floor=ABS_FLOOR):
    """max |got - ref| / max(|ref|, floor / tol-equivalent); NaN patterns must match."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    if not np.array_equal(nan_g, nan_r):
        return np.inf
    ok = ~nan_r
    if not ok.any():
        return 0.0
    diff = np.abs(got[ok] - ref[ok])
    diff = np.where(np.isinf(ref[ok]) & (got[ok] == ref[ok]), 0.0, diff)
    return float(np.max(np.maximum(diff - floor, 0.0) / np.maximum(np.abs(ref[ok]), 1e-300)))


class Case:
    """A small synthetic tile with everything both sides need."""

    def __init__(self, n_cells=96, years=1, first_year=2001, seed=syn.SEED, species="An. gambiae",
                 habitats=None, aggregate_by="month", included=None, default_dates=False,
                 out_of_range_fraction=0.05):
        self.lh = syn.synthetic_lifehistory(species)
        self.tables = self.lh.tables()
        self.start, self.end = syn.month_table(first_year, years)
        if default_dates:
            self.dates_predict = dates.default_dates_predict(self.start, self.end)
        else:
            self.dates_predict = syn.year_dates(first_year, years)
        self.mid_day, self.day_predict = dates.datenums(self.start, self.end, self.dates_predict)
        self.monthly, self.altitude = syn.synthetic_forcing(n_cells, self.start, seed=seed,
                                                            out_of_range_fraction=out_of_range_fraction)
        self.period_of_day, self.p_start, self.p_end = dates.period_index(self.dates_predict, aggregate_by, included)
        self.P = len(self.p_start)
        self.habitats = habitats or [syn.habitat_cone()]
        dm = self.lh.das_densmod
        self.densmod = dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2)
        self.C, self.D, self.H = n_cells, len(self.dates_predict), 24 * len(self.dates_predict)

    def oracle(self, want_hourly=True, burnin_years_water=0, burnin_years_pop=1):
        return oracle.run_chain(self.monthly, self.mid_day, self.day_predict, self.altitude,
                                [oracle.habitat(**h) for h in self.habitats],
                                oracle.densmod(self.densmod["type_"], self.densmod["dd_effect"],
                                               self.densmod["surface_area_cm2"]),
                                self.tables, self.period_of_day, self.P,
                                burnin_years_water=burnin_years_water, burnin_years_pop=burnin_years_pop,
                                want_hourly=want_hourly)

    def gpu_kwargs(self, **over):
        kw = dict(mid_day=self.mid_day, day_predict=self.day_predict, habitats=self.habitats,
                  densmod=self.densmod, tables=self.tables, period_of_day=self.period_of_day,
                  n_periods=self.P)
        kw.update(over)
        return kw


# oracle variable id (12 vars) for each product variable name
ORACLE_VAR = {n: i for i, n in enumerate(oracle.VAR_NAMES)}


# ---- the plain-C caller of the ABI (tests/c_driver/abi_driver.c) ------------------------------------------

def build_c_driver(tmp_path) -> str:
    """gcc -std=c99 -Wall -Wextra -Werror on the driver, linked against the in-tree libmmsgpu.so."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.join(root, "mosmicrosim_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c_driver", "abi_driver.c"), "-o", exe, "-L", libdir, "-lmmsgpu",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def write_c_driver_inputs(case, names, d, ns, nf):
    """The driver's input directory: dims.txt and raw little-endian arrays / structs."""
    from mosmicrosim_b200 import backend
    t = case.tables
    (d / "dims.txt").write_text(f"{case.C} {len(case.mid_day)} {case.D} {len(case.habitats)} {case.P} "
                                f"{len(t['mdr']['x'])} {len(t['ds']['x'])} {len(t['ds']['y'])} "
                                f"{backend.var_mask(names)} {ns} {nf}\n")

    def f64(name, a):
        np.ascontiguousarray(a, dtype=np.float64).tofile(str(d / name))

    f64("mid_day.f64", case.mid_day); f64("day_predict.f64", case.day_predict)
    np.ascontiguousarray(case.period_of_day, dtype=np.int32).tofile(str(d / "period_of_day.i32"))
    (d / "habitats.bin").write_bytes(b"".join(bytes(backend.make_habitat(**h)) for h in case.habitats))
    (d / "densmod.bin").write_bytes(bytes(backend.make_densmod(**case.densmod)))
    for stem in ("mdr", "efd", "das"):
        for part in "xybcd":
            f64(f"{stem}_{part}.f64", t[stem][part])
    f64("ds_x.f64", t["ds"]["x"]); f64("ds_y.f64", t["ds"]["y"]); f64("ds_z.f64", t["ds"]["z"])
    f64("monthly.f64", case.monthly); f64("altitude.f64", case.altitude)
