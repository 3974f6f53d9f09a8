"""Deterministic synthetic inputs for tests and benchmarks.

The reference's fitted life-history tables (R/sysdata.rda) and any real
TerraClimate forcing are unavailable offline, so every measured or tested
number in this repo is on "synthetic tables / synthetic forcing": curves of
the functional forms the reference fits (Briere, concave quadratic, bell,
smooth surface) sampled at the reference's knots, and TerraClimate-style
monthly series with a seasonal cycle.
"""
from __future__ import annotations

import numpy as np

from . import lifehistory as lh

SEED = 20251218     # echoes data-raw/estimate_lifehistory_parameters.R:49

# closed-form constants of the reference
DD_EFFECT_AG = -0.003935158408341286       # mean logit slope, Muriu rows
MURIU_AREA_CM2 = 962.1127501618741
EVANS_AREA_CM2 = 28.58146874284323


def synthetic_lifehistory(species="An. gambiae") -> lh.Lifehistory:
    """Stand-ins for lifehistory_functions$An_gambiae / $An_stephensi with the
    reference's functional forms and knots (temps_out = seq(-20, 80, by = 0.1),
    R/lifehistory_estimation_functions.R:259)."""
    temps_out = -20.0 + np.arange(1001, dtype=np.float64) * 0.1
    t = temps_out
    shift = 0.0 if species == "An. gambiae" else 1.5
    # Briere: c T (T - T0) sqrt(Tm - T) on (T0, Tm), else 0  (:68-76)
    t0, tm, c = 14.0 + shift, 38.5 + shift, 1.1e-4
    inside = (t > t0) & (t < tm)
    mdr = np.where(inside, c * t * (t - t0) * np.sqrt(np.where(inside, tm - t, 0.0)), 0.0)
    # concave quadratic: -qd (T - T0)(T - Tm) on (T0, Tm), else 0  (:77-84)
    e0, em, qd = 15.0 + shift, 36.0 + shift, 0.09
    efd = np.where((t > e0) & (t < em), -qd * (t - e0) * (t - em), 0.0)
    # daily aquatic survival at zero density: smooth bell inside (0, 1)
    das = 0.95 * np.exp(-((t - (25.0 + shift)) / 11.0) ** 2)

    def ds_fun(temperature, humidity):
        return 0.9 * np.exp(-((temperature - (24.0 + shift)) / 14.0) ** 2) * (0.55 + 0.45 * humidity / 100.0)

    if species == "An. gambiae":
        dens = lh.make_surv_densmod_function(DD_EFFECT_AG, MURIU_AREA_CM2, "logit")
    else:
        dens = lh.make_surv_densmod_function(0.002, EVANS_AREA_CM2, "cox_ph")
    return lh.Lifehistory(
        mdr_temp=lh.positive_spline(mdr, temps_out),
        efd_temp=lh.positive_spline(efd, temps_out),
        das_temp=lh.positive_spline(das, temps_out),
        ds_temp_humid=lh.make_temp_humid_interpolator(ds_fun),
        das_densmod=dens,
    )


def month_table(first_year: int, years: int):
    """Monthly knots from the December before first_year to the January after
    the last year (M = 12*years + 2), i.e. what monthly_summary_days() with its
    +-21-day buffer produces for whole-year requests (R/terraclimate.R:91-98).
    Returns (start[M], end[M]) as datetime64[D]."""
    first = np.datetime64(f"{first_year - 1:04d}-12", "M")
    months = first + np.arange(12 * years + 2)
    start = months.astype("datetime64[D]")
    end = (months + 1).astype("datetime64[D]") - np.timedelta64(1, "D")
    return start, end


def year_dates(first_year: int, years: int) -> np.ndarray:
    """Every calendar day of the simulated years."""
    a = np.datetime64(f"{first_year:04d}-01-01", "D")
    b = np.datetime64(f"{first_year + years:04d}-01-01", "D")
    return np.arange(a, b)


def synthetic_forcing(n_cells: int, start, seed: int = SEED, out_of_range_fraction: float = 0.01,
                      dtype=np.float64):
    """TerraClimate-style monthly model variables for n_cells cells.
    Returns (monthly[9][M][C] cell-fastest, altitude[C]); variable order is the
    reference's column order tmin, tmax, rhmin, rhmax, wsmin, wsmax, ccmin,
    ccmax, rainfall_daily (R/terraclimate.R:633-642)."""
    start = np.asarray(start, dtype="datetime64[D]")
    M = len(start)
    rng = np.random.default_rng(seed)
    C = int(n_cells)
    month_of_year = (start.astype("datetime64[M]").astype(np.int64) % 12).astype(np.float64)
    phase = rng.uniform(0, 2 * np.pi, size=C)
    season = np.sin(2 * np.pi * month_of_year[:, None] / 12.0 + phase[None, :])      # [M, C]
    monthly = np.empty((9, M, C), dtype=dtype)
    tmax = 30.0 + 5.0 * season + rng.normal(0.0, 1.0, size=(M, C))
    # a small fraction of cells leave the [0, 45] C range of the ds surface (NA path)
    n_bad = int(round(out_of_range_fraction * C))
    if n_bad:
        bad = rng.choice(C, size=n_bad, replace=False)
        tmax[:, bad] += np.where(rng.random(n_bad) < 0.5, 14.0, -24.0)[None, :]
    tmin = tmax - rng.uniform(8.0, 14.0, size=(M, C))
    rhmax = rng.uniform(60.0, 100.0, size=(M, C))
    rhmin = rhmax * rng.uniform(0.3, 0.8, size=(M, C))
    wsmax = rng.uniform(0.5, 5.0, size=(M, C))
    wsmin = wsmax * 0.1                                     # R/terraclimate.R:618
    cloud = rng.uniform(0.0, 80.0, size=(M, C))
    ccmax = np.clip(cloud * 1.2, 0.0, 100.0)                # R/terraclimate.R:624-625
    ccmin = np.clip(cloud * 0.8, 0.0, 100.0)
    wet_season = np.clip(0.5 + 0.8 * np.sin(2 * np.pi * month_of_year[:, None] / 12.0 + phase[None, :] + 1.0), 0.0, None)
    rain = rng.gamma(shape=2.0, scale=2.0, size=(M, C)) * wet_season     # mean 0-8 mm/day
    for v, arr in enumerate((tmin, tmax, np.clip(rhmin, 0.01, 100.0), np.clip(rhmax, 0.01, 100.0),
                             wsmin, wsmax, ccmin, ccmax, rain)):
        monthly[v] = arr
    altitude = rng.uniform(0.0, 2500.0, size=C)
    return monthly, altitude


def synthetic_forcing_large(n_cells: int, start, seed: int = SEED, base_cells: int = 100_000,
                            out_of_range_fraction: float = 0.01):
    """The same kind of forcing for the multi-million-cell configurations, generated fast: a base tile of
    base_cells cells from synthetic_forcing(), repeated with a replica-specific temperature shift, rainfall
    scale and altitude offset (so no two replicas are the same cells).  Returns (monthly[9][M][C], altitude[C])."""
    C = int(n_cells)
    base, alt = synthetic_forcing(min(base_cells, C), start, seed=seed, out_of_range_fraction=out_of_range_fraction)
    nb = base.shape[2]
    if nb == C:
        return base, alt
    monthly = np.empty((9, base.shape[1], C))
    altitude = np.empty(C)
    for k, c0 in enumerate(range(0, C, nb)):
        n = min(nb, C - c0)
        monthly[:, :, c0:c0 + n] = base[:, :, :n]
        monthly[0:2, :, c0:c0 + n] += 0.37 * ((k * 7) % 11 - 5)          # tmin, tmax: -1.85 .. +1.85 deg C
        monthly[8, :, c0:c0 + n] *= 1.0 + 0.05 * ((k * 5) % 9 - 4)        # rainfall: x 0.8 .. 1.2
        altitude[c0:c0 + n] = (alt[:n] + 137.0 * k) % 2500.0
    return monthly, altitude


def synthetic_terraclimate(n_cells: int, start, seed: int = SEED, out_of_range_fraction: float = 0.01):
    """Raw TerraClimate-style monthly variables for n_cells cells:
    returns (terraclimate[6][M][C] in the order tmax, tmin, ppt, ws, vpd, srad; clear_sky[12][C];
    altitude[C]).  Ranges follow TerraClimate's units: deg C, mm / month, m/s at 10 m, kPa, W/m2."""
    start = np.asarray(start, dtype="datetime64[D]")
    M, C = len(start), int(n_cells)
    rng = np.random.default_rng(seed + 1)
    month_of_year = (start.astype("datetime64[M]").astype(np.int64) % 12)
    phase = rng.uniform(0, 2 * np.pi, size=C)
    season = np.sin(2 * np.pi * month_of_year[:, None] / 12.0 + phase[None, :])
    tmax = 30.0 + 5.0 * season + rng.normal(0.0, 1.0, size=(M, C))
    n_bad = int(round(out_of_range_fraction * C))
    if n_bad:
        bad = rng.choice(C, size=n_bad, replace=False)
        tmax[:, bad] += np.where(rng.random(n_bad) < 0.5, 14.0, -24.0)[None, :]
    tmin = tmax - rng.uniform(8.0, 14.0, size=(M, C))
    wet_season = np.clip(0.5 + 0.8 * np.sin(2 * np.pi * month_of_year[:, None] / 12.0 + phase[None, :] + 1.0), 0.0, None)
    ppt = rng.gamma(shape=2.0, scale=60.0, size=(M, C)) * wet_season           # mm / month
    ws = rng.uniform(0.5, 7.0, size=(M, C))
    vpd = rng.uniform(0.05, 3.0, size=(M, C))                                  # some exceed esat(tmean): vp < 0 -> rh clamps
    clear_sky = rng.uniform(180.0, 330.0, size=(12, C))
    srad = clear_sky[month_of_year] * rng.uniform(0.3, 1.1, size=(M, C))       # > clear sky in places: cloud clamps at 0
    altitude = rng.uniform(0.0, 2500.0, size=C)
    return np.stack([tmax, tmin, ppt, ws, vpd, srad]), clear_sky, altitude


# habitat settings used by the benchmark configurations ----------------------

def habitat_cone(max_depth=1.0, inflow_multiplier=1.0, leak_per_hour=0.0, initial_volume=0.0):
    return dict(shape=0, max_depth=max_depth, radius=1.0, inflow_multiplier=inflow_multiplier,
                leak_per_hour=leak_per_hour, initial_volume=initial_volume)


def habitat_cylinder(radius=0.5, max_depth=0.5, inflow_multiplier=1.0, leak_per_hour=0.0, initial_volume=0.0):
    return dict(shape=1, max_depth=max_depth, radius=radius, inflow_multiplier=inflow_multiplier,
                leak_per_hour=leak_per_hour, initial_volume=initial_volume)


def habitat_ensemble(n: int = 64):
    """cfg5: depth x inflow x leak x shape grid."""
    out = []
    depths = (0.25, 0.5, 1.0, 2.0)
    inflows = (0.5, 1.0, 2.0, 4.0)
    leaks = (0.0, 0.002)
    for shape in (0, 1):
        for depth in depths:
            for inflow in inflows:
                for leak in leaks:
                    if shape == 0:
                        out.append(habitat_cone(depth, inflow, leak))
                    else:
                        out.append(habitat_cylinder(0.5, depth, inflow, leak))
    return out[:n]
