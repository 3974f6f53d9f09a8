"""Writes tests/golden/r_reference/inputs/: this repository's synthetic inputs as flat little-endian binaries for
baseline/run_reference.R (which runs the reference's own R functions on them; see its header for the layouts).

    python tests/golden/make_r_inputs.py

24 cells, one simulated year (2001) between December and January knots (no extrapolation branch), both species'
synthetic life-history tables (the fitted tables of the reference, R/sysdata.rda, are not available: SURVEY 0.3).
Two of the cells leave the 0-45 C grid of the ds surface (the NA path).  Deterministic: re-running reproduces the
committed files bit for bit."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mosmicrosim_b200 import synthetic as syn  # noqa: E402

N_CELLS = 24
SPECIES = {"An_gambiae": "An. gambiae", "An_stephensi": "An. stephensi"}


def inputs():
    start, end = syn.month_table(2001, 1)
    dates_predict = syn.year_dates(2001, 1)
    monthly, altitude = syn.synthetic_forcing(N_CELLS, start, seed=syn.SEED + 3, out_of_range_fraction=2.0 / N_CELLS)
    return start, end, dates_predict, monthly, altitude


def main(out_dir=os.path.join(HERE, "r_reference", "inputs")):
    os.makedirs(out_dir, exist_ok=True)
    start, end, dates_predict, monthly, altitude = inputs()
    days = lambda d: np.asarray(d, dtype="datetime64[D]").astype(np.int64).astype("<i4")
    days(start).tofile(os.path.join(out_dir, "start_days.i32"))
    days(end).tofile(os.path.join(out_dir, "end_days.i32"))
    days(dates_predict).tofile(os.path.join(out_dir, "dates_predict_days.i32"))
    monthly.astype("<f8").tofile(os.path.join(out_dir, "monthly.f64"))          # [9][M][C], cell fastest
    altitude.astype("<f8").tofile(os.path.join(out_dir, "altitude.f64"))
    n_knots = n_grid = None
    for d, species in SPECIES.items():
        lh = syn.synthetic_lifehistory(species)
        sd = os.path.join(out_dir, d)
        os.makedirs(sd, exist_ok=True)
        t = lh.tables()
        n_knots, n_grid = len(t["mdr"]["x"]), len(t["ds"]["x"])
        t["mdr"]["x"].astype("<f8").tofile(os.path.join(sd, "temps_out.f64"))
        for k in ("mdr", "efd", "das"):
            t[k]["y"].astype("<f8").tofile(os.path.join(sd, f"{k}_y.f64"))       # R recomputes the fmm coefficients
        t["ds"]["z"].astype("<f8").tofile(os.path.join(sd, "ds_z.f64"))          # [iT + 250 iRH], temperature fastest
        dm = lh.das_densmod
        with open(os.path.join(sd, "densmod.txt"), "w") as f:
            f.write(f"{dm.type} {dm.dd_effect!r} {dm.surface_area_cm2!r}\n")
    with open(os.path.join(out_dir, "dims.txt"), "w") as f:
        f.write(f"{N_CELLS} {len(start)} {len(dates_predict)} {n_knots} {n_grid}\n")
    return out_dir


if __name__ == "__main__":
    print(main())
