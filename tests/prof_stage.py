"""Timing driver for the stage-level twins and state-free verbs (development aid): achieved
algorithmic HBM GB/s of their kernels, CUDA events around the kernel (mms_stage_kernel_ms)."""
import sys
sys.path.insert(0, ".."); sys.path.insert(0, ".")
import numpy as np
from helpers import Case
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8760
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
what = sys.argv[4] if len(sys.argv) > 4 else "all"
rng = np.random.default_rng(1)


def report(name, ms, nbytes, note=""):
    print(f"{name:18s} {mode}: kernel {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s algorithmic {note}")


if what in ("all", "twins"):
    shape = (H, C)
    mk = lambda lo, hi: np.asfortranarray(rng.uniform(lo, hi, size=shape))
    rain, airt, rh, wind = mk(0, 0.4), mk(12, 36), mk(25, 100), mk(0.1, 5)
    alt = rng.uniform(0, 2500, C)
    for rep in range(3):
        w = backend.water_vectorised(rain, airt, rh, wind, alt, syn.habitat_cone(), math_mode=mode)
        report("water twin", backend.stage_kernel_ms(), 40.0 * H * C, "(40 B/cell-hour)")
    area = w["water_surface_area"]
    del rain, wind
    mdr, efd, das, ds = mk(0, 0.16), mk(0, 10), mk(0.2, 0.95), mk(0.3, 0.9)
    dm = backend.make_densmod("logit", syn.DD_EFFECT_AG, syn.MURIU_AREA_CM2)
    burn = 1 if H >= 8760 else 0
    for rep in range(3):
        backend.population_vectorised(area, mdr, efd, das, ds, dm, burnin_years=burn, math_mode=mode)
        report("population twin", backend.stage_kernel_ms(), 56.0 * H * C,
               f"(56 B/cell-hour over {H} recorded steps; {H + 8760 * burn} executed)")
    del area, mdr, efd, das, ds

if what in ("all", "verbs"):
    case = Case(n_cells=C, out_of_range_fraction=0.01)
    D = case.D
    for rep in range(3):
        daily = backend.daily_from_monthly(case.monthly, case.mid_day, case.day_predict, math_mode=mode)
        report("daily verb", backend.stage_kernel_ms(), 72.0 * D * C, "(9 doubles out per cell-day)")
    for rep in range(3):
        hourly = backend.hourly_from_daily(daily, math_mode=mode)
        report("hourly verb", backend.stage_kernel_ms(), (72.0 + 6 * 24 * 8.0) * D * C,
               "(9 in + 144 out doubles per cell-day)")
    tabs = case.lh.tables()
    for rep in range(3):
        backend.lifehistory_vectorised(hourly["water_temperature"], hourly["air_temperature"], hourly["humidity"],
                                       tabs, math_mode=mode)
        report("life-history verb", backend.stage_kernel_ms(), 56.0 * 24 * D * C, "(3 in + 4 out doubles per value)")
