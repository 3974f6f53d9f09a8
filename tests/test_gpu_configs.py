"""BASELINE.json configs[2], [3] and [4] at FULL size through ONE plan over every visible GPU (the call bench.py's
`configs` block times), with the oracle on a random sample of 1000 cells of each output: integer outputs equal,
doubles within 1e-9.  On a one-GPU box the plan has one shard; on a multi-GPU box the same test covers the
multi-device gather (tests/test_gpu_pipeline.py::test_two_devices_equal_one asserts shard equivalence bit for bit)."""
import time

import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, rel_err
from mosmicrosim_b200 import backend, dates, synthetic as syn

pytestmark = pytest.mark.gpu

SAMPLE = 1000


def run_config(n_dev, cells, years, habitats, agg, names, wet, oracle_settings=None):
    lh = syn.synthetic_lifehistory("An. gambiae")
    dm = lh.das_densmod
    densmod = dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2)
    start, end = syn.month_table(2001, years)
    dp = syn.year_dates(2001, years)
    mid_day, day_predict = dates.datenums(start, end, dp)
    pod, ps, _ = dates.period_index(dp, agg)
    monthly, altitude = syn.synthetic_forcing_large(cells, start, seed=syn.SEED + 7)
    t0 = time.perf_counter()
    with backend.ChainPlan(cells, mid_day=mid_day, day_predict=day_predict, habitats=habitats, densmod=densmod,
                           tables=lh.tables(), period_of_day=pod, n_periods=len(ps), burnin_years_pop=1,
                           mean_vars=names, want_wet_hours=wet, devices=list(range(n_dev))) as plan:
        info = plan.info()
        out = plan.run(monthly, altitude)
        kernel_s = plan.kernel_ms() * 1e-3
    t1 = time.perf_counter()
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(cells, size=SAMPLE, replace=False))
    sets = list(range(len(habitats))) if oracle_settings is None else list(oracle_settings)
    ref = oracle.run_chain(np.ascontiguousarray(monthly[:, :, pick]), mid_day, day_predict, altitude[pick],
                           [oracle.habitat(**habitats[s]) for s in sets],
                           oracle.densmod(dm.type, dm.dd_effect, dm.surface_area_cm2), lh.tables(), pod, len(ps),
                           burnin_years_pop=1, want_hourly=False)
    drift = {}
    for k, v in enumerate(out["mean_state_vars"]):
        drift[v] = rel_err(out["mean_state"][sets, k][..., pick], ref["mean"][:, ORACLE_VAR[v]])
    for k, v in enumerate(out["mean_forcing_vars"]):
        drift[v] = rel_err(out["mean_forcing"][k][..., pick], ref["mean"][0, ORACLE_VAR[v]])
    flips = int((out["wet_hours"][sets][..., pick] != ref["wet_hours"]).sum()) if wet else 0
    print(f"{cells} cells x {years} y x {len(habitats)} settings on {n_dev} GPU(s) {info}: {t1 - t0:.2f} s with plan "
          f"creation (kernels {kernel_s:.2f} s); {SAMPLE}-cell sample: wet-hour mismatches {flips}, drift {drift}")
    assert flips == 0
    assert all(d <= REL_TOL_FAST for d in drift.values()), drift


def test_config3_africa_ten_years_monthly(gpu):
    """configs[2]: ~1.5 M cells, 10 years hourly (+ burn-in year), monthly means of the six variables."""
    run_config(gpu, 1_500_000, 10, [syn.habitat_cone()], "month", list(backend.VAR_NAMES[:6]), True)


def test_config4_one_tile_five_years_daily(gpu):
    """configs[3]: one 1.25 M-cell tile of the ~10 M-cell global grid (bench.py walks 8 of them through one plan),
    5 years hourly, daily means of adult and aquatic."""
    run_config(gpu, 1_250_000, 5, [syn.habitat_cone()], "day", ["adult", "aquatic"], True)


def test_config5_ensemble_64_settings(gpu):
    """configs[4]: 250k cells x 64 water-body settings x 1 year; the oracle runs 8 of the 64 settings (every
    eighth) on the sampled cells."""
    run_config(gpu, 250_000, 1, syn.habitat_ensemble(64), "month", ["adult", "aquatic"], True,
               oracle_settings=range(3, 64, 8))
