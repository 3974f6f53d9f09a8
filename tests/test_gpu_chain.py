"""GPU parity: the fused sm_100a chain (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from helpers import ABS_FLOOR, ORACLE_VAR, REL_TOL_FAST, REL_TOL_STRICT, Case, build_c_driver, rel_err, write_c_driver_inputs
from mosmicrosim_b200 import backend, synthetic as syn

pytestmark = pytest.mark.gpu

# every variable the oracle carries too (cloud cover is device-only; the water temperature -- the air temperature
# unless it is modelled -- lives in the monolithic kernel form and has its own tests, test_gpu_water_temperature.py)
ALL_VARS = [v for v in backend.VAR_NAMES if v not in ("cloudcover", "water_temperature")]


def compare(case, out, ref, tol, what=("mean", "hourly")):
    worst = {}
    flips = 0
    if "hourly" in what and "hourly_state" in out:
        for k, v in enumerate(out["hourly_state_vars"]):
            e = rel_err(out["hourly_state"][:, k], ref["hourly"][:, ORACLE_VAR[v]])
            worst["hourly " + v] = e
        for k, v in enumerate(out["hourly_forcing_vars"]):
            if v == "cloudcover":
                continue
            e = rel_err(out["hourly_forcing"][k], ref["hourly"][0, ORACLE_VAR[v]])
            worst["hourly " + v] = e
        if "wet_flags" in out:
            flips = int((out["wet_flags"] != ref["wet"]).sum())
    if "mean" in what and "mean_state" in out:
        for k, v in enumerate(out["mean_state_vars"]):
            worst["mean " + v] = rel_err(out["mean_state"][:, k], ref["mean"][:, ORACLE_VAR[v]])
        for k, v in enumerate(out["mean_forcing_vars"]):
            if v == "cloudcover":
                continue
            worst["mean " + v] = rel_err(out["mean_forcing"][k], ref["mean"][0, ORACLE_VAR[v]])
        if "wet_hours" in out:
            flips += int((out["wet_hours"] != ref["wet_hours"]).sum())
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, f"parity beyond {tol}: {bad} (all: {worst})"
    assert flips == 0, f"{flips} wet/dry flag or wet-hour-count mismatches"
    return worst


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_chain_hourly_and_monthly_cone(gpu, mode, tol):
    """cfg1-like: cone/no-leak (the reference's water body), 1 year, full series + monthly means."""
    case = Case(n_cells=70)
    ref = case.oracle()
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, hourly_vars=ALL_VARS, want_wet_hours=True, want_wet_flags=True, math_mode=mode))
    worst = compare(case, out, ref, tol)
    print(mode, worst)


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_chain_two_habitats_pipeline_vars(gpu, mode, tol):
    """cfg2-like: cone + cylinder with leakage, the six reference variables, monthly means only."""
    case = Case(n_cells=300, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode=mode))
    compare(case, out, ref, tol, what=("mean",))


def test_chain_cox_ph_species(gpu):
    case = Case(n_cells=64, species="An. stephensi")
    ref = case.oracle()
    for mode, tol in (("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)):
        out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
            mean_vars=ALL_VARS, hourly_vars=["adult", "aquatic"], want_wet_flags=True, math_mode=mode))
        compare(case, out, ref, tol)


def test_chain_four_settings_and_daily_periods(gpu):
    habs = syn.habitat_ensemble(4)
    case = Case(n_cells=40, habitats=habs, aggregate_by="day")
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, want_wet_hours=True, math_mode="fast"))
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))


def test_chain_default_dates_included_dates_and_water_burnin(gpu):
    """dates_predict = mid-first-month .. mid-last-month (the reference default), summaries only
    over included_dates, water burn-in year on."""
    case0 = Case(n_cells=33, default_dates=True)
    included = syn.year_dates(2001, 1)[40:300]
    case = Case(n_cells=33, default_dates=True, included=included)
    assert case.D == case0.D and case.P < case0.P
    ref = case.oracle(want_hourly=False, burnin_years_water=1, burnin_years_pop=2)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, want_wet_hours=True, math_mode="strict", burnin_years_water=1, burnin_years_pop=2))
    compare(case, out, ref, REL_TOL_STRICT, what=("mean",))


def test_chain_three_settings_odd(gpu):
    """S = 3 exercises the one-setting-per-thread kernel with grid.y = 3."""
    case = Case(n_cells=20, habitats=syn.habitat_ensemble(3))
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:8]), math_mode="fast"))
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))


@pytest.mark.parametrize("form", ["0", "1"])
@pytest.mark.parametrize("n_settings", [1, 2, 4])
def test_both_kernel_forms(gpu, monkeypatch, form, n_settings):
    """fast mode has two kernel forms (monolithic, and forcing/state warp pairs); the library picks
    one by problem shape.  Force each and check it: hourly series, means, flags, NA cells, a ragged
    number of cells, and summaries restricted to included dates."""
    monkeypatch.setenv("MMS_CHAIN_WS", form)
    included = syn.year_dates(2001, 1)[20:200]
    case = Case(n_cells=77, habitats=syn.habitat_ensemble(8)[2:2 + n_settings], included=included,
                out_of_range_fraction=0.2)
    ref = case.oracle()
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, hourly_vars=ALL_VARS, want_wet_hours=True, want_wet_flags=True, math_mode="fast"))
    compare(case, out, ref, REL_TOL_FAST)
    # the pipeline-variables-only instantiation (kFull = false), water burn-in on
    ref2 = case.oracle(want_hourly=False, burnin_years_water=1)
    out2 = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode="fast", burnin_years_water=1))
    compare(case, out2, ref2, REL_TOL_FAST, what=("mean",))


@pytest.mark.parametrize("n_settings", [1, 2, 6])
def test_team_form(gpu, monkeypatch, n_settings):
    """The team form (MMS_CHAIN_TEAM=1: climate / rates / state warps, chain_team.cuh; an experiment kept behind its
    switch, profiles/r02_pair_kernel_binder.md): against the oracle like the other forms, and -- same device functions on
    the same values -- bit for bit what the pair form gives.  The pair form's state warp runs the software-pipelined
    hour loop (state_pass_pipe), the team form's state warps the plain one: the comparison covers that too."""
    included = syn.year_dates(2001, 1)[20:200]
    case = Case(n_cells=77, habitats=syn.habitat_ensemble(8)[2:2 + n_settings], included=included,
                out_of_range_fraction=0.2)
    kw = case.gpu_kwargs(mean_vars=ALL_VARS, hourly_vars=ALL_VARS, want_wet_hours=True, want_wet_flags=True,
                         math_mode="fast")
    kw2 = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode="fast",
                          burnin_years_water=1)
    monkeypatch.setenv("MMS_CHAIN_WS", "1")
    pair, pair2 = backend.run_chain(case.monthly, case.altitude, **kw), backend.run_chain(case.monthly, case.altitude, **kw2)
    monkeypatch.setenv("MMS_CHAIN_TEAM", "1")
    team, team2 = backend.run_chain(case.monthly, case.altitude, **kw), backend.run_chain(case.monthly, case.altitude, **kw2)
    compare(case, team, case.oracle(), REL_TOL_FAST)
    compare(case, team2, case.oracle(want_hourly=False, burnin_years_water=1), REL_TOL_FAST, what=("mean",))
    for a, b in ((pair, team), (pair2, team2)):
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_chain_two_years_with_leap_day(gpu, mode, tol):
    """2004-2005: 731 days, M = 26 knots; the burn-in still repeats the first 8760 steps
    (365 * 24 is hard-coded in the reference, R/sim_pop_dynamics.R:222), annual periods."""
    case = Case(n_cells=24, years=2, first_year=2004, aggregate_by="year")
    assert case.D == 731 and case.P == 2
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, want_wet_hours=True, math_mode=mode))
    compare(case, out, ref, tol, what=("mean",))


def test_plan_matches_one_shot_and_is_repeatable(gpu):
    case = Case(n_cells=200)
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True)
    one = backend.run_chain(case.monthly, case.altitude, **kw)
    with backend.ChainPlan(case.C, **kw) as plan:
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        a = plan.download(plan.allocate_outputs())
        plan.launch()                       # relaunch on resident inputs: identical bits
        b = plan.download(plan.allocate_outputs())
        assert plan.kernel_ms() > 0
        assert plan.launch_count() >= 3
    for k in ("mean_state", "mean_forcing", "wet_hours"):
        assert np.array_equal(one[k], a[k], equal_nan=True)
        assert np.array_equal(a[k], b[k], equal_nan=True)


def test_errors(gpu):
    case = Case(n_cells=8)
    kw = case.gpu_kwargs(mean_vars=["adult"])
    # extrapolation: a prediction day before the first knot
    bad = dict(kw, day_predict=case.day_predict - 400.0)
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain(case.monthly, case.altitude, **bad)
    assert e.value.code == "MMS_ERR_EXTRAPOLATION"
    # burn-in longer than the series
    short = Case(n_cells=8)
    kw2 = dict(short.gpu_kwargs(mean_vars=["adult"]))
    kw2["day_predict"] = short.day_predict[:100]
    kw2["period_of_day"] = short.period_of_day[:100]
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain(short.monthly, short.altitude, **kw2)
    assert e.value.code == "MMS_ERR_BURNIN"
    # NA in the forcing
    m = case.monthly.copy()
    m[1, 3, 5] = np.nan
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain(m, case.altitude, **kw)
    assert e.value.code == "MMS_ERR_NONFINITE"


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
@pytest.mark.parametrize("species", ["An. gambiae", "An. stephensi"])
def test_chain_extreme_cells(gpu, mode, tol, species):
    """Cells that live on the guards of the branch-free math: never any rain (volume 0, log(0), 0/eps density),
    a deluge (volume pinned at its maximum), frost (all life-history rates 0), heat beyond the ds grid (NA every
    hour), saturated air (evaporation 0), a gale, and a population driven to extinction / to huge densities."""
    case = Case(n_cells=8, species=species, out_of_range_fraction=0.0,
                habitats=[syn.habitat_cone(), syn.habitat_cylinder(radius=0.2, max_depth=0.1, leak_per_hour=0.05)])
    m = case.monthly
    F = {n: i for i, n in enumerate(("tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax", "rain"))}
    m[F["rain"], :, 0] = 0.0                                         # never wet
    m[F["rain"], :, 1] = 80.0                                        # always full
    m[F["tmin"], :, 2] = -12.0; m[F["tmax"], :, 2] = -1.0            # frost
    m[F["tmin"], :, 3] = 38.0; m[F["tmax"], :, 3] = 52.0             # beyond the 0-45 C grid for part of every day
    m[F["rhmin"], :, 4] = 100.0; m[F["rhmax"], :, 4] = 100.0         # saturated air
    m[F["wsmin"], :, 5] = 4.0; m[F["wsmax"], :, 5] = 40.0            # gale: fast evaporation
    m[F["rain"], ::2, 6] = 0.0; m[F["rain"], 1::2, 6] = 30.0         # alternating drought / flood months
    m[F["tmin"], :, 7] = 44.0; m[F["tmax"], :, 7] = 45.0; m[F["rain"], :, 7] = 0.05   # hot trickle: tiny pools, huge densities
    ref = case.oracle()
    out = backend.run_chain(m, case.altitude, **case.gpu_kwargs(
        mean_vars=ALL_VARS, hourly_vars=ALL_VARS, want_wet_hours=True, want_wet_flags=True, math_mode=mode))
    compare(case, out, ref, tol)
    wet = ref["wet"]
    assert wet[0][:, 0].sum() == 0 and wet[0][:, 1].mean() > 0.99       # the cases are what they claim to be
    assert np.isnan(ref["hourly"][0, ORACLE_VAR["ds"]][:, 3]).any()


def test_chain_single_cell(gpu):
    """BASELINE configs[0]: one cell, one year, every output, both math modes and both species."""
    for species in ("An. gambiae", "An. stephensi"):
        case = Case(n_cells=1, species=species, out_of_range_fraction=0.0)
        ref = case.oracle()
        for mode, tol in (("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)):
            out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
                mean_vars=ALL_VARS, hourly_vars=ALL_VARS, want_wet_hours=True, want_wet_flags=True, math_mode=mode))
            compare(case, out, ref, tol)


def test_chain_ten_years_monthly(gpu):
    """BASELINE configs[2] in miniature: ten years (two leap days, 87 648 hourly steps + burn-in), 120 monthly means."""
    case = Case(n_cells=40, years=10, first_year=1995,
                habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:8]), want_wet_hours=True))
    assert out["mean_state"].shape[2] == 120
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))


@pytest.mark.parametrize("species", ["An. gambiae", "An. stephensi"])
def test_chain_pair_form_with_compiled_in_density_dependence(gpu, species):
    """Two habitats, summaries of the six reference variables only: the pair kernel instantiations that have the
    density-dependence form (logit / cox_ph) compiled in."""
    case = Case(n_cells=200, species=species, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True))
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))


@pytest.mark.parametrize("species", ["An. gambiae", "An. stephensi"])
@pytest.mark.parametrize("hab", [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.001)])
def test_chain_one_setting_specialised_instantiations(gpu, species, hab):
    """One habitat per cell, summaries only: the monolithic instantiations with the density-dependence form and
    the water-body shape compiled in (all four combinations), and three same-shape settings (shape compiled in,
    one setting per thread)."""
    for habitats in ([hab], [hab, dict(hab, inflow_multiplier=2.0), dict(hab, max_depth=0.3)]):
        case = Case(n_cells=100, species=species, habitats=habitats)
        ref = case.oracle(want_hourly=False)
        out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
            mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True))
        compare(case, out, ref, REL_TOL_FAST, what=("mean",))


def test_plain_c_caller(gpu, tmp_path):
    """tests/c_driver/abi_driver.c -- a C99 program that fills the structs of include/mmsgpu.h by hand and calls
    mms_run_chain like the .Call shim does -- gives bit-identical results to the ctypes route (and an invalid
    call comes back as a status code with a message)."""
    import subprocess
    exe = build_c_driver(tmp_path)
    case = Case(n_cells=150, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    names = list(backend.VAR_NAMES[:6])
    ref = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=names, want_wet_hours=True))
    d = tmp_path
    write_c_driver_inputs(case, names, d, ref["mean_state"].shape[1], ref["mean_forcing"].shape[0])
    r = subprocess.run([exe, str(d)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout, r.stderr)
    got_state = np.fromfile(str(d / "mean_state.f64")).reshape(ref["mean_state"].shape)
    got_forc = np.fromfile(str(d / "mean_forcing.f64")).reshape(ref["mean_forcing"].shape)
    got_wet = np.fromfile(str(d / "wet_hours.i32"), dtype=np.int32).reshape(ref["wet_hours"].shape)
    assert np.array_equal(got_state, ref["mean_state"], equal_nan=True)
    assert np.array_equal(got_forc, ref["mean_forcing"], equal_nan=True)
    assert np.array_equal(got_wet, ref["wet_hours"])


@pytest.mark.parametrize("shapes", ["cone,cone", "cylinder,cone", "cylinder,cylinder"])
def test_chain_pair_form_shape_patterns(gpu, shapes):
    """The pair-kernel instantiations with both water-body shapes compiled in ((cone, cylinder) is covered above)."""
    mk = {"cone": lambda k: syn.habitat_cone(max_depth=1.0 - 0.3 * k),
          "cylinder": lambda k: syn.habitat_cylinder(leak_per_hour=0.001 * (k + 1))}
    pair = [mk[name](k) for k, name in enumerate(shapes.split(","))]
    case = Case(n_cells=150, habitats=pair)
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True))
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))
