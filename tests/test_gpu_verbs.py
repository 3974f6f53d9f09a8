"""GPU parity of the state-free pipeline verbs (daily, hourly microclimate, life history) and of the
verb-by-verb pipeline against the fused chain."""
import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, REL_TOL_STRICT, Case, rel_err
from mosmicrosim_b200 import backend, pipeline, synthetic as syn

pytestmark = pytest.mark.gpu


def oracle_daily(case):
    return np.stack([oracle.daily_from_monthly(case.monthly[:, :, c], case.mid_day, case.day_predict)
                     for c in range(case.C)], axis=2)                     # [9][D][C]


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_daily_from_monthly(gpu, mode, tol):
    for kw in (dict(n_cells=37), dict(n_cells=130, default_dates=True), dict(n_cells=5, years=2, first_year=2004)):
        case = Case(**kw)
        ref = oracle_daily(case)
        got = backend.daily_from_monthly(case.monthly, case.mid_day, case.day_predict, math_mode=mode)
        assert got.shape == ref.shape
        assert rel_err(got, ref) <= tol
    # the linear branch reproduces the knots exactly on knot days (approx1: v == x[j] -> y[j])
    case = Case(n_cells=4, default_dates=True)
    got = backend.daily_from_monthly(case.monthly, case.mid_day, case.day_predict, math_mode="strict")
    on_knot = np.isin(case.day_predict, case.mid_day)
    idx = np.searchsorted(case.mid_day, case.day_predict[on_knot])
    assert on_knot.sum() >= 8             # the 31-day months have integral mid-dates
    assert np.array_equal(got[:8][:, on_knot, :], case.monthly[:8][:, idx, :])
    with pytest.raises(backend.MmsError) as e:
        backend.daily_from_monthly(case.monthly, case.mid_day, case.day_predict + 40.0)
    assert e.value.code == "MMS_ERR_EXTRAPOLATION"


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_hourly_from_daily(gpu, mode, tol):
    for n_cells, days in ((33, None), (3, 1), (2, 2)):
        case = Case(n_cells=n_cells)
        daily = oracle_daily(case)
        if days:
            daily = np.ascontiguousarray(daily[:, :days, :])
        got = backend.hourly_from_daily(daily, math_mode=mode)
        for c in range(n_cells):
            ref = oracle.hourly_from_daily_ambient(daily[:, :, c])
            for name in ("air_temperature", "humidity", "cloudcover", "windspeed", "water_temperature", "rainfall"):
                assert rel_err(got[name][:, c], ref[name]) <= tol, (name, c, days)
    # anchors of the reference's demos: max T at 13 h, min at 6 h; humidity max at 6 h, min at 13 h
    case = Case(n_cells=8)
    daily = oracle_daily(case)
    got = backend.hourly_from_daily(daily, math_mode="strict")
    air = got["air_temperature"].reshape(case.D, 24, 8)
    np.testing.assert_allclose(air[:, 12, :], daily[1], rtol=1e-12)
    np.testing.assert_allclose(air[:, 5, :], daily[0], rtol=1e-12)
    rh = got["humidity"].reshape(case.D, 24, 8)
    np.testing.assert_allclose(rh[:, 5, :], daily[3], rtol=1e-14)
    np.testing.assert_allclose(rh[:, 12, :], daily[2], rtol=1e-14)


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_lifehistory_vectorised(gpu, mode, tol):
    rng = np.random.default_rng(7)
    n = 20_000
    tw = rng.uniform(-25, 85, n)          # beyond the knots on both sides: end-cubic extrapolation
    ta = rng.uniform(-5, 50, n)           # partly outside the ds grid -> NA
    rh = rng.uniform(-2, 103, n)
    ta[:4] = [0.0, 45.0, 45.0, 22.5]      # exactly on the grid edges
    rh[:4] = [0.0, 100.0, 50.0, 100.0]
    for species in ("An. gambiae", "An. stephensi"):
        tabs = syn.synthetic_lifehistory(species).tables()
        got = backend.lifehistory_vectorised(tw, ta, rh, tabs, math_mode=mode)
        assert rel_err(got["mdr"], oracle.spline_eval(tabs["mdr"], tw, positive=True)) <= tol
        assert rel_err(got["efd"], oracle.spline_eval(tabs["efd"], ta, positive=True)) <= tol
        assert rel_err(got["das_zerodensity"], oracle.spline_eval(tabs["das"], tw, positive=True)) <= tol
        ref_ds = oracle.interp_surface(tabs["ds"], ta, rh)
        assert np.isnan(ref_ds).sum() > 100 and not np.isnan(ref_ds[:4]).any()
        assert rel_err(got["ds"], ref_ds) <= tol


def test_verb_by_verb_pipeline_equals_fused_chain(gpu):
    """interpolate_daily_climate |> simulate_hourly_microclimate |> water |> lifehistory |> population,
    one GPU call per verb like the reference pipeline, against the fused kernel's hourly series."""
    case = Case(n_cells=40)
    pmc = {"start": case.start, "end": case.end, "altitude": case.altitude,
           "longitude": np.arange(40.0), "latitude": np.arange(40.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = case.monthly[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    daily = pipeline.interpolate_daily_climate(pmc, dates_predict=case.dates_predict, math_mode="strict")
    hourly = pipeline.simulate_hourly_microclimate(daily, math_mode="strict")
    assert hourly["hour"][:25].tolist() == list(range(1, 25)) + [1]
    water = pipeline.simulate_ephemeral_habitat_vectorised(hourly["rainfall"], hourly["air_temperature"],
                                                           hourly["humidity"], hourly["windspeed"], case.altitude,
                                                           math_mode="strict")
    cond = dict(hourly, water_surface_area=water["water_surface_area"])
    life = pipeline.simulate_hourly_lifehistory(cond, lhf, species="An. gambiae", math_mode="strict")
    pop = pipeline.simulate_population_vectorised(life["larval_habitat_area"], life["mdr"], life["efd"],
                                                  life["das_zerodensity"], life["ds"], life["das_densmod_function"],
                                                  math_mode="strict")
    fused = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        hourly_vars=["adult", "aquatic", "larval_habitat_area", "mdr", "ds", "air_temperature"], math_mode="strict"))
    st = {v: fused["hourly_state"][0, k] for k, v in enumerate(fused["hourly_state_vars"])}
    fo = {v: fused["hourly_forcing"][k] for k, v in enumerate(fused["hourly_forcing_vars"])}
    # same device functions, same order of operations: bit-identical
    assert np.array_equal(hourly["air_temperature"], fo["air_temperature"])
    assert np.array_equal(water["water_surface_area"], st["larval_habitat_area"])
    assert np.array_equal(life["mdr"], fo["mdr"]) and np.array_equal(life["ds"], fo["ds"], equal_nan=True)
    assert np.array_equal(pop["adult"], st["adult"], equal_nan=True)
    assert np.array_equal(pop["aquatic"], st["aquatic"], equal_nan=True)
    ref = case.oracle()
    assert rel_err(pop["adult"], ref["hourly"][0, ORACLE_VAR["adult"]]) <= REL_TOL_STRICT


# ---- TerraClimate variables -> monthly forcing -----------------------------

@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_terraclimate_to_monthly_matches_oracle(gpu, mode, tol):
    from mosmicrosim_b200 import dates
    start, end = syn.month_table(2001, 2)
    month_id, n_days = dates.terraclimate_month_info(start, end)
    for C in (1, 131, 1000):
        tc, clear_sky, _ = syn.synthetic_terraclimate(C, start)
        tc[0, 3, 0] = -5.0; tc[1, 3, 0] = -15.0                     # frost: the over-ice branch of VAPPRS
        ref = oracle.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
        got = backend.terraclimate_to_monthly(tc, clear_sky, month_id, n_days, math_mode=mode)
        for v, name in enumerate(pipeline.FORCING_VARS):
            assert rel_err(got[v], ref[v]) <= tol, (name, C)
        # the clamps are hit at identical cells
        assert np.array_equal(got[2] == 0.01, ref[2] == 0.01) and np.array_equal(got[7] == 0.0, ref[7] == 0.0)


def test_tc_to_vectors_equals_the_two_step_route(gpu):
    """tc_to_vectors() (raw variables up, forcing derived on the device in front of the fused chain) is bit
    identical to terraclimate_to_monthly_climate() followed by monthly_climate_to_vectors(), and matches
    the oracle's chain fed with the oracle's own monthly variables."""
    from mosmicrosim_b200 import dates
    case = Case(n_cells=40)
    tc, clear_sky, altitude = syn.synthetic_terraclimate(40, case.start, out_of_range_fraction=0.05)
    ptd = {"start": case.start, "end": case.end, "altitude": altitude, "clear_sky_SOLR": clear_sky,
           "longitude": np.arange(40.0), "latitude": -np.arange(40.0)}
    for v, name in enumerate(backend.TERRACLIMATE_VARS):
        ptd[name] = tc[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    one = pipeline.tc_to_vectors(ptd, lhf, species="An. gambiae", dates_predict=case.dates_predict)
    pmc = pipeline.terraclimate_to_monthly_climate(ptd)
    two = pipeline.monthly_climate_to_vectors(pmc, lhf, species="An. gambiae", dates_predict=case.dates_predict)
    for name in pipeline.VECTOR_VARIABLES:
        assert np.array_equal(one[name], two[name], equal_nan=True), name
    month_id, n_days = dates.terraclimate_month_info(case.start, case.end)
    monthly = oracle.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
    ref = oracle.run_chain(monthly, case.mid_day, case.day_predict, altitude, [oracle.habitat()],
                           oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"]),
                           case.tables, case.period_of_day, case.P, want_hourly=False)
    for name in pipeline.VECTOR_VARIABLES:
        assert rel_err(one[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_FAST, name
    feb = pipeline.raster_slice(one, "adult", 2001, 2)
    assert feb.shape == (1, 40) and np.array_equal(feb[0], one["adult"][1], equal_nan=True)


def test_run_chain_terraclimate_one_shot(gpu):
    """mms_run_chain_terraclimate() == terraclimate_to_monthly + mms_run_chain, bit for bit, over two shards too."""
    from mosmicrosim_b200 import dates
    case = Case(n_cells=300, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    tc, clear_sky, altitude = syn.synthetic_terraclimate(300, case.start)
    month_id, n_days = dates.terraclimate_month_info(case.start, case.end)
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:8]), want_wet_hours=True)
    monthly = backend.terraclimate_to_monthly(tc, clear_sky, month_id, n_days)
    two = backend.run_chain(monthly, altitude, **kw)
    for devices in ([0], [0, 0]):
        one = backend.run_chain_terraclimate(tc, clear_sky, month_id, n_days, altitude, **{**kw, "devices": devices})
        for k in ("mean_state", "mean_forcing", "wet_hours"):
            assert np.array_equal(one[k], two[k], equal_nan=True), (k, devices)
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain_terraclimate(tc, clear_sky, month_id * 0, n_days, altitude, **kw)
    assert e.value.code == "MMS_ERR_INVALID_ARG"


@pytest.mark.parametrize("aggregate_by", ["month", "day"])
def test_all_six_verbs_one_gpu_call_each(gpu, aggregate_by):
    """interpolate_daily_climate |> simulate_hourly_microclimate |> simulate_hourly_conditions |>
    simulate_hourly_lifehistory |> simulate_hourly_vectors |> summarise_vectors with the reference's signatures,
    one GPU call per verb; the summaries equal the fused chain's bit for bit (same device functions, same order of
    the hourly accumulation), with included_dates dropping days."""
    case = Case(n_cells=37)
    pmc = {"start": case.start, "end": case.end, "altitude": case.altitude,
           "longitude": np.arange(37.0), "latitude": np.arange(37.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = case.monthly[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    included = np.concatenate([case.dates_predict[:40], case.dates_predict[100:333]])
    kw = dict(math_mode="strict")
    vec = pipeline.simulate_hourly_vectors(
        pipeline.simulate_hourly_lifehistory(
            pipeline.simulate_hourly_conditions(
                pipeline.simulate_hourly_microclimate(
                    pipeline.interpolate_daily_climate(pmc, dates_predict=case.dates_predict, **kw), **kw), **kw),
            lhf, species="An. gambiae", **kw), **kw)
    got = pipeline.summarise_vectors(vec, aggregate_by=aggregate_by, included_dates=included)
    fused = pipeline.monthly_climate_to_vectors(pmc, lhf, species="An. gambiae", included_dates=included,
                                                aggregate_by=aggregate_by, dates_predict=case.dates_predict, **kw)
    assert np.array_equal(got["start"], fused["start"]) and np.array_equal(got["end"], fused["end"])
    assert len(got["start"]) == (273 if aggregate_by == "day" else 10)       # Jan, Feb, Apr .. Nov
    for name in pipeline.VECTOR_VARIABLES:
        assert np.array_equal(got[name], fused[name], equal_nan=True), name
    # and against the oracle's long-double two-pass mean() of the same periods
    from mosmicrosim_b200 import dates as mms_dates
    pod, _, _ = mms_dates.period_index(case.dates_predict, aggregate_by, included)
    ref = oracle.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude, [oracle.habitat()],
                           oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"]),
                           case.tables, pod, len(got["start"]), want_hourly=False)
    for name in pipeline.VECTOR_VARIABLES:
        assert rel_err(got[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_STRICT, name


def test_summarise_vectors_argument_checks(gpu):
    h = np.zeros((2, 48, 5))
    with pytest.raises(backend.MmsError) as e:
        backend.summarise_vectors(h, [1, 0], 2)                 # decreasing periods
    assert e.value.code == "MMS_ERR_INVALID_ARG"
    with pytest.raises(backend.MmsError):
        backend.summarise_vectors(h, [0, 3], 2)                 # period >= n_periods
    m = backend.summarise_vectors(h + 2.5, [-1, 1], 3)          # periods 0 and 2 have no day: NaN
    assert np.isnan(m[:, 0]).all() and np.isnan(m[:, 2]).all() and np.all(m[:, 1] == 2.5)


def test_golden_front_end_and_twins_gpu(gpu):
    """The committed golden vectors of the TerraClimate front end and of the one-pixel twins
    (tests/golden/front_end_and_twins.npz, from the numpy restatement) against the CUDA path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "front_end_and_twins.npz"))
    got = backend.terraclimate_to_monthly(g["terraclimate"], g["clear_sky"], g["month_id"], g["n_days"])
    assert rel_err(got, g["monthly"]) <= REL_TOL_FAST
    s = dict(zip(oracle.VAR_NAMES, g["series"]))
    area = pipeline.simulate_ephemeral_habitat(s, float(g["altitude"]), burnin_years=2)
    assert rel_err(area, g["water_surface_area_burnin2"]) <= REL_TOL_FAST
    pop = pipeline.simulate_population(s, syn.synthetic_lifehistory("An. gambiae").das_densmod, burnin_years=2)
    assert rel_err(pop["aquatic"], g["aquatic_burnin2"]) <= REL_TOL_FAST
    assert rel_err(pop["adult"], g["adult_burnin2"]) <= REL_TOL_FAST
