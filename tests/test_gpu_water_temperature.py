"""The water-temperature model (an EXTENSION: the reference only sketches it, query_terraclimate.R:370-497, and stop()s
when it is requested, R/pipeline_functions.R:334-338; spec authored in DESIGN.md section 5 -- parity unpinned by
construction): the CUDA path against the oracle's restatement of the same spec, and the default path unchanged."""
import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, REL_TOL_STRICT, Case, rel_err
from mosmicrosim_b200 import backend, dates, pipeline, synthetic as syn
from test_gpu_chain import ALL_VARS, compare

pytestmark = pytest.mark.gpu


def location(case, seed=3):
    rng = np.random.default_rng(seed)
    lat = rng.uniform(-35.0, 35.0, size=case.C)
    lat[: min(3, case.C)] = (-66.0, 0.0, 71.0)[: min(3, case.C)]          # polar-circle days, the equator
    clear_sky = rng.uniform(180.0, 330.0, size=(12, case.C))
    return lat, clear_sky


def wt_options(case, **over):
    lat, clear_sky = location(case)
    w = dict(shade_proportion=0.4, albedo=0.05, emissivity=0.97, convection_w_m2_k=0.0,
             day_of_year=dates.day_of_year(case.dates_predict), month_of_day=dates.month_of_year(case.dates_predict),
             latitude=lat, clear_sky=clear_sky)
    w.update(over)
    return w


def test_default_water_temperature_is_the_air_temperature(gpu, monkeypatch):
    """model_water_temperature = FALSE: water_temperature = air_temperature (R/sim_microclimate.R:163), and asking for
    it changes no other output bit (same kernel form: the water temperature is an output of the monolithic form, so
    that form is pinned for the run without it too)."""
    monkeypatch.setenv("MMS_CHAIN_WS", "0")
    case = Case(n_cells=50, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    kw = case.gpu_kwargs(want_wet_hours=True, want_wet_flags=True)
    a = backend.run_chain(case.monthly, case.altitude, mean_vars=ALL_VARS, hourly_vars=ALL_VARS, **kw)
    b = backend.run_chain(case.monthly, case.altitude, mean_vars=ALL_VARS + ["water_temperature"],
                          hourly_vars=ALL_VARS + ["water_temperature"], **kw)
    assert b["hourly_state_vars"][-1] == "water_temperature"
    air = b["hourly_forcing"][b["hourly_forcing_vars"].index("air_temperature")]
    for s in range(2):
        assert np.array_equal(b["hourly_state"][s, -1], air)
        assert np.array_equal(b["hourly_state"][s, :-1], a["hourly_state"][s], equal_nan=True)
        assert np.array_equal(b["mean_state"][s, :-1], a["mean_state"][s], equal_nan=True)
    assert np.array_equal(a["hourly_forcing"], b["hourly_forcing"], equal_nan=True)
    assert np.array_equal(a["wet_flags"], b["wet_flags"])


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
@pytest.mark.parametrize("habitat,over", [
    (syn.habitat_cone(), dict(shade_proportion=1.0)),                                    # the notes' "simple model"
    (syn.habitat_cone(), dict(shade_proportion=0.0, albedo=0.0)),                        # full sun
    (syn.habitat_cone(max_depth=0.3, initial_volume=0.01), dict(convection_w_m2_k=15.0)),
    (syn.habitat_cylinder(radius=0.3, max_depth=0.2, leak_per_hour=0.004), dict(shade_proportion=0.7, convection_w_m2_k=8.0)),
])
def test_water_temperature_model_matches_oracle(gpu, mode, tol, habitat, over):
    case = Case(n_cells=45, habitats=[habitat], out_of_range_fraction=0.05)
    wt = wt_options(case, **over)
    dm = case.densmod
    ref = oracle.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude, [oracle.habitat(**habitat)],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables,
                           case.period_of_day, case.P, burnin_years_water=1, burnin_years_pop=1, want_hourly=True,
                           water_temperature=wt)
    names = ALL_VARS + ["water_temperature"]
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=names, hourly_vars=names, want_wet_hours=True, want_wet_flags=True, math_mode=mode,
        burnin_years_water=1, water_temperature=wt))
    worst = compare(case, out, ref, tol)
    tw = ref["hourly"][0, ORACLE_VAR["water_temperature"]]
    ta = ref["hourly"][0, ORACLE_VAR["air_temperature"]]
    assert tw.min() >= 0.0 and tw.max() <= 100.0 and np.abs(tw - ta).max() > 1.0       # the model does something
    print(mode, over, {k: v for k, v in worst.items() if "water_temperature" in k or "adult" in k})


def test_water_temperature_cox_ph_summaries_and_plan_reuse(gpu):
    """The other species, summaries only over included dates, the plan reused for a second tile."""
    included = syn.year_dates(2001, 1)[30:330]
    case = Case(n_cells=40, species="An. stephensi", included=included)
    wt = wt_options(case, shade_proportion=0.2)
    dm = case.densmod
    ref = oracle.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude, [oracle.habitat()],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables,
                           case.period_of_day, case.P, want_hourly=False, water_temperature=wt)
    names = list(backend.VAR_NAMES[:8]) + ["water_temperature"]
    with backend.ChainPlan(case.C, **case.gpu_kwargs(mean_vars=names, want_wet_hours=True, water_temperature=wt)) as plan:
        out = plan.run(case.monthly, case.altitude)
        compare(case, out, ref, REL_TOL_FAST, what=("mean",))
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        again = plan.download(plan.allocate_outputs())
        assert np.array_equal(out["mean_state"], again["mean_state"], equal_nan=True)


def test_water_temperature_errors(gpu):
    case = Case(n_cells=8, habitats=[syn.habitat_cone(), syn.habitat_cylinder()])
    wt = wt_options(case)
    with pytest.raises(backend.MmsError) as e:       # a property of the water body: one setting per call
        backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=["adult"], water_temperature=wt))
    assert e.value.code == "MMS_ERR_UNSUPPORTED"
    one = Case(n_cells=8)
    wt = wt_options(one)
    loc = {k: wt.pop(k) for k in ("latitude", "clear_sky")}
    with backend.ChainPlan(one.C, **one.gpu_kwargs(mean_vars=["adult"], water_temperature=wt)) as plan:
        with pytest.raises(backend.MmsError) as e:   # no mms_chain_set_location() yet
            plan.run(one.monthly, one.altitude)
        assert e.value.code == "MMS_ERR_INVALID_ARG"
        plan.set_location(loc["latitude"], loc["clear_sky"])
        plan.run(one.monthly, one.altitude)
    with pytest.raises(backend.MmsError):
        backend.run_chain(one.monthly, one.altitude, **one.gpu_kwargs(
            mean_vars=["adult"], water_temperature=dict(wt, shade_proportion=1.5, **loc)))


def test_tc_to_vectors_with_water_temperature(gpu):
    """The reference-facing mirror: tc_to_vectors(..., model_water_temperature = TRUE, water_shade_proportion = s) no
    longer stops; it equals the oracle run on the same monthly forcing."""
    case = Case(n_cells=30)
    tc, clear_sky, altitude = syn.synthetic_terraclimate(30, case.start, out_of_range_fraction=0.0)
    lat = np.linspace(-30.0, 30.0, 30)
    ptd = {"start": case.start, "end": case.end, "altitude": altitude, "clear_sky_SOLR": clear_sky,
           "longitude": np.arange(30.0), "latitude": lat}
    for k, name in enumerate(backend.TERRACLIMATE_VARS):
        ptd[name] = tc[k]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    with pytest.warns(UserWarning):                   # R/pipeline_functions.R:327-331
        pipeline.tc_to_vectors(ptd, lhf, dates_predict=case.dates_predict, water_shade_proportion=0.5)
    res = pipeline.tc_to_vectors(ptd, lhf, dates_predict=case.dates_predict, model_water_temperature=True,
                                 water_shade_proportion=0.5)
    month_id, n_days = dates.terraclimate_month_info(case.start, case.end)
    monthly = oracle.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
    dm = case.densmod
    wt = dict(shade_proportion=0.5, day_of_year=dates.day_of_year(case.dates_predict),
              month_of_day=dates.month_of_year(case.dates_predict), latitude=lat, clear_sky=clear_sky)
    ref = oracle.run_chain(monthly, case.mid_day, case.day_predict, altitude, [oracle.habitat()],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables,
                           case.period_of_day, case.P, want_hourly=False, water_temperature=wt)
    for name in pipeline.VECTOR_VARIABLES + ("water_temperature",):
        assert rel_err(res[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_FAST, name
    with pytest.raises(NotImplementedError):          # the stage-level verb points to the fused chain
        pipeline.simulate_hourly_conditions({}, model_water_temperature=True)
