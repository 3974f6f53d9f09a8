"""GPU tests of the host pipeline (mms_chain_run: chunked, overlapped upload / kernels / download through
per-device host threads), of the multi-device plan, and of parity at the sizes and shapes the north-star
configurations use (whole 100k-cell tile, 64 settings per cell, five years of daily summaries)."""
import time

import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, Case, rel_err
from mosmicrosim_b200 import backend, dates, lifehistory as lh, pipeline, synthetic as syn
from test_gpu_chain import compare

pytestmark = pytest.mark.gpu

KEYS = ("mean_state", "mean_forcing", "wet_hours", "hourly_state", "hourly_forcing", "wet_flags")


def same_bits(a, b):
    for k in KEYS:
        assert (k in a) == (k in b), k
        if k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k


def pinned_copy(a):
    p = backend.PinnedArray(a.shape, a.dtype)
    p.array[...] = a
    return p


@pytest.mark.parametrize("chunks,workers", [("1", "1"), ("3", "1"), ("5", "3"), ("7", "4")])
def test_pipelined_run_equals_resident_path(gpu, monkeypatch, chunks, workers):
    """mms_chain_run (chunks of the shard pipelined through host worker threads and pinned bounce buffers) gives
    the bits of upload + launch + download, from pageable and from pinned host memory, for summaries, hourly series
    and flags, with a ragged number of cells."""
    case = Case(n_cells=1000 + 77, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)],
                out_of_range_fraction=0.05)
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:8]), hourly_vars=["adult", "ds"], want_wet_hours=True,
                         want_wet_flags=True)
    monkeypatch.setenv("MMS_CHAIN_CHUNKS", "1")
    monkeypatch.setenv("MMS_HOST_WORKERS", "1")
    with backend.ChainPlan(case.C, devices=[0], **kw) as plan:
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        ref = plan.download(plan.allocate_outputs())
    monkeypatch.setenv("MMS_CHAIN_CHUNKS", chunks)
    monkeypatch.setenv("MMS_HOST_WORKERS", workers)
    with backend.ChainPlan(case.C, devices=[0], **kw) as plan:
        assert plan.info() == {"shards": 1, "chunks_per_shard": int(chunks), "workers_per_shard": int(workers)}
        same_bits(ref, plan.run(case.monthly, case.altitude))                  # pageable in, pageable out
        same_bits(ref, plan.run(case.monthly, case.altitude))                  # again on the same plan
        pm, pa = pinned_copy(case.monthly), pinned_copy(case.altitude)
        out = plan.allocate_outputs(pinned=True)
        same_bits(ref, plan.run(pm.array, pa.array, out))                      # pinned in, pinned out
        same_bits(ref, plan.run(pm.array, pa.array))                           # pinned in, pageable out
        assert plan.kernel_ms() > 0
        # and the resident calls on the chunked plan
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        same_bits(ref, plan.download(plan.allocate_outputs()))


def test_pipelined_terraclimate_route(gpu, monkeypatch):
    """tc_to_vectors() whole through the pipelined call: the six raw variables in, the front end and the chain
    chunk by chunk on the device; equal to the resident route bit for bit."""
    start, end = syn.month_table(2001, 1)
    C = 900
    tc, clear_sky, altitude = syn.synthetic_terraclimate(C, start)
    mid, nd = dates.terraclimate_month_info(start, end)
    case = Case(n_cells=C)
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True)
    monkeypatch.setenv("MMS_CHAIN_CHUNKS", "1")
    with backend.ChainPlan(C, devices=[0], **kw) as plan:
        plan.upload_terraclimate(tc, clear_sky, mid, nd, altitude)
        plan.launch()
        ref = plan.download(plan.allocate_outputs())
    monkeypatch.setenv("MMS_CHAIN_CHUNKS", "4")
    monkeypatch.setenv("MMS_HOST_WORKERS", "2")
    with backend.ChainPlan(C, devices=[0], **kw) as plan:
        same_bits(ref, plan.run_terraclimate(tc, clear_sky, mid, nd, altitude))
    same_bits(ref, backend.run_chain_terraclimate(tc, clear_sky, mid, nd, altitude, **kw))


def test_nonfinite_tile_does_not_poison_the_plan(gpu):
    """A tile with NA forcing is refused (MMS_ERR_NONFINITE); the next, clean tile on the same plan is fine --
    the kernel-side flag is zeroed at every launch."""
    case = Case(n_cells=300)
    kw = case.gpu_kwargs(mean_vars=["adult", "aquatic"])
    good = backend.run_chain(case.monthly, case.altitude, **kw)
    bad = case.monthly.copy()
    bad[1, 3, 5] = np.nan
    with backend.ChainPlan(case.C, **kw) as plan:
        with pytest.raises(backend.MmsError) as e:
            plan.run(bad, case.altitude)
        assert e.value.code == "MMS_ERR_NONFINITE"
        same_bits(good, plan.run(case.monthly, case.altitude))
        plan.upload(bad, case.altitude)
        plan.launch()
        with pytest.raises(backend.MmsError):
            plan.sync()
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        same_bits(good, plan.download(plan.allocate_outputs()))


def test_two_devices_equal_one(gpu):
    """The plan's own multi-device path on real devices: devices = [0, 1] == devices = [0] bit for bit
    (skipped on a one-GPU box; the shard logic itself also runs as devices = [0, 0, 0] in test_gpu_stage.py)."""
    if gpu < 2:
        pytest.skip("needs two GPUs")
    case = Case(n_cells=5000 + 13, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:8]), want_wet_hours=True)
    one = backend.run_chain(case.monthly, case.altitude, devices=[0], **kw)
    devs = list(range(gpu))
    for d in ([0, 1], devs):
        with backend.ChainPlan(case.C, devices=d, **kw) as plan:
            assert plan.info()["shards"] == len(d)
            same_bits(one, plan.run(case.monthly, case.altitude))
            plan.upload(case.monthly, case.altitude)
            plan.launch()
            same_bits(one, plan.download(plan.allocate_outputs()))


def test_whole_tile_flip_count_and_drift(gpu):
    """BASELINE configs[1], ALL 100 000 cells against the oracle (not a sample): the integer outputs (wet-hour
    counts per month: 1.75e9 wet/dry decisions behind them) must be equal everywhere -- the flip count is
    reported separately from the floating-point drift, as SURVEY section 7 asks."""
    habs = [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)]
    case = Case(n_cells=100_000, habitats=habs, out_of_range_fraction=0.01)
    names = list(backend.VAR_NAMES[:6])
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=names, want_wet_hours=True))
    t0 = time.perf_counter()
    ref = case.oracle(want_hourly=False)
    t1 = time.perf_counter()
    flips = int((out["wet_hours"] != ref["wet_hours"]).sum())
    flipped_cells = int((out["wet_hours"] != ref["wet_hours"]).any(axis=(0, 1)).sum())
    drift = {}
    for k, v in enumerate(out["mean_state_vars"]):
        drift[v] = rel_err(out["mean_state"][:, k], ref["mean"][:, ORACLE_VAR[v]])
    for k, v in enumerate(out["mean_forcing_vars"]):
        drift[v] = rel_err(out["mean_forcing"][k], ref["mean"][0, ORACLE_VAR[v]])
    print(f"whole tile: {case.C} cells x {len(habs)} habitats x {case.P} months; wet-hour count mismatches {flips} "
          f"in {flipped_cells} cells; max relative drift {drift}; oracle {t1 - t0:.1f} s")
    assert flips == 0
    assert all(d <= REL_TOL_FAST for d in drift.values()), drift


def test_sixty_four_settings(gpu):
    """BASELINE configs[4] in miniature: the 64-setting habitat ensemble (depth x inflow x leak x shape), every
    setting of every cell against the oracle."""
    habs = syn.habitat_ensemble(64)
    case = Case(n_cells=70, habitats=habs, out_of_range_fraction=0.05)
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=["adult", "aquatic", "larval_habitat_area", "water_volume", "mdr", "ds"], want_wet_hours=True))
    assert out["mean_state"].shape[0] == 64
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))
    out2 = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=["adult", "aquatic"],
                                                                            want_wet_hours=True))
    compare(case, out2, ref, REL_TOL_FAST, what=("mean",))


def test_five_years_daily_summaries(gpu):
    """BASELINE configs[3] in miniature: five years (one leap day), daily means of adult and aquatic only."""
    case = Case(n_cells=48, years=5, first_year=2003, aggregate_by="day")
    assert case.P == case.D == 1826
    ref = case.oracle(want_hourly=False)
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=["adult", "aquatic"],
                                                                           want_wet_hours=True))
    compare(case, out, ref, REL_TOL_FAST, what=("mean",))


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("species", ["An. stephensi", "An. gambiae"])
def test_das_spline_at_and_above_one(gpu, mode, species):
    """positive_spline() clamps at 0 only: a das_temp table that reaches 1 exactly and overshoots it.  The reference
    then takes log(-log p) of p = 1 (-Inf: das_t = 1) or of p > 1 (NaN: that hour's states are NA and are reset at
    the start of the next step, R/sim_pop_dynamics.R:33-34) -- on a habitat that dries out (density up to a / eps)."""
    case = Case(n_cells=40, species=species, out_of_range_fraction=0.0,
                habitats=[syn.habitat_cylinder(radius=0.2, max_depth=0.1, leak_per_hour=0.05)])
    t = case.tables["das"]["x"]
    das = np.clip(1.04 * np.exp(-((t - 26.0) / 9.0) ** 2), 0.0, None)
    das[(t > 21.95) & (t < 22.55)] = 1.0                   # a plateau of exact ones
    tab = lh.positive_spline(das, t).table
    assert tab.y.max() > 1.0 and (tab.y == 1.0).any()
    case.tables = dict(case.tables, das=tab.as_dict())
    case.monthly[0] = 21.0 + 0.0 * case.monthly[0]          # tmin 21, tmax 29: every hour near the plateau / overshoot
    case.monthly[1] = 29.0 + 0.0 * case.monthly[1]
    case.monthly[0, :, :8] = 22.0; case.monthly[1, :, :8] = 22.0    # eight cells sit on a knot of the plateau: das0 == 1 exactly
    ref = case.oracle()
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=["adult", "aquatic", "das_zerodensity"], hourly_vars=["adult", "aquatic", "das_zerodensity"],
        want_wet_flags=True, math_mode=mode))
    das_h = ref["hourly"][0, ORACLE_VAR["das_zerodensity"]]
    assert (das_h > 1.0).any() and (das_h == 1.0).any()
    assert np.isnan(ref["hourly"][0, ORACLE_VAR["aquatic"]]).any()
    compare(case, out, ref, 1e-11 if mode == "strict" else REL_TOL_FAST)


@pytest.mark.parametrize("devices,chunks,workers", [([0], "1", "1"), ([0], "4", "3"), ([0, 0, 0], "2", "2")])
def test_outputs_scattered_into_raster_order(gpu, monkeypatch, devices, chunks, workers):
    """N2: with raster_cell_idx the rows come back as whole raster slices -- pixel c in column cell_idx[c], NaN (0 in
    the integer outputs) where the raster has no pixel of the tile -- exactly what the numpy scatter
    raster[:, cell_idx] = values of the plain run gives (create_vector_raster_slice(), R/pipeline_functions.R:1316-1322);
    summaries, hourly series and flags, pageable and pinned buffers, several chunks / workers / shards."""
    monkeypatch.setenv("MMS_CHAIN_CHUNKS", chunks)
    monkeypatch.setenv("MMS_HOST_WORKERS", workers)
    case = Case(n_cells=700 + 19, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    rng = np.random.default_rng(9)
    R = 2600
    cell_idx = np.sort(rng.choice(np.arange(5, R - 40), size=case.C, replace=False))     # gaps in front, inside, behind
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:8]), hourly_vars=["aquatic", "mdr"], want_wet_hours=True,
                         want_wet_flags=True, devices=devices)
    plain = backend.run_chain(case.monthly, case.altitude, **kw)
    with backend.ChainPlan(case.C, raster_cell_idx=cell_idx, n_raster_cells=R, **kw) as plan:
        got = plan.run(case.monthly, case.altitude)
        pinned = plan.run(case.monthly, case.altitude, plan.allocate_outputs(pinned=True))
        plan.upload(case.monthly, case.altitude)
        plan.launch()
        resident = plan.download(plan.allocate_outputs())
    for k in KEYS:
        fill = np.nan if plain[k].dtype == np.float64 else 0
        want = np.full(plain[k].shape[:-1] + (R,), fill, dtype=plain[k].dtype)
        want[..., cell_idx] = plain[k]
        for out in (got, pinned, resident):
            assert out[k].shape == want.shape
            assert np.array_equal(out[k], want, equal_nan=True), k
    with pytest.raises(backend.MmsError) as e:          # not in raster order
        backend.ChainPlan(case.C, raster_cell_idx=cell_idx[::-1].copy(), n_raster_cells=R, **kw)
    assert e.value.code == "MMS_ERR_UNSUPPORTED"


@pytest.mark.parametrize("n_settings", [8, 12, 64])
@pytest.mark.parametrize("fan", ["0", "1"])
def test_ensemble_forms(gpu, monkeypatch, n_settings, fan):
    """Habitat ensembles: the fan-out form (one forcing warp feeding the state warps of all settings of its cells,
    chosen automatically for 8, 12, ... settings) and the form it replaces (every group of four settings recomputes the
    forcing), both against the oracle: hourly series, flags, summaries over included dates, NA cells, ragged cells."""
    monkeypatch.setenv("MMS_CHAIN_FAN", fan)
    included = syn.year_dates(2001, 1)[15:250]
    case = Case(n_cells=45, habitats=syn.habitat_ensemble(64)[5:5 + n_settings] if n_settings < 64 else syn.habitat_ensemble(64),
                included=included, out_of_range_fraction=0.1)
    ref = case.oracle()
    names = ["adult", "aquatic", "mdr", "ds", "larval_habitat_area", "water_volume", "air_temperature"]
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(
        mean_vars=names, hourly_vars=["adult", "larval_habitat_area", "efd"], want_wet_hours=True, want_wet_flags=True))
    compare(case, out, ref, REL_TOL_FAST)
    for species in ("An. gambiae", "An. stephensi"):          # summaries only: density dependence compiled in
        c2 = Case(n_cells=45, species=species, habitats=case.habitats, out_of_range_fraction=0.1)
        out2 = backend.run_chain(c2.monthly, c2.altitude, **c2.gpu_kwargs(mean_vars=["adult", "aquatic"], want_wet_hours=True,
                                                                          burnin_years_water=1))
        compare(c2, out2, c2.oracle(want_hourly=False, burnin_years_water=1), REL_TOL_FAST, what=("mean",))


def test_april_start_series_errors_by_default_and_runs_clipped(gpu):
    """The default dates of a series that starts in a 30-day month overhang the knots by half a day (the reference's
    GAM branch): MMS_ERR_EXTRAPOLATION by default, the clipped run equals the oracle on the clipped dates."""
    months = np.datetime64("2001-04", "M") + np.arange(23)                 # April 2001 .. February 2003
    start = months.astype("datetime64[D]")
    end = (months + 1).astype("datetime64[D]") - np.timedelta64(1, "D")
    C = 25
    monthly, altitude = syn.synthetic_forcing(C, start, out_of_range_fraction=0.0)
    case = Case(n_cells=C)
    pmc = {"start": start, "end": end, "altitude": altitude, "longitude": np.arange(C * 1.0), "latitude": np.arange(C * 1.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = monthly[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    with pytest.raises(backend.MmsError) as e:
        pipeline.monthly_climate_to_vectors(pmc, lhf)
    assert e.value.code == "MMS_ERR_EXTRAPOLATION"
    with pytest.warns(UserWarning, match="dropped"):
        res = pipeline.monthly_climate_to_vectors(pmc, lhf, on_extrapolation="clip")
    dp, dropped = dates.clip_dates_predict(start, end, dates.default_dates_predict(start, end))
    assert dropped == 2
    mid_day, day_predict = dates.datenums(start, end, dp)
    pod, ps, _ = dates.period_index(dp, "month")
    dm = case.densmod
    ref = oracle.run_chain(monthly, mid_day, day_predict, altitude, [oracle.habitat()],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables, pod, len(ps),
                           want_hourly=False)
    for name in pipeline.VECTOR_VARIABLES:
        assert rel_err(res[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_FAST, name


def test_na_knots_refused_by_default_and_interpolated_on_request(gpu):
    """NA in the monthly forcing: MMS_ERR_NONFINITE by default; with na_knots="interpolate" the knots are filled on the
    host the way stats::approx() -- which drops them -- interpolates there, and the run equals the oracle (whose
    restatement of approx() drops them) on the tile WITH the holes."""
    start, end = syn.month_table(2001, 1)
    C = 25
    monthly, altitude = syn.synthetic_forcing(C, start, out_of_range_fraction=0.0)
    rng = np.random.default_rng(9)
    holes = monthly.copy()
    for c in range(0, C, 3):
        v = int(rng.integers(0, 9))
        m0 = int(rng.integers(2, holes.shape[1] - 4))
        holes[v, m0:m0 + int(rng.integers(1, 3)), c] = np.nan
    case = Case(n_cells=C)
    pmc = {"start": start, "end": end, "altitude": altitude, "longitude": np.arange(C * 1.0), "latitude": np.arange(C * 1.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = holes[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    with pytest.raises(backend.MmsError) as e:
        pipeline.monthly_climate_to_vectors(pmc, lhf)
    assert e.value.code == "MMS_ERR_NONFINITE"
    res = pipeline.monthly_climate_to_vectors(pmc, lhf, na_knots="interpolate")
    dp = dates.default_dates_predict(start, end)
    mid_day, day_predict = dates.datenums(start, end, dp)
    pod, ps, _ = dates.period_index(dp, "month")
    dm = case.densmod
    ref = oracle.run_chain(holes, mid_day, day_predict, altitude, [oracle.habitat()],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables, pod, len(ps),
                           want_hourly=False)
    for name in pipeline.VECTOR_VARIABLES:
        assert rel_err(res[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_FAST, name
