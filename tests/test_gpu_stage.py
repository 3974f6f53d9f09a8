"""GPU parity of the stage-level twins (R matrices in / out), of the multi-shard
path, and size-independent properties at the benchmark's full size."""
import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, REL_TOL_STRICT, Case, rel_err
from mosmicrosim_b200 import backend, lifehistory as lh, pipeline, synthetic as syn

pytestmark = pytest.mark.gpu


def hourly_inputs(n_cells=50, years=1, seed=11):
    """Hourly microclimate / life-history matrices [H, C] from the oracle's own chain."""
    case = Case(n_cells=n_cells, years=years, seed=seed)
    ref = case.oracle()
    h = ref["hourly"][0]
    return case, {name: np.asfortranarray(h[ORACLE_VAR[name]]) for name in ORACLE_VAR}


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
@pytest.mark.parametrize("hab,burnin", [(syn.habitat_cone(), 0), (syn.habitat_cone(max_depth=0.5, inflow_multiplier=2.0, initial_volume=0.2), 1),
                                        (syn.habitat_cylinder(leak_per_hour=0.003), 2)])
def test_water_twin(gpu, mode, tol, hab, burnin):
    case, m = hourly_inputs(n_cells=131)
    ref = oracle.simulate_ephemeral_habitat_vectorised(m["rainfall"], m["air_temperature"], m["humidity"],
                                                       m["windspeed"], case.altitude, oracle.habitat(**hab), burnin)
    got = backend.water_vectorised(m["rainfall"], m["air_temperature"], m["humidity"], m["windspeed"],
                                   case.altitude, hab, burnin_years=burnin, math_mode=mode, want_volume=True,
                                   want_wet=True)
    assert rel_err(got["water_surface_area"], ref["water_surface_area"]) <= tol
    assert rel_err(got["water_volume"], ref["water_volume"]) <= tol
    assert np.array_equal(got["wet"], ref["wet"])                   # wet/dry flags bit exact


RAGGED = ((8759, 3), (101, 1), (8, 2), (7, 3), (10, 33), (46, 40), (3, 65), (1, 1), (12, 32))


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_water_twin_ragged_shapes(gpu, mode, tol):
    """H odd (8-byte copies), H = 2 mod 4 (16-byte copies, no 32-byte stores), H not a multiple of the
    time tile, C = 1 and C not a multiple of the 32-pixel unit."""
    case, m = hourly_inputs(n_cells=65)
    for H, C in RAGGED:
        sub = {k: np.asfortranarray(v[:H, :C]) for k, v in m.items()}
        ref = oracle.simulate_ephemeral_habitat_vectorised(sub["rainfall"], sub["air_temperature"], sub["humidity"],
                                                           sub["windspeed"], case.altitude[:C], oracle.habitat(), 0)
        got = backend.water_vectorised(sub["rainfall"], sub["air_temperature"], sub["humidity"], sub["windspeed"],
                                       case.altitude[:C], syn.habitat_cone(), math_mode=mode, want_volume=True,
                                       want_wet=True)
        assert rel_err(got["water_surface_area"], ref["water_surface_area"]) <= tol, (H, C)
        assert rel_err(got["water_volume"], ref["water_volume"]) <= tol, (H, C)
        assert np.array_equal(got["wet"], ref["wet"]), (H, C)


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_population_twin_ragged_shapes(gpu, mode, tol):
    case, m = hourly_inputs(n_cells=65)
    dmh = syn.synthetic_lifehistory("An. gambiae").das_densmod
    dm = oracle.densmod(dmh.type, dmh.dd_effect, dmh.surface_area_cm2)
    names = ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")
    for H, C in RAGGED:
        sub = [np.asfortranarray(m[k][:H, :C]) for k in names]
        ref = oracle.simulate_population_vectorised(*sub, dm, burnin_years=0)
        got = pipeline.simulate_population_vectorised(*sub, dmh, burnin_years=0, math_mode=mode)
        assert rel_err(got["aquatic"], ref["aquatic"]) <= tol, (H, C)
        assert rel_err(got["adult"], ref["adult"]) <= tol, (H, C)


def test_twins_several_rounds_of_units(gpu):
    """More 32-pixel units than one round of persistent warps holds (150k pixels, short series)."""
    rng = np.random.default_rng(5)
    H, C = 52, 150_001
    mk = lambda lo, hi: np.asfortranarray(rng.uniform(lo, hi, size=(H, C)))
    rain, airt, rh, wind, alt = mk(0, 0.4), mk(12, 36), mk(25, 100), mk(0.1, 5), rng.uniform(0, 2500, C)
    hab = syn.habitat_cone()
    ref = oracle.simulate_ephemeral_habitat_vectorised(rain, airt, rh, wind, alt, oracle.habitat(**hab), 0)
    got = backend.water_vectorised(rain, airt, rh, wind, alt, hab, want_volume=True, want_wet=True)
    assert rel_err(got["water_surface_area"], ref["water_surface_area"]) <= REL_TOL_FAST
    assert rel_err(got["water_volume"], ref["water_volume"]) <= REL_TOL_FAST
    assert np.array_equal(got["wet"], ref["wet"])
    area = ref["water_surface_area"]
    mdr, efd, das, ds = mk(0, 0.16), mk(0, 10), mk(0.2, 0.95), mk(0.3, 0.9)
    dmh = syn.synthetic_lifehistory("An. gambiae").das_densmod
    refp = oracle.simulate_population_vectorised(area, mdr, efd, das, ds,
                                                 oracle.densmod(dmh.type, dmh.dd_effect, dmh.surface_area_cm2),
                                                 burnin_years=0)
    gotp = pipeline.simulate_population_vectorised(area, mdr, efd, das, ds, dmh, burnin_years=0)
    assert rel_err(gotp["aquatic"], refp["aquatic"]) <= REL_TOL_FAST
    assert rel_err(gotp["adult"], refp["adult"]) <= REL_TOL_FAST


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
@pytest.mark.parametrize("species,burnin", [("An. gambiae", 1), ("An. stephensi", 0), ("An. stephensi", 2)])
def test_population_twin(gpu, mode, tol, species, burnin):
    case, m = hourly_inputs(n_cells=77)
    dmh = syn.synthetic_lifehistory(species).das_densmod
    dm = oracle.densmod(dmh.type, dmh.dd_effect, dmh.surface_area_cm2)
    ref = oracle.simulate_population_vectorised(m["larval_habitat_area"], m["mdr"], m["efd"], m["das_zerodensity"],
                                                m["ds"], dm, burnin_years=burnin)
    got = pipeline.simulate_population_vectorised(m["larval_habitat_area"], m["mdr"], m["efd"], m["das_zerodensity"],
                                                  m["ds"], dmh, burnin_years=burnin, math_mode=mode)
    assert rel_err(got["aquatic"], ref["aquatic"]) <= tol
    assert rel_err(got["adult"], ref["adult"]) <= tol


def test_reference_signature_mirror_defaults(gpu):
    """simulate_ephemeral_habitat_vectorised() with the reference's defaults (cone, depth 1, no burn-in)."""
    case, m = hourly_inputs(n_cells=20)
    got = pipeline.simulate_ephemeral_habitat_vectorised(m["rainfall"], m["air_temperature"], m["humidity"],
                                                         m["windspeed"], case.altitude)
    assert list(got) == ["water_surface_area"] and got["water_surface_area"].shape == (case.H, 20)
    assert rel_err(got["water_surface_area"], m["larval_habitat_area"]) <= REL_TOL_FAST
    with pytest.raises(backend.MmsError) as e:       # R: subscript out of bounds
        pipeline.simulate_ephemeral_habitat_vectorised(m["rainfall"][:100], m["air_temperature"][:100],
                                                       m["humidity"][:100], m["windspeed"][:100], case.altitude,
                                                       burnin_years=1)
    assert e.value.code == "MMS_ERR_BURNIN"


def test_monthly_climate_to_vectors_mirror(gpu):
    case = Case(n_cells=30)
    pmc = {"start": case.start, "end": case.end, "altitude": case.altitude,
           "longitude": np.arange(30.0), "latitude": np.arange(30.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = case.monthly[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    res = pipeline.monthly_climate_to_vectors(pmc, lhf, species="An. gambiae", dates_predict=case.dates_predict)
    ref = case.oracle(want_hourly=False)
    assert len(res["start"]) == 12 and str(res["end"][1]) == "2001-02-28"
    for name in pipeline.VECTOR_VARIABLES:
        assert rel_err(res[name], ref["mean"][0, ORACLE_VAR[name]]) <= REL_TOL_FAST, name


def test_several_shards_equal_one(gpu):
    """The cell tile cut into 3 contiguous shards (all on device 0 here; one per GPU on a
    multi-GPU box) gives bit-identical results: no reduction crosses cells."""
    case = Case(n_cells=700, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    kw = case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:8]), want_wet_hours=True)
    one = backend.run_chain(case.monthly, case.altitude, devices=[0], **kw)
    three = backend.run_chain(case.monthly, case.altitude, devices=[0, 0, 0], **kw)
    for k in ("mean_state", "mean_forcing", "wet_hours"):
        assert np.array_equal(one[k], three[k], equal_nan=True), k


def test_full_size_properties_and_sampled_parity(gpu):
    """BASELINE configs[1] at full size (100k cells): invariants over all cells, and oracle
    parity on a random sample of the same cells."""
    habs = [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)]
    case = Case(n_cells=100_000, habitats=habs, out_of_range_fraction=0.01)
    names = list(backend.VAR_NAMES[:8])
    out = backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=names, want_wet_hours=True))
    ms = {v: out["mean_state"][:, k] for k, v in enumerate(out["mean_state_vars"])}
    eps = np.sqrt(np.finfo(float).eps)
    for s, hb in enumerate(habs):
        vmax = oracle.lib().mmso_habitat_max_volume(oracle.habitat(**hb))
        v = ms["water_volume"][s]
        assert np.nanmin(v) >= 0.0 and np.nanmax(v) <= vmax * (1 + 1e-12)     # a mean of values <= vmax
        assert np.nanmin(ms["adult"][s]) >= eps
        wh = out["wet_hours"][s]
        assert wh.min() >= 0 and np.all(wh <= 24 * np.bincount(case.period_of_day)[:, None])
        assert np.all((wh == 0) <= (v == 0.0))                       # never wet <=> zero mean volume
    nan_cells = np.isnan(out["mean_forcing"][3]).any(axis=0)          # ds NA months only in out-of-grid cells
    assert 0 < nan_cells.sum() < 0.02 * case.C
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(case.C, size=96, replace=False))
    ref = oracle.run_chain(np.ascontiguousarray(case.monthly[:, :, pick]), case.mid_day, case.day_predict,
                           case.altitude[pick], [oracle.habitat(**h) for h in habs],
                           oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"]),
                           case.tables, case.period_of_day, case.P, want_hourly=False)
    for k, v in enumerate(out["mean_state_vars"]):
        assert rel_err(out["mean_state"][:, k][..., pick], ref["mean"][:, ORACLE_VAR[v]]) <= REL_TOL_FAST, v
    for k, v in enumerate(out["mean_forcing_vars"]):
        assert rel_err(out["mean_forcing"][k][..., pick], ref["mean"][0, ORACLE_VAR[v]]) <= REL_TOL_FAST, v
    assert np.array_equal(out["wet_hours"][..., pick], ref["wet_hours"])


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_scalar_twins_whole_series_burnin(gpu, mode, tol):
    """simulate_ephemeral_habitat() / simulate_population(): one pixel, vectors in and out, and the
    burn-in repeats the whole series (so a 500-step series with burn-in is legal)."""
    case, m = hourly_inputs(n_cells=2)
    col = {k: v[:, 1] for k, v in m.items()}
    dmh = syn.synthetic_lifehistory("An. gambiae").das_densmod
    dm = oracle.densmod(dmh.type, dmh.dd_effect, dmh.surface_area_cm2)
    for n, burn in ((case.H, 0), (case.H, 1), (500, 2), (7, 3)):
        sub = {k: v[:n] for k, v in col.items()}
        ref = oracle.simulate_ephemeral_habitat(sub["rainfall"], sub["air_temperature"], sub["humidity"],
                                                sub["windspeed"], float(case.altitude[1]), oracle.habitat(), burn)
        got = pipeline.simulate_ephemeral_habitat(sub, float(case.altitude[1]), burnin_years=burn, math_mode=mode)
        assert got.shape == (n,) and rel_err(got, ref) <= tol, (n, burn)
        refp = oracle.simulate_population(sub["larval_habitat_area"], sub["mdr"], sub["efd"], sub["das_zerodensity"],
                                          sub["ds"], dm, burnin_years=burn)
        gotp = pipeline.simulate_population(sub, dmh, burnin_years=burn, math_mode=mode)
        assert rel_err(gotp["aquatic"], refp["aquatic"]) <= tol, (n, burn)
        assert rel_err(gotp["adult"], refp["adult"]) <= tol, (n, burn)


@pytest.mark.parametrize("mode,tol", [("strict", REL_TOL_STRICT), ("fast", REL_TOL_FAST)])
def test_twins_ragged_shapes_with_burnin_and_cylinder(gpu, mode, tol):
    """Both twins on the ragged shapes again, with the other species / water body, a burn-in year where the
    series allows it, and pixel counts around the 32-pixel unit and the 148-SM grid."""
    case, m = hourly_inputs(n_cells=131)
    dmh = syn.synthetic_lifehistory("An. stephensi").das_densmod
    dm = oracle.densmod(dmh.type, dmh.dd_effect, dmh.surface_area_cm2)
    names = ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")
    hab = syn.habitat_cylinder(leak_per_hour=0.003)
    for H, C in RAGGED + ((8760, 131), (30, 64), (30, 97)):
        burn = 1 if H >= 8760 else 0
        sub = {k: np.asfortranarray(v[:H, :C]) for k, v in m.items()}
        ref = oracle.simulate_ephemeral_habitat_vectorised(sub["rainfall"], sub["air_temperature"], sub["humidity"],
                                                           sub["windspeed"], case.altitude[:C], oracle.habitat(**hab), burn)
        got = backend.water_vectorised(sub["rainfall"], sub["air_temperature"], sub["humidity"], sub["windspeed"],
                                       case.altitude[:C], hab, burnin_years=burn, math_mode=mode, want_volume=True,
                                       want_wet=True)
        assert rel_err(got["water_surface_area"], ref["water_surface_area"]) <= tol, (H, C)
        assert rel_err(got["water_volume"], ref["water_volume"]) <= tol, (H, C)
        assert np.array_equal(got["wet"], ref["wet"]), (H, C)
        cols = [sub[k] for k in names]
        refp = oracle.simulate_population_vectorised(*cols, dm, burnin_years=burn)
        gotp = pipeline.simulate_population_vectorised(*cols, dmh, burnin_years=burn, math_mode=mode)
        assert rel_err(gotp["aquatic"], refp["aquatic"]) <= tol, (H, C)
        assert rel_err(gotp["adult"], refp["adult"]) <= tol, (H, C)


@pytest.mark.parametrize("chunks,pass_pixels,workers", [("3", "0", "2"), ("2", "64", "1"), ("5", "96", "3")])
def test_twins_chunked_pipeline_and_passes(gpu, monkeypatch, chunks, pass_pixels, workers):
    """The twins' host pipeline: pixels in several chunks (upload / kernel / download overlapped), several host workers,
    and several passes when the matrices do not fit in device memory (forced here) -- the same bits as one chunk in
    one pass, from pageable and from pinned matrices."""
    case, m = hourly_inputs(n_cells=211)
    hab = syn.habitat_cylinder(leak_per_hour=0.003)
    dmh = syn.synthetic_lifehistory("An. gambiae").das_densmod
    names = ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")

    def both():
        w = backend.water_vectorised(m["rainfall"], m["air_temperature"], m["humidity"], m["windspeed"], case.altitude, hab,
                                     burnin_years=1, want_volume=True, want_wet=True)
        p = pipeline.simulate_population_vectorised(*[m[k] for k in names], dmh, burnin_years=1)
        return w, p

    monkeypatch.setenv("MMS_TWIN_CHUNKS", "1")
    w0, p0 = both()
    assert backend.stage_kernel_ms() > 0
    monkeypatch.setenv("MMS_TWIN_CHUNKS", chunks)
    monkeypatch.setenv("MMS_HOST_WORKERS", workers)
    if pass_pixels != "0":
        monkeypatch.setenv("MMS_TWIN_PASS_PIXELS", pass_pixels)
    w1, p1 = both()
    for k in ("water_surface_area", "water_volume", "wet"):
        assert np.array_equal(w0[k], w1[k], equal_nan=True), k
    for k in ("aquatic", "adult"):
        assert np.array_equal(p0[k], p1[k], equal_nan=True), k
    # pinned matrices: the DMA runs on them in place
    pins = [backend.PinnedArray((case.C, case.H)) for _ in range(5)]             # [pixel][time] = R's column-major H x C
    for pa, k in zip(pins, ("rainfall", "air_temperature", "humidity", "windspeed")):
        pa.array[...] = m[k].T
    w2 = backend.water_vectorised(*[pa.array.T for pa in pins[:4]], case.altitude, hab, burnin_years=1, want_volume=True,
                                  want_wet=True)
    for k in ("water_surface_area", "water_volume", "wet"):
        assert np.array_equal(w0[k], w2[k], equal_nan=True), k
