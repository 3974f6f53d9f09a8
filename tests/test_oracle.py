"""The CPU oracle against (i) the known-answer constants derivable from the R
source, (ii) the properties of the reference's commented
demos, (iii) the independently written numpy restatement, (iv) scipy/numpy for the
third-party algorithms it restates, and (v) the committed golden vectors."""
import os

import numpy as np
import pytest

import oracle
from oracle import restate_numpy as rn
from helpers import Case
from mosmicrosim_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "chain_small.npz")


# ---- known-answer constants ---------------------------------------------------

def test_kat_constants():
    L = oracle.lib()
    assert np.sqrt(np.finfo(float).eps) == 1.4901161193847656e-08
    assert L.mmso_cone_depth_to_volume(1.0) == 1.0471975511965976
    assert 3.046474 * L.mmso_R_pow(1.0471975511965976, 0.666667) == 3.141592812544754
    hb = oracle.habitat()
    assert L.mmso_habitat_catchment(hb) == 3.141592812544754          # "pi", but not exactly
    assert L.mmso_evaporation_rate_kg_h(25.0, 60.0, 500.0, 2.0, 1.0) == pytest.approx(0.5498675133833498, rel=1e-15)
    assert L.mmso_R_pow(2.0 / 10.0, 0.15) == pytest.approx(0.7855150302317644, rel=1e-15)
    assert L.mmso_R_pow(3.0, 2.0) == 9.0 and L.mmso_R_pow(0.0, 0.5) == 0.0 and L.mmso_R_pow(-0.0, 0.666667) == 0.0
    assert L.mmso_cone_volume_to_surface(0.0) == 0.0


def test_kat_sinec_constants():
    # a day with tmin = 0, tmax = 2 (A = 1) exposes the hour-of-day constants directly
    air = oracle.interpolate_daily_air_temp([2.0, 2.0, 2.0], [0.0, 0.0, 0.0])[24:48]
    light = np.array([-0.999999999999966, -0.9009689487307901, -0.6234898892473616, -0.2225209702802615,
                      0.2225209702802615, 0.6234898892473616, 0.9009689487307901, 0.999999999999966,
                      0.9009687224112726, 0.6234894814336972, 0.2225204617449742, -0.22252147881548834,
                      -0.6234902970608567])
    np.testing.assert_allclose(air[5:18], light + 1.0, rtol=0, atol=2e-16)
    z_night = -0.6234902970608567
    e = np.array([1.2840254166877414, 1.6487212707001282, 2.117000016612675, 2.718281828459045,
                  3.4903429574618414, 4.4816890703380645])
    np.testing.assert_allclose(air[18:24], (z_night + 1.0) / e, rtol=1e-15)


def test_kat_vsine_weights():
    w = oracle.vsine_weights(np.arange(1.0, 25.0))
    expect = [7 / 12, 8 / 12, 9 / 12, 10 / 12, 11 / 12, 1, 6 / 7, 5 / 7, 4 / 7, 3 / 7, 2 / 7, 1 / 7, 0] + \
             [k / 22 for k in range(1, 12)]
    np.testing.assert_allclose(w, expect, rtol=0, atol=1.2e-16)


def test_densmod_forms():
    lg = oracle.densmod("logit", syn.DD_EFFECT_AG, syn.MURIU_AREA_CM2)
    cx = oracle.densmod("cox_ph", 0.002, syn.EVANS_AREA_CM2)
    L = oracle.lib()
    assert L.mmso_densmod_apply(lg, 0.9, 0.0) == pytest.approx(0.9, rel=1e-15)
    assert L.mmso_densmod_apply(cx, 0.9, 0.0) == pytest.approx(0.9, rel=1e-15)
    assert L.mmso_densmod_apply(lg, 0.9, 10.0) < 0.9 and L.mmso_densmod_apply(cx, 0.9, 10.0) < 0.9
    assert L.mmso_densmod_apply(lg, 0.9, 1e12) == 0.0          # dry habitat: density / eps -> survival 0
    assert L.mmso_densmod_apply(lg, 0.0, 1.0) == 0.0           # logit(0) = -Inf


# ---- properties from the reference's commented demos ----------------------------

def test_hourly_curve_anchor_properties():
    rng = np.random.default_rng(1)
    n = 12
    tmax, tmin = rng.uniform(25, 32, n), rng.uniform(11, 15, n)
    air = oracle.interpolate_daily_air_temp(tmax, tmin).reshape(n, 24)
    np.testing.assert_allclose(air[:, 12], tmax, rtol=1e-12)      # hour 13 = daily max (R/sim_microclimate.R:300-301)
    np.testing.assert_allclose(air[:, 5], tmin, rtol=1e-12)       # hour 6 = daily min (:302-303)
    hmax, hmin = rng.uniform(85, 92, n), rng.uniform(51, 55, n)
    rh = oracle.vsine_interpolate_daily(hmax, hmin, "max").reshape(n, 24)
    np.testing.assert_allclose(rh[:, 5], hmax, rtol=1e-15)        # max at 6 (:343-344)
    np.testing.assert_allclose(rh[:, 12], hmin, rtol=1e-15)       # min at 13 (:345-346)
    wmax, wmin = rng.uniform(3, 5, n), rng.uniform(1, 1.5, n)
    ws = oracle.vsine_interpolate_daily(wmax, wmin, "min").reshape(n, 24)
    np.testing.assert_allclose(ws[:, 5], wmin, rtol=1e-15)        # min at 6 (:385-386)
    np.testing.assert_allclose(ws[:, 12], wmax, rtol=1e-15)       # max at 13 (:383-384)


def test_water_and_population_invariants():
    case = Case(n_cells=24, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.01)])
    ref = case.oracle()
    for s, hb in enumerate(case.habitats):
        h = oracle.habitat(**hb)
        vmax = oracle.lib().mmso_habitat_max_volume(h)
        vol = ref["hourly"][s, 7]
        area = ref["hourly"][s, 6]
        assert np.nanmin(vol) >= 0.0 and np.nanmax(vol) <= vmax
        assert np.all(area[vol == 0.0] == 0.0)
        assert np.array_equal(ref["wet"][s].astype(bool), vol > 0)
        adult = ref["hourly"][s, 0]
        assert np.nanmin(adult) >= np.sqrt(np.finfo(float).eps)   # imported adults keep it positive
    # leak = 0 cylinder equals ... itself with leak 0.0 bit for bit through the extension term
    a = Case(n_cells=8, habitats=[syn.habitat_cone(leak_per_hour=0.0)]).oracle()
    b = Case(n_cells=8, habitats=[syn.habitat_cone()]).oracle()
    assert np.array_equal(a["hourly"], b["hourly"], equal_nan=True)


def test_na_outside_ds_grid_is_reset_next_step():
    case = Case(n_cells=40, out_of_range_fraction=0.5)
    ref = case.oracle()
    ds, adult = ref["hourly"][0, 5], ref["hourly"][0, 0]
    assert np.isnan(ds).any()
    assert np.array_equal(np.isnan(adult), np.isnan(ds))          # NA that hour only (R/sim_pop_dynamics.R:33-34)
    t, c = np.argwhere(np.isnan(ds[:-1]) & ~np.isnan(ds[1:]))[0]
    assert adult[t + 1, c] >= np.sqrt(np.finfo(float).eps)


def test_extrapolation_and_burnin_errors():
    case = Case(n_cells=2)
    with pytest.raises(ValueError):
        oracle.spline_seasonal(case.monthly[0, :, 0], case.mid_day, case.day_predict - 400.0)
    with pytest.raises(ValueError):
        oracle.run_chain(case.monthly, case.mid_day, case.day_predict[:100], case.altitude,
                         [oracle.habitat()], oracle.densmod("logit", -0.004, 900.0), case.tables,
                         case.period_of_day[:100], case.P)      # burn-in needs >= 8760 steps


# ---- third-party algorithms the oracle restates -----------------------------------

def test_approx_matches_numpy_interp():
    rng = np.random.default_rng(3)
    x = np.cumsum(rng.uniform(0.5, 2.0, 30))
    y = rng.normal(size=30)
    v = rng.uniform(x[0], x[-1], 200)
    got = np.array([oracle.approx1(t, x, y) for t in v])
    np.testing.assert_allclose(got, np.interp(v, x, y), rtol=1e-13, atol=1e-15)
    assert oracle.approx1(x[7], x, y) == y[7] and oracle.approx1(x[-1], x, y) == y[-1]
    assert np.isnan(oracle.approx1(x[0] - 1, x, y)) and np.isnan(oracle.approx1(x[-1] + 1, x, y))


def test_fmm_spline_reproduces_cubics_and_matches_host_builder():
    from mosmicrosim_b200.lifehistory import splinefun_fmm
    x = np.linspace(-3, 5, 41)
    y = 0.3 * x ** 3 - 1.1 * x ** 2 + 0.2 * x + 4.0       # fmm end conditions make cubics exact
    b, c, d = oracle.fmm_spline(x, y)
    tab = {"x": x, "y": y, "b": b, "c": c, "d": d}
    u = np.linspace(-3, 5, 333)
    np.testing.assert_allclose(oracle.spline_eval(tab, u), 0.3 * u ** 3 - 1.1 * u ** 2 + 0.2 * u + 4.0, rtol=1e-11)
    host = splinefun_fmm(x, y)
    assert np.array_equal(host.b, b) and np.array_equal(host.c, c) and np.array_equal(host.d, d)
    # extrapolation uses the end cubics, like R's splinefun
    assert oracle.spline_eval(tab, [6.0])[0] == pytest.approx(0.3 * 216 - 1.1 * 36 + 1.2 + 4.0, rel=1e-9)


def test_interp_surface_matches_scipy_bilinear():
    from scipy.interpolate import RegularGridInterpolator
    tabs = syn.synthetic_lifehistory().tables()["ds"]
    z2 = tabs["z"].reshape(len(tabs["y"]), len(tabs["x"])).T
    rgi = RegularGridInterpolator((tabs["x"], tabs["y"]), z2, method="linear")
    rng = np.random.default_rng(5)
    t, h = rng.uniform(0, 45, 500), rng.uniform(0, 100, 500)
    np.testing.assert_allclose(oracle.interp_surface(tabs, t, h), rgi(np.stack([t, h], 1)), rtol=1e-12)
    out = oracle.interp_surface(tabs, [-0.1, 45.0001, 20.0, 45.0], [50.0, 50.0, 100.5, 100.0])
    assert np.isnan(out[:3]).all() and out[3] == z2[-1, -1]        # NA outside, exact on the last grid line


def test_r_mean_is_long_double_two_pass():
    x = np.array([1e16, 1.0, -1e16, 1.0])
    assert oracle.R_mean(x) == 0.5 and rn.r_mean(x) == 0.5
    assert np.isnan(oracle.R_mean(np.array([1.0, np.nan])))


# ---- double-entry: C restatement vs numpy restatement --------------------------------

@pytest.mark.parametrize("species", ["An. gambiae", "An. stephensi"])
def test_c_oracle_matches_numpy_restatement(species):
    case = Case(n_cells=5, species=species,
                habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    ref = case.oracle()
    dm = dict(type=case.densmod["type_"], dd_effect=case.densmod["dd_effect"],
              surface_area_cm2=case.densmod["surface_area_cm2"])
    alt = rn.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude, case.habitats, dm,
                       case.tables, case.period_of_day, case.P)
    for v, name in enumerate(rn.VAR_NAMES):      # (the oracle's 13th, water_temperature, is the air temperature here)
        a, b = ref["hourly"][:, v], alt["hourly"][:, v]
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(a)
        rel = np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(a[ok]), 1e-300) * (np.abs(a[ok] - b[ok]) > 1e-15))
        assert rel < 1e-11, (name, rel)        # glibc scalar vs numpy SIMD exp/log/pow last bits
    # the state-independent stages are bit identical
    for name in ("mdr", "efd", "das_zerodensity", "ds", "air_temperature", "humidity", "windspeed"):
        v = oracle.VAR_NAMES.index(name)
        assert np.array_equal(ref["hourly"][0, v], alt["hourly"][0, v], equal_nan=True), name
    assert np.array_equal(ref["wet"], alt["wet"]) and np.array_equal(ref["wet_hours"], alt["wet_hours"])
    np.testing.assert_allclose(ref["mean"][:, :len(rn.VAR_NAMES)], alt["mean"], rtol=1e-11, atol=1e-15)


# ---- golden vectors ----------------------------------------------------------------------

def test_golden_vectors():
    g = np.load(GOLDEN)
    case = Case(n_cells=int(g["n_cells"]), habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
    assert np.array_equal(case.monthly, g["monthly"]) and np.array_equal(case.altitude, g["altitude"])
    ref = case.oracle(want_hourly=True)
    nv = g["mean"].shape[1]                      # the golden file holds the restatement's 12 variables
    np.testing.assert_allclose(ref["mean"][:, :nv], g["mean"], rtol=1e-11, atol=1e-15)
    assert np.array_equal(ref["wet_hours"], g["wet_hours"])
    sel = g["hourly_steps"]
    np.testing.assert_allclose(ref["hourly"][:, :nv][:, :, sel, :], g["hourly_sample"], rtol=1e-11, atol=1e-15)


# ---- the one-pixel twins (whole-series burn-in rule) ------------------------------------------

def test_scalar_twins_burnin_repeats_the_whole_series():
    """simulate_ephemeral_habitat() / simulate_population() (R/sim_water.R:151-152,
    R/sim_pop_dynamics.R:147-148): index <- rep(seq_len(n_times), burnin_years + 1)."""
    case = Case(n_cells=2)
    h = case.oracle()["hourly"][0]
    v = {n: h[oracle.VAR_NAMES.index(n)][:, 1] for n in oracle.VAR_NAMES}
    hb, dm = oracle.habitat(), oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"],
                                              case.densmod["surface_area_cm2"])
    wcols = [v[k] for k in ("rainfall", "air_temperature", "humidity", "windspeed")]
    pcols = [v[k] for k in ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")]
    alt = float(case.altitude[1])
    as_mat = lambda cols: [c.reshape(-1, 1) for c in cols]
    # no burn-in, and one burn-in year on exactly 8760 rows: identical to the vectorised functions
    for burn in (0, 1):
        a = oracle.simulate_ephemeral_habitat(*wcols, alt, hb, burn)
        b = oracle.simulate_ephemeral_habitat_vectorised(*as_mat(wcols), [alt], hb, burn)["water_surface_area"][:, 0]
        assert np.array_equal(a, b)
        p = oracle.simulate_population(*pcols, dm, burnin_years=burn)
        q = oracle.simulate_population_vectorised(*as_mat(pcols), dm, burnin_years=burn)
        assert np.array_equal(p["adult"], q["adult"][:, 0]) and np.array_equal(p["aquatic"], q["aquatic"][:, 0])
    # a short series with two burn-in repeats = the series tiled three times, last copy kept
    n = 500
    short_w, short_p = [c[:n] for c in wcols], [c[:n] for c in pcols]
    a = oracle.simulate_ephemeral_habitat(*short_w, alt, hb, 2)
    b = oracle.simulate_ephemeral_habitat(*[np.tile(c, 3) for c in short_w], alt, hb, 0)[-n:]
    assert np.array_equal(a, b)
    p = oracle.simulate_population(*short_p, dm, burnin_years=2)
    q = oracle.simulate_population(*[np.tile(c, 3) for c in short_p], dm, burnin_years=0)
    assert np.array_equal(p["adult"], q["adult"][-n:]) and np.array_equal(p["aquatic"], q["aquatic"][-n:])


# ---- TerraClimate variables -> monthly forcing ------------------------

def test_terraclimate_known_answers():
    L = oracle.lib()
    assert L.mmso_vapprs(0.0) == pytest.approx(610.71, rel=1e-12)          # triple point: 10^log10(6.1071) hPa
    assert L.mmso_vapprs(100.0) == pytest.approx(101324.6, rel=1e-12)      # boiling point: 1013.246 hPa
    assert L.mmso_vapprs(20.0) == pytest.approx(2338.0, rel=2e-3)          # Goff-Gratch, ~2.34 kPa at 20 C
    assert L.mmso_vapprs(-10.0) == pytest.approx(259.9, rel=2e-3)          # over ice
    assert L.mmso_adjust_wind_speed(1.0, 2.0, 10.0, 0.15) == 0.7855150302317644     # SURVEY 8c KAT
    assert L.mmso_cloud_cover(100.0, 200.0, 1.2) == pytest.approx(60.0) and L.mmso_cloud_cover(100.0, 200.0, 0.8) == pytest.approx(40.0)
    assert L.mmso_cloud_cover(250.0, 200.0, 1.2) == 0.0 and L.mmso_cloud_cover(1.0, 200.0, 1.2) == 100.0   # clamps
    vp = L.mmso_vapour_pressure(20.0, 1.0)
    assert vp == pytest.approx(L.mmso_vapprs(20.0) - 1000.0)
    assert L.mmso_relative_humidity(20.0, vp) == pytest.approx(100 * vp / L.mmso_vapprs(20.0))
    assert L.mmso_relative_humidity(20.0, -5.0) == 0.01 and L.mmso_relative_humidity(20.0, 1e6) == 100.0


def test_process_terraclimate_pixel_vars_and_month_bookkeeping():
    from mosmicrosim_b200 import dates
    start, end = syn.month_table(2001, 1)
    month_id, n_days = dates.terraclimate_month_info(start, end)
    assert list(month_id) == [12] + list(range(1, 13)) + [1]
    assert list(n_days[:4]) == [30.0, 30.0, 27.0, 30.0]                  # days in the month MINUS ONE, as the reference
    tc, clear_sky, _ = syn.synthetic_terraclimate(7, start)
    m = oracle.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
    names = ["tmin", "tmax", "rhmin", "rhmax", "wsmin", "wsmax", "ccmin", "ccmax", "rainfall_daily"]
    v = dict(zip(names, m))
    assert np.array_equal(v["tmax"], tc[0]) and np.array_equal(v["tmin"], tc[1])
    assert np.all(v["rhmin"] <= v["rhmax"]) and v["rhmin"].min() >= 0.01 and v["rhmax"].max() <= 100.0
    assert np.array_equal(v["wsmin"], v["wsmax"] * 0.1) and np.allclose(v["wsmax"], tc[3] * 0.7855150302317644, rtol=1e-15)
    assert np.all(v["ccmin"] <= v["ccmax"]) and v["ccmin"].min() >= 0 and v["ccmax"].max() <= 100
    assert (v["ccmax"] == 0).any() and (v["rhmin"] == 0.01).any()        # the synthetic data reaches the clamps
    assert np.array_equal(v["rainfall_daily"], tc[2] / n_days[:, None])
    # clear-sky values are looked up by calendar month, not by row
    assert np.array_equal(m[7][0], m[7][12] * 0 + oracle.process_terraclimate_pixel_vars(
        tc[:, :1], clear_sky, month_id[:1], n_days[:1])[7][0])


def test_pad_last_month():
    from mosmicrosim_b200 import dates
    start, end = syn.month_table(2001, 3)                                 # Dec 2000 .. Jan 2004
    vals = np.arange(2 * len(start) * 2, dtype=np.float64).reshape(2, len(start), 2) ** 1.5
    s2, e2, v2 = dates.pad_last_month(start, end, vals)
    assert str(s2[-1]) == "2004-02-01" and str(e2[-1]) == "2004-02-29" and v2.shape == (2, len(start) + 1, 2)
    feb = [i for i, s in enumerate(start) if str(s)[5:7] == "02"]          # Feb 2001, 2002, 2003
    assert len(feb) == 3
    np.testing.assert_allclose(v2[:, -1], vals[:, feb].mean(axis=1), rtol=1e-15)
    assert np.array_equal(v2[:, :-1], vals)
    assert dates.r_mean([1.0, 2.0, 4.0]) == oracle.R_mean(np.array([1.0, 2.0, 4.0]))
    # fewer matching years than n_previous: the mean of those present; none: NaN like mean(numeric(0))
    s3, e3, v3 = dates.pad_last_month(start[:3], end[:3], vals[:, :3])
    assert str(s3[-1]) == "2001-03-01" and np.isnan(v3[:, -1]).all()


def test_terraclimate_c_oracle_matches_numpy_restatement():
    from mosmicrosim_b200 import dates
    start, end = syn.month_table(2003, 2)
    month_id, n_days = dates.terraclimate_month_info(start, end)
    tc, clear_sky, _ = syn.synthetic_terraclimate(50, start)
    tc[0, 5, :3] = [-2.0, 0.0, 0.5]; tc[1, 5, :3] = [-20.0, -9.0, -0.5]        # both branches of VAPPRS
    a = oracle.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
    b = rn.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
    np.testing.assert_allclose(a, b, rtol=1e-13, atol=0)            # glibc scalar vs numpy log10 / pow last bits
    for v in (0, 1, 4, 5, 8):                                       # no transcendental involved: bit identical
        assert np.array_equal(a[v], b[v])


def test_golden_front_end_and_twins():
    """tests/golden/front_end_and_twins.npz (written by the numpy restatement, make_golden.py): the TerraClimate front
    end and the one-pixel twins with two whole-series burn-in repeats, against the C oracle."""
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "front_end_and_twins.npz"))
    got = oracle.process_terraclimate_pixel_vars(g["terraclimate"], g["clear_sky"], g["month_id"], g["n_days"])
    np.testing.assert_allclose(got, g["monthly"], rtol=1e-13, atol=0)
    s = dict(zip(oracle.VAR_NAMES, g["series"]))
    case = Case(n_cells=2)
    hab = oracle.habitat(**syn.habitat_cone())
    dm = oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"])
    area = oracle.simulate_ephemeral_habitat(s["rainfall"], s["air_temperature"], s["humidity"], s["windspeed"],
                                             float(g["altitude"]), hab, 2)
    np.testing.assert_allclose(area, g["water_surface_area_burnin2"], rtol=1e-11, atol=1e-15)
    pop = oracle.simulate_population(s["larval_habitat_area"], s["mdr"], s["efd"], s["das_zerodensity"], s["ds"], dm,
                                     burnin_years=2)
    np.testing.assert_allclose(pop["aquatic"], g["aquatic_burnin2"], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(pop["adult"], g["adult_burnin2"], rtol=1e-11, atol=1e-15)


# ---- water-temperature EXTENSION (query_terraclimate.R:370-497) ------------------------------

def test_solar_geometry_known_answers():
    """Declination: ~0 at the March equinox (doy 80), +-23.44 degrees at the solstices; hour angles symmetric about
    noon; the hourly clear-sky curve keeps the daily mean it is given, is zero at night and peaks at noon."""
    sd, cd = oracle.solar_declination(80)
    assert abs(sd) < 1e-12 and abs(cd - 1) < 1e-12                       # zeta = 0 exactly on day 80
    assert abs(np.degrees(np.arcsin(oracle.solar_declination(172)[0])) - 23.44) < 0.05
    assert abs(np.degrees(np.arcsin(oracle.solar_declination(355)[0])) + 23.44) < 0.05
    assert oracle.cos_hour_angle(12) == 1.0 and abs(oracle.cos_hour_angle(18)) < 1e-15
    assert oracle.cos_hour_angle(7) == pytest.approx(oracle.cos_hour_angle(17), abs=1e-15)
    for lat, doy in ((0.0, 80), (-33.0, 200), (52.0, 355), (66.0, 172)):
        irr = oracle.clear_sky_hourly(lat, doy, 250.0)
        assert irr.mean() == pytest.approx(250.0, rel=1e-12) and irr.min() == 0.0 and irr.argmax() == 11
    eq = oracle.clear_sky_hourly(0.0, 80, 250.0)                          # equator, equinox: cos Z = cos h, sun up 7h..17h
    assert np.count_nonzero(eq > 1e-9) == 11 and eq[11] == pytest.approx(250.0 * 24 / sum(max(0.0, np.cos(np.radians(15 * (k - 12)))) for k in range(1, 25)))
    assert np.all(oracle.clear_sky_hourly(80.0, 355, 250.0) == 0.0)       # polar night


def test_water_temperature_model_properties():
    """Default off: the water temperature is the air temperature and nothing else changes.  On: a fully shaded water
    body is never warmer than the air (evaporative loss, no solar gain); sun warms it; strong convection ties it to
    the air; a deep, full body barely moves in an hour; the volume series does not depend on the temperature model."""
    from helpers import Case, ORACLE_VAR
    from mosmicrosim_b200 import dates
    case = Case(n_cells=6, out_of_range_fraction=0.0)
    dm = case.densmod
    args = (case.monthly, case.mid_day, case.day_predict, case.altitude, [oracle.habitat()],
            oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), case.tables, case.period_of_day, case.P)
    base = oracle.run_chain(*args, want_hourly=True)
    h0 = base["hourly"][0]
    assert np.array_equal(h0[ORACLE_VAR["water_temperature"]], h0[ORACLE_VAR["air_temperature"]])
    wt = dict(day_of_year=dates.day_of_year(case.dates_predict), month_of_day=dates.month_of_year(case.dates_predict),
              latitude=np.linspace(-30, 30, 6), clear_sky=np.full((12, 6), 280.0))

    def run(**over):
        h = oracle.run_chain(*args, want_hourly=True, water_temperature=dict(wt, **over))["hourly"][0]
        return h[ORACLE_VAR["water_temperature"]], h[ORACLE_VAR["air_temperature"]], h[ORACLE_VAR["water_volume"]]

    tw, ta, vol = run(shade_proportion=1.0)
    assert np.array_equal(vol, h0[ORACLE_VAR["water_volume"]])           # same water, whatever its temperature
    assert np.all(tw <= ta + 1e-9) and (ta - tw).mean() > 5.0            # the notes' shaded model: evaporative cooling
    tw_sun, _, _ = run(shade_proportion=0.0)
    assert (tw_sun - tw).max() > 10.0 and np.all(tw_sun >= tw - 1e-9)
    tw_conv, _, _ = run(shade_proportion=1.0, convection_w_m2_k=1e4)
    assert np.abs(tw_conv - ta).max() < 0.5
    assert np.array_equal(tw[vol == 0.0], ta[vol == 0.0])                # a dry body has the air's temperature
