"""CPU-only tests: host-side mirror of the reference interface (dates, tables,
argument handling), the C-ABI library loads and exports every symbol that
include/mmsgpu.h declares, there is no CPU fallback, and the one-process-per-GPU
sharding logic is equivalent to a single run (gloo, world size 2)."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest

import oracle
from helpers import Case
from mosmicrosim_b200 import backend, dates, lifehistory as lh, pipeline, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the C ABI ---------------------------------------------------------------------

def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mmsgpu.h")).read()
    declared = set(re.findall(r"MMS_API\s+[\w\s\*]*?\b(mms_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    L = ctypes.CDLL(backend.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(backend.EXPORTS), declared ^ set(backend.EXPORTS)
    assert L.mms_abi_version() == 2


def test_struct_layouts_match_the_header(tmp_path):
    # sizes the C compiler gives the structs in include/mmsgpu.h
    src = '#include <stdio.h>\n#include "mmsgpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu",' \
          "sizeof(mms_habitat),sizeof(mms_densmod),sizeof(mms_spline),sizeof(mms_surface)," \
          "sizeof(mms_lifehistory),sizeof(mms_chain_config),sizeof(mms_chain_outputs));return 0;}"
    exe = str(tmp_path / "mms_sizes")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(),
                   check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    got = [ctypes.sizeof(t) for t in (backend.Habitat, backend.Densmod, backend.Spline, backend.Surface,
                                      backend.LifehistoryC, backend.ChainConfig, backend.ChainOutputs)]
    assert got == sizes


def test_r_shim_compiles_against_the_abi_header():
    """r-binding/src/mmsgpu_shim.c (the .Call shim for the reference's src/) against a stub of
    R's C API: every libmmsgpu call in it must match include/mmsgpu.h."""
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror=implicit-function-declaration",
                        "-Werror=incompatible-pointer-types", "-Werror=int-conversion", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "r-binding", "src", "mmsgpu_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _r_call_sites(text):
    """(.Call name, number of positional arguments after the name) for every .Call in an R source."""
    out, k = [], 0
    while True:
        k = text.find(".Call(", k)
        if k < 0:
            return out
        i, depth, args, cur, quote = k + 6, 1, [], "", None
        while depth:
            ch = text[i]
            if quote:
                quote = None if ch == quote else quote
            elif ch in "\"'":
                quote = ch
            elif ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            elif ch == "," and depth == 1:
                args.append(cur.strip())
                cur, ch = "", ""
            cur += ch
            i += 1
        args.append(cur.strip())
        positional = [a for a in args[1:] if not a.startswith("PACKAGE")]
        out.append((args[0].strip("\"'"), len(positional)))
        k = i


def test_r_calls_match_the_registration_table():
    """Every .Call in r-binding/R/gpu_backend.R names an entry point the shim registers, with the registered number
    of arguments (R checks this only at call time; R is not installed here)."""
    import re
    shim = open(os.path.join(ROOT, "r-binding", "src", "mmsgpu_shim.c")).read()
    registered = {m.group(1): int(m.group(2)) for m in
                  re.finditer(r'\{"(mms_R_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', shim)}
    defined = {m.group(1): m.group(2).count("SEXP") for m in
               re.finditer(r"^SEXP (mms_R_\w+)\(([^)]*)\)", shim, re.M)}
    assert registered and registered == defined, (registered, defined)
    sites = _r_call_sites(open(os.path.join(ROOT, "r-binding", "R", "gpu_backend.R")).read())
    assert sites
    for name, n_args in sites:
        assert registered.get(name) == n_args, (name, n_args, registered.get(name))
    assert {n for n, _ in sites} | {"mms_R_device_count", "mms_R_terraclimate_to_monthly"} >= set(registered)


def test_host_copy_and_worker_windows(tmp_path):
    """copy_rows() (AVX2 streaming stores, ragged heads / tails / pitches) against memcpy, and the tiling of a chunk by
    its workers' windows: tests/c_driver/host_copy_test.cpp, g++ only."""
    exe = str(tmp_path / "host_copy_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror",
                    os.path.join(ROOT, "tests", "c_driver", "host_copy_test.cpp"),
                    os.path.join(ROOT, "mosmicrosim_b200", "csrc", "host_copy.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_na_knots_are_filled_like_approx_drops_them():
    """stats::approx() drops NA knots (restated in the oracle's spline_seasonal); pipeline.fill_na_knots() replaces an
    NA knot by the linear interpolation of its neighbours -- the same daily series to rounding.  Leading / trailing NA
    knots cannot be filled: the reference has NA days there."""
    start, end = syn.month_table(2001, 1)
    dp = syn.year_dates(2001, 1)
    mid_day, day_predict = dates.datenums(start, end, dp)
    monthly, _ = syn.synthetic_forcing(8, start, seed=11)                  # [9][M][C]
    rng = np.random.default_rng(5)
    holes = monthly.copy()
    M = holes.shape[1]
    for c in range(6):                                                     # interior NA knots, single and in runs
        for v in rng.choice(9, size=3, replace=False):
            m0 = int(rng.integers(1, M - 3))
            holes[v, m0:m0 + int(rng.integers(1, 3)), c] = np.nan
    filled, n = pipeline.fill_na_knots(holes, start, end)
    assert n == int(np.isnan(holes).sum()) and np.isfinite(filled).all()
    assert np.array_equal(filled[~np.isnan(holes)], holes[~np.isnan(holes)])          # the other knots keep their bits
    for c in range(8):
        ref = oracle.daily_from_monthly(holes[:, :, c], mid_day, day_predict)          # drops the NA knots
        got = oracle.daily_from_monthly(filled[:, :, c], mid_day, day_predict)
        assert np.isfinite(ref).all()
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-12
    # a leading NA knot stays NA (the reference: NA days, not the GAM) ...
    edge = monthly.copy()
    edge[2, 0, 0] = np.nan
    filled, n = pipeline.fill_na_knots(edge, start, end)
    assert n == 0 and np.isnan(filled[2, 0, 0])
    ref = oracle.daily_from_monthly(edge[:, :, 0], mid_day, day_predict)
    assert np.isnan(ref[2, 0]) and np.isfinite(ref[2, -1])
    # ... and approx() needs two knots
    edge[2, 1:, 0] = np.nan
    with pytest.raises(ValueError, match="two non-NA"):
        oracle.daily_from_monthly(edge[:, :, 0], mid_day, day_predict)


def test_no_cpu_fallback_without_a_device():
    if backend.device_count() > 0:
        pytest.skip("a GPU is visible")
    case = Case(n_cells=4)
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain(case.monthly, case.altitude, **case.gpu_kwargs(mean_vars=["adult"]))
    assert e.value.code == "MMS_ERR_NO_DEVICE"
    z = np.zeros((24, 2))
    with pytest.raises(backend.MmsError) as e:
        backend.water_vectorised(z, z, z, z, np.zeros(2), backend.make_habitat())
    assert e.value.code == "MMS_ERR_NO_DEVICE"
    with pytest.raises(backend.MmsError) as e:
        pipeline.simulate_population_vectorised(z, z, z, z, z, lh.make_surv_densmod_function(-0.004, 900.0, "logit"),
                                                burnin_years=0)
    assert e.value.code == "MMS_ERR_NO_DEVICE"


def test_argument_validation_happens_before_any_device_work():
    case = Case(n_cells=4)
    bad = dict(case.gpu_kwargs(mean_vars=["adult"]))
    bad["habitats"] = []                                        # n_settings = 0
    with pytest.raises(backend.MmsError) as e:
        backend.run_chain(case.monthly, case.altitude, **bad)
    assert e.value.code == "MMS_ERR_INVALID_ARG"
    z = np.zeros((100, 2))
    with pytest.raises(backend.MmsError) as e:                  # R: subscript out of bounds
        backend.population_vectorised(z, z, z, z, z, backend.make_densmod("logit", -0.004, 900.0), burnin_years=1)
    assert e.value.code == "MMS_ERR_BURNIN"


def test_shard_ranges_cover_the_raster():
    for n, g in ((100_000, 8), (1_500_000, 8), (1000, 3), (127, 4), (129, 2), (5, 8)):
        ranges = [backend.shard_range(n, g, i) for i in range(g)]
        pos = 0
        for c0, cnt in ranges:
            assert c0 == pos and cnt >= 0
            pos += cnt
            assert c0 % 128 == 0 or cnt == 0
        assert pos == n
        assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 256      # balanced to 2 blocks


# ---- dates ---------------------------------------------------------------------------

def test_mid_dates_and_default_dates_predict():
    start, end = syn.month_table(2001, 1)
    assert len(start) == 14 and str(start[0]) == "2000-12-01" and str(end[-1]) == "2002-01-31"
    mid = dates.month_mid_days(start, end)
    d0 = np.datetime64("2000-12-01").astype("datetime64[D]").astype(np.int64)
    assert mid[0] == d0 + 15.0                                   # 31-day month: start + 15
    assert mid[2] - (d0 + 31 + 31) == 13.5                       # Feb 2001 (28 days): start + 13.5
    dp = dates.default_dates_predict(start, end)
    assert str(dp[0]) == "2000-12-16" and str(dp[-1]) == "2002-01-16"
    mid_day, day_predict = dates.datenums(start, end, dp)
    assert mid_day[0] == 0.0 and day_predict[0] == 0.0 and not dates.needs_extrapolation(mid_day, day_predict)
    # a 30-day first month (mid = start + 14.5, rounded limit start + 14) triggers the GAM branch in R
    s2 = np.array(["2001-04-01", "2001-05-01", "2001-06-01"], "datetime64[D]")
    e2 = np.array(["2001-04-30", "2001-05-31", "2001-06-30"], "datetime64[D]")
    md2, dp2 = dates.datenums(s2, e2, dates.default_dates_predict(s2, e2))
    assert dates.needs_extrapolation(md2, dp2)


def test_period_index_month_day_included():
    dp = syn.year_dates(2001, 1)
    pod, ps, pe = dates.period_index(dp, "month")
    assert len(ps) == 12 and str(ps[1]) == "2001-02-01" and str(pe[1]) == "2001-02-28"
    assert np.bincount(pod).tolist() == [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    inc = dp[40:100]
    pod2, ps2, _ = dates.period_index(dp, "month", inc)
    assert (pod2 >= 0).sum() == 60 and len(ps2) == 3 and pod2[39] == -1 and pod2[40] == 0
    pod3, ps3, pe3 = dates.period_index(dp, "day", inc)
    assert len(ps3) == 60 and np.array_equal(ps3, pe3)
    with pytest.raises(ValueError):
        dates.period_index(dp, "week")


# ---- life-history table carriers --------------------------------------------------------

def test_lifehistory_tables_follow_the_reference_shapes():
    L = syn.synthetic_lifehistory("An. gambiae")
    t = L.tables()
    assert len(t["mdr"]["x"]) == 1001 and t["mdr"]["x"][0] == -20.0 and t["mdr"]["x"][-1] == 80.0
    assert np.array_equal(t["mdr"]["x"], t["efd"]["x"]) and np.array_equal(t["mdr"]["x"], t["das"]["x"])
    g = t["ds"]
    assert len(g["x"]) == 250 and len(g["y"]) == 250 and g["z"].shape == (62500,)
    assert g["x"][0] == 0.0 and g["x"][-1] == 45.0 and g["y"][-1] == 100.0
    assert g["x"][1] == 45.0 / 249                              # seq(0, 45, length.out = 250)
    assert L.das_densmod.type == "logit" and syn.synthetic_lifehistory("An. stephensi").das_densmod.type == "cox_ph"
    with pytest.raises(ValueError):
        lh.make_surv_densmod_function(1.0, 1.0, "probit")
    assert not callable(L.mdr_temp)                             # table carriers, no CPU evaluation path


def test_pipeline_argument_handling_mirrors_the_reference():
    case = Case(n_cells=3)
    pmc = {"start": case.start, "end": case.end, "altitude": case.altitude,
           "longitude": np.arange(3.0), "latitude": np.arange(3.0)}
    for v, name in enumerate(pipeline.FORCING_VARS):
        pmc[name] = case.monthly[v]
    lhf = {"An_gambiae": case.lh, "An_stephensi": syn.synthetic_lifehistory("An. stephensi")}
    with pytest.raises(ValueError):
        pipeline.monthly_climate_to_vectors(pmc, lhf, species="An. funestus")
    # the reference stop()s on model_water_temperature = TRUE (R/pipeline_functions.R:334-338); here it is modelled
    # (an extension) and needs the pixels' clear-sky radiation next to their latitude
    with pytest.raises(ValueError, match="clear_sky_SOLR"):
        pipeline.monthly_climate_to_vectors(pmc, lhf, model_water_temperature=True)
    with pytest.raises(NotImplementedError):                     # the stage-level verb points to the fused chain
        pipeline.simulate_hourly_conditions({}, model_water_temperature=True)
    if backend.device_count() == 0:
        with warnings.catch_warnings(record=True) as wlist:      # R/pipeline_functions.R:327-332
            warnings.simplefilter("always")
            with pytest.raises(backend.MmsError):
                pipeline.monthly_climate_to_vectors(pmc, lhf, water_shade_proportion=0.5)
            assert any("water_shade_proportion" in str(x.message) for x in wlist)


def test_half_day_overhang_of_the_default_dates_is_detected_and_clipped():
    """A series that starts in a 30-day month (April) and ends in a non-leap February: the reference's default
    dates_predict (midpoint rounded half to even, R/sim_microclimate.R:14-24) starts half a day BEFORE the first knot
    (unrounded midpoint, :31) and ends half a day AFTER the last one -> its GAM branch.  The library refuses
    (MMS_ERR_EXTRAPOLATION); the mirrors can clip those days with a warning."""
    months = np.datetime64("2001-04", "M") + np.arange(11)                 # April 2001 .. February 2002
    start = months.astype("datetime64[D]")
    end = (months + 1).astype("datetime64[D]") - np.timedelta64(1, "D")
    dp = dates.default_dates_predict(start, end)
    mid_day, day_predict = dates.datenums(start, end, dp)
    assert mid_day[0] == 0.0 and day_predict[0] == -0.5 and day_predict[-1] == mid_day[-1] + 0.5
    assert dates.needs_extrapolation(mid_day, day_predict)
    clipped, dropped = dates.clip_dates_predict(start, end, dp)
    assert dropped == 2 and len(clipped) == len(dp) - 2
    assert not dates.needs_extrapolation(*dates.datenums(start, end, clipped))
    # a December .. January series (what monthly_summary_days() produces, R/terraclimate.R:91-98) needs no clipping
    s2, e2 = syn.month_table(2001, 1)
    assert dates.clip_dates_predict(s2, e2, dates.default_dates_predict(s2, e2))[1] == 0


# ---- one process per GPU: shards == whole (gloo, world size 2, CPU) --------------------------

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle
import oracle
from helpers import Case
from mosmicrosim_b200 import backend
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = Case(n_cells=300)                       # the same raster on every rank
c0, n = backend.shard_range(case.C, world, rank)
sub = oracle.run_chain(np.ascontiguousarray(case.monthly[:, :, c0:c0 + n]), case.mid_day, case.day_predict,
                       case.altitude[c0:c0 + n], [oracle.habitat(**h) for h in case.habitats],
                       oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"]),
                       case.tables, case.period_of_day, case.P, want_hourly=False, n_threads=2)
mine = torch.zeros((1, oracle.NVARS, case.P, 256), dtype=torch.float64)
mine[..., :n] = torch.from_numpy(sub["mean"])
parts = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)                   # the gather at the end: the only communication
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)       # bench.py's max-over-ranks timing reduction
if rank == 0:
    whole = oracle.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude,
                             [oracle.habitat(**h) for h in case.habitats],
                             oracle.densmod(case.densmod["type_"], case.densmod["dd_effect"], case.densmod["surface_area_cm2"]),
                             case.tables, case.period_of_day, case.P, want_hourly=False, n_threads=2)["mean"]
    cat = []
    for r in range(world):
        _, nr = backend.shard_range(case.C, world, r)
        cat.append(parts[r][..., :nr].numpy())
    cat = np.concatenate(cat, axis=-1)
    assert cat.shape == whole.shape, (cat.shape, whole.shape)
    assert np.array_equal(cat, whole, equal_nan=True), "concatenated shards differ from the single run"
    assert float(t.item()) == world
    print("SHARDS_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_sharding_equals_single_run(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDS_OK" in r.stdout


def test_plain_c_caller_builds_and_fails_loudly_without_a_device(tmp_path):
    """tests/c_driver/abi_driver.c compiles as C99 with -Wall -Wextra -Werror against include/mmsgpu.h and links
    against libmmsgpu.so; on a box without a GPU its argument check still answers (status + message) and the
    real call comes back with MMS_ERR_NO_DEVICE -- no CPU fallback."""
    from helpers import Case, build_c_driver, write_c_driver_inputs
    exe = build_c_driver(tmp_path)
    if backend.lib().mms_device_count() > 0:
        pytest.skip("a GPU is visible: the run itself is tests/test_gpu_chain.py::test_plain_c_caller")
    case = Case(n_cells=4)
    write_c_driver_inputs(case, list(backend.VAR_NAMES[:6]), tmp_path, 2, 4)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 5 and "no CUDA device" in r.stderr, (r.returncode, r.stdout, r.stderr)
