"""Parity against the REAL reference, for whoever has R: baseline/run_reference.R runs goldingn/mosmicrosim's own
functions on the committed inputs (tests/golden/r_reference/inputs, written by tests/golden/make_r_inputs.py) and
writes every stage's output to tests/golden/r_reference/outputs/.  When those files exist, the CPU oracle (here) and
the CUDA path (-m gpu) are compared with them stage by stage; when they do not (R is not installed in this
repository's development image: parity is "unpinned") the comparisons are skipped -- but the
reader and the comparison code still run on a stand-in written by the oracle in the R script's file layout, so the
hook itself is tested.

    Rscript baseline/run_reference.R --source /path/to/mosmicrosim && python -m pytest tests/test_r_reference.py
"""
import os
import sys

import numpy as np
import pytest

import oracle
from helpers import ORACLE_VAR, REL_TOL_FAST, rel_err
from mosmicrosim_b200 import dates, synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_r_inputs  # noqa: E402

R_DIR = os.path.join(HERE, "golden", "r_reference")
SPECIES = make_r_inputs.SPECIES
HOURLY_VARS = ("air_temperature", "humidity", "cloudcover", "windspeed", "water_temperature", "rainfall")
VECTOR_VARS = ("adult", "aquatic", "mdr", "efd", "das_zerodensity", "ds")
MATRICES = ("water_surface_area", "mdr", "efd", "das_zerodensity", "ds", "aquatic", "adult")
# the reference's mean() is a long-double two-pass mean, the restatements sum in double: far below 1e-9
TOL = 1e-9


def r_array(path, dims):
    """An array written by R's writeBin(as.vector(x)) with dim(x) = dims (first index fastest)."""
    return np.fromfile(path, dtype="<f8").reshape(dims[::-1]).transpose(*range(len(dims) - 1, -1, -1))


def load_r_outputs(out_dir, species_dir):
    d = os.path.join(out_dir, species_dir)
    C, D, H, P = (int(x) for x in open(os.path.join(d, "dims.txt")).read().split())
    out = {"daily": r_array(os.path.join(d, "daily.f64"), (D, 9, C)),
           "hourly": r_array(os.path.join(d, "hourly.f64"), (H, 6, C)),
           "monthly_means": r_array(os.path.join(d, "monthly_means.f64"), (P, 6, C))}
    for name in MATRICES:
        out[name] = r_array(os.path.join(d, name + ".f64"), (H, C))
    return out


def case(species):
    start, end, dates_predict, monthly, altitude = make_r_inputs.inputs()
    lh = syn.synthetic_lifehistory(species)
    mid_day, day_predict = dates.datenums(start, end, dates_predict)
    pod, ps, _ = dates.period_index(dates_predict, "month")
    dm = lh.das_densmod
    return dict(monthly=monthly, altitude=altitude, lh=lh, mid_day=mid_day, day_predict=day_predict, pod=pod,
                P=len(ps), densmod=dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2))


def oracle_stages(c):
    """The oracle's outputs in the shapes load_r_outputs() returns."""
    dm = c["densmod"]
    ref = oracle.run_chain(c["monthly"], c["mid_day"], c["day_predict"], c["altitude"], [oracle.habitat()],
                           oracle.densmod(dm["type_"], dm["dd_effect"], dm["surface_area_cm2"]), c["lh"].tables(),
                           c["pod"], c["P"], burnin_years_pop=1, want_hourly=True)
    h = ref["hourly"][0]                       # [var][H][C]
    C = c["monthly"].shape[2]
    # the two state-free stages cell by cell, as the reference runs them (cloud cover is not in the chain's hourly set)
    daily = np.stack([oracle.daily_from_monthly(c["monthly"][:, :, k], c["mid_day"], c["day_predict"]) for k in range(C)], axis=2)
    hourly = [oracle.hourly_from_daily_ambient(daily[:, :, k]) for k in range(C)]
    out = {"daily": np.transpose(daily, (1, 0, 2)),                                        # [D][9][C]
           "hourly": np.stack([np.stack([hk[v] for hk in hourly], axis=1) for v in HOURLY_VARS], axis=1),   # [H][6][C]
           "monthly_means": np.stack([ref["mean"][0, ORACLE_VAR[v]] for v in VECTOR_VARS], axis=1),
           "water_surface_area": h[ORACLE_VAR["larval_habitat_area"]]}
    assert np.array_equal(out["hourly"][:, 0], h[ORACLE_VAR["air_temperature"]])            # the chain's own hourly series
    for v in ("mdr", "efd", "das_zerodensity", "ds", "aquatic", "adult"):
        out[v] = h[ORACLE_VAR[v]]
    return out


def write_like_r(stages, out_dir, species_dir):
    """The stand-in for R: the same files, the same (column-major) layout."""
    d = os.path.join(out_dir, species_dir)
    os.makedirs(d, exist_ok=True)
    for name, a in stages.items():
        np.ascontiguousarray(np.transpose(a, range(a.ndim - 1, -1, -1)), dtype="<f8").tofile(os.path.join(d, name + ".f64"))
    H, C = stages["adult"].shape
    with open(os.path.join(d, "dims.txt"), "w") as f:
        f.write(f"{C} {stages['daily'].shape[0]} {H} {stages['monthly_means'].shape[0]}\n")


def compare_stages(got, ref, tol, skip=()):
    worst = {}
    for name in ref:
        if name in skip:
            continue
        worst[name] = rel_err(got[name], ref[name])
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, f"beyond {tol}: {bad} (all: {worst})"
    return worst


def r_outputs_present():
    return all(os.path.exists(os.path.join(R_DIR, "outputs", d, "dims.txt")) for d in SPECIES)


def test_committed_inputs_are_reproducible(tmp_path):
    """tests/golden/make_r_inputs.py regenerates the committed input files bit for bit."""
    out = make_r_inputs.main(str(tmp_path / "inputs"))
    for root, _, files in os.walk(out):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), out)
            assert open(os.path.join(root, f), "rb").read() == open(os.path.join(R_DIR, "inputs", rel), "rb").read(), rel


@pytest.mark.parametrize("species_dir", list(SPECIES))
def test_hook_reads_the_r_layout(tmp_path, species_dir):
    """The reader and the stage-by-stage comparison, exercised on a stand-in written in the R script's layout."""
    stages = oracle_stages(case(SPECIES[species_dir]))
    write_like_r(stages, str(tmp_path), species_dir)
    back = load_r_outputs(str(tmp_path), species_dir)
    assert back["daily"].shape == (365, 9, make_r_inputs.N_CELLS) and back["adult"].shape == (8760, make_r_inputs.N_CELLS)
    compare_stages(back, stages, 0.0)
    assert np.isnan(stages["ds"]).any()                     # the committed cells do exercise the NA path


@pytest.mark.parametrize("species_dir", list(SPECIES))
def test_oracle_matches_r(species_dir):
    if not r_outputs_present():
        pytest.skip("no R outputs: run `Rscript baseline/run_reference.R --source <mosmicrosim checkout>` (R is not "
                    "installed in this image; parity unpinned)")
    ref = load_r_outputs(os.path.join(R_DIR, "outputs"), species_dir)
    print(compare_stages(oracle_stages(case(SPECIES[species_dir])), ref, TOL))


def gpu_stages(c, mode):
    """The CUDA path's outputs (one fused-chain call + the daily verb) in the shapes load_r_outputs() returns."""
    from mosmicrosim_b200 import backend
    out = backend.run_chain(c["monthly"], c["altitude"], mid_day=c["mid_day"], day_predict=c["day_predict"],
                            habitats=[syn.habitat_cone()], densmod=c["densmod"], tables=c["lh"].tables(),
                            period_of_day=c["pod"], n_periods=c["P"], burnin_years_pop=1, mean_vars=list(VECTOR_VARS),
                            hourly_vars=list(backend.VAR_NAMES[:7]) + ["air_temperature", "humidity", "windspeed",
                                                                        "rainfall", "cloudcover"],
                            math_mode=mode)
    hs = {v: out["hourly_state"][0, k] for k, v in enumerate(out["hourly_state_vars"])}
    hf = {v: out["hourly_forcing"][k] for k, v in enumerate(out["hourly_forcing_vars"])}
    got = {"hourly": np.stack([hf["air_temperature"], hf["humidity"], hf["cloudcover"], hf["windspeed"],
                               hf["air_temperature"], hf["rainfall"]], axis=1),       # water_temperature = air_temperature
           "water_surface_area": hs["larval_habitat_area"], "aquatic": hs["aquatic"], "adult": hs["adult"],
           "monthly_means": np.stack([out["mean_state"][0, k] for k in range(2)] +
                                     [out["mean_forcing"][k] for k in range(4)], axis=1)}
    for v in ("mdr", "efd", "das_zerodensity", "ds"):
        got[v] = hf[v]
    got["daily"] = np.transpose(backend.daily_from_monthly(c["monthly"], c["mid_day"], c["day_predict"], math_mode=mode),
                                (1, 0, 2))
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fast", REL_TOL_FAST), ("strict", 1e-11)])
@pytest.mark.parametrize("species_dir", list(SPECIES))
def test_gpu_stages_on_the_committed_inputs(gpu, species_dir, mode, tol):
    """The CUDA path on the inputs R would get, every stage against the oracle (and, below, against R when present)."""
    c = case(SPECIES[species_dir])
    print(compare_stages(gpu_stages(c, mode), oracle_stages(c), tol))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "strict"])
@pytest.mark.parametrize("species_dir", list(SPECIES))
def test_gpu_matches_r(gpu, species_dir, mode):
    if not r_outputs_present():
        pytest.skip("no R outputs: run `Rscript baseline/run_reference.R --source <mosmicrosim checkout>`")
    ref = load_r_outputs(os.path.join(R_DIR, "outputs"), species_dir)
    print(compare_stages(gpu_stages(case(SPECIES[species_dir]), mode), ref, TOL))
