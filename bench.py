#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend for the mosmicrosim chain.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-configs]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): cell-years / s, hourly steps, fp64.  Workload at N = 1 is
BASELINE.json configs[1]: a synthetic 100k-cell tile, 1 simulated year at hourly
steps, two water bodies per cell (the reference's cone and a cylinder with
leakage), monthly summaries of the reference's six vector variables.  With N > 1
every rank (one process per GPU) owns its own 100k-cell tile -- cells never
communicate, so there is no collective on the data path ("scaling": "weak").

A step = one pass of the fused chain over the tile: 8760 burn-in hours (executed,
not credited) + 8760 recorded hours per cell.  `value` is timed with CUDA events
on the library's own stream with the inputs resident in HBM; `e2e` is ONE C-ABI
call per step (mms_chain_run: upload, kernels and download pipelined in chunks)
on pinned HOST buffers, and `e2e.pageable` the same call on ordinary numpy
memory (what an R caller hands over).  `--impl reference` times the CPU
restatement of the reference (oracle/, all host threads) on a bounded sample of
the same workload.

`configs` (rank 0, after the per-rank run): BASELINE.json configs[2], [3] and [4]
at FULL size through ONE multi-device plan of the library over the N GPUs of the
launch (mms_chain_create(devices = 0..N-1), one host-thread team per device, one
gather into one host buffer) -- fixed total work, so seconds_e2e at N = 1, 2, 4, 8
is the strong-scaling curve of the north-star target.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mosmicrosim_b200 import dates, synthetic as syn  # noqa: E402

N_CELLS = 100_000
YEARS = 1
FIRST_YEAR = 2001
CPU_SAMPLE_CELLS = 40_000      # ~12 s of host work on 16 cores
REFERENCE_ARM_BUDGET_S = 100.0  # --impl reference: all its steps together

# ---- algorithmic work per cell-hour (DESIGN.md "Roofline accounting") ----
# arithmetic counted off the R expressions, transcendental weights exp = 28, log = 36, pow = 130
FLOP_SHARED = 309          # microclimate 15, evaporation 19 + exp, splines 27, bilinear 22,
                           # emergence exp + 2, ds^(1/24) pow, logit(das0) log + 2
FLOP_CONE = 321            # cone step 11 + pow, densmod 8 + exp, das^(1/24) pow, population 14
FLOP_CYLINDER = 191        # as the cone without V^0.666667

# ncu summary of the bench kernel (one launch of this workload), committed with the kernel it describes
NCU_PROFILES = ("profiles/r02_ncu_chain_ws.txt", "profiles/r01_ncu_chain_ws.txt")


def workload(rank: int):
    lh = syn.synthetic_lifehistory("An. gambiae")
    start, end = syn.month_table(FIRST_YEAR, YEARS)
    dates_predict = syn.year_dates(FIRST_YEAR, YEARS)
    mid_day, day_predict = dates.datenums(start, end, dates_predict)
    monthly, altitude = syn.synthetic_forcing(N_CELLS, start, seed=syn.SEED + rank)
    period_of_day, p_start, _ = dates.period_index(dates_predict, "month")
    habitats = [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)]
    dm = lh.das_densmod
    return dict(lh=lh, mid_day=mid_day, day_predict=day_predict, monthly=monthly, altitude=altitude,
                period_of_day=period_of_day, n_periods=len(p_start), habitats=habitats,
                densmod=dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2),
                D=len(dates_predict))


CONFIG = {
    "workload": "configs[1]: synthetic 100k-cell tile, 1 year hourly (8760 steps + 8760 burn-in steps "
                "executed, not credited), cone + cylinder-with-leakage per cell, monthly means of 6 variables",
    "cells_per_gpu": N_CELLS, "years": YEARS, "hours": 8760 * YEARS, "habitat_settings": 2,
    "burnin_years_pop": 1, "species": "An. gambiae (logit density dependence)",
    "tables": "synthetic (reference sysdata.rda unavailable)", "math_mode": "fast",
    "l2": "flushed between timed iterations (256 MB memset, outside the timed events)",
    "sharding": "one 100k-cell tile per GPU, no collective",
}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        busy = sorted(sm)[len(sm) // 2:] if sm else []      # upper half = samples taken under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads() -> int:
    # every host core this process may use, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(w, n_cells: int, steps: int = 1, warmup: int = 0):
    """The reference's algorithm on the host cores (oracle port, OpenMP over cells)."""
    import oracle
    tabs = w["lh"].tables()
    hab = [oracle.habitat(**h) for h in w["habitats"]]
    dm = oracle.densmod(w["densmod"]["type_"], w["densmod"]["dd_effect"], w["densmod"]["surface_area_cm2"])
    monthly = np.ascontiguousarray(w["monthly"][:, :, :n_cells])
    alt = w["altitude"][:n_cells]
    n_threads = host_threads()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.run_chain(monthly, w["mid_day"], w["day_predict"], alt, hab, dm, tabs, w["period_of_day"],
                         w["n_periods"], burnin_years_water=0, burnin_years_pop=1, want_hourly=False,
                         n_threads=n_threads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = n_cells * YEARS / sec
    return {"value": value, "unit": "cell-years/s", "cores": n_threads, "value_per_core": value / n_threads,
            "kind": "port", "seconds_per_step": sec, "sample_cells": int(n_cells),
            "sample": f"first {n_cells} cells of the same tile, same settings and outputs, "
                      f"{n_threads} OpenMP threads over cells (R itself is not installed on this image)"}


def reference_arm(args):
    """--impl reference: exactly --steps timed and --warmup untimed steps of the CPU restatement, each on a
    sample of the tile sized (from a short calibration run) so that all of them end within ~100 s."""
    w = workload(0)
    cal_cells = 2_000
    cal = cpu_reference(w, cal_cells, steps=1, warmup=0)
    per_cell = cal["seconds_per_step"] / cal_cells
    n_steps = max(1, args.steps) + max(0, args.warmup)
    sample = int(REFERENCE_ARM_BUDGET_S / n_steps / per_cell)
    sample = max(2_000, min(CPU_SAMPLE_CELLS, sample // 1000 * 1000))
    cb = cpu_reference(w, sample, steps=max(1, args.steps), warmup=max(0, args.warmup))
    keys = ("value", "unit", "cores", "value_per_core", "kind", "sample", "sample_cells")
    return {"impl": "reference", "metric": "cell-years/sec (hourly steps)", "value": cb["value"],
            "unit": "cell-years/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {k: cb[k] for k in keys},
            "e2e": {"value": cb["value"], "unit": "cell-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def ncu_profile():
    """dram bytes / FP64-pipe activity of one launch of the bench kernel, read from the committed ncu summary."""
    for rel in NCU_PROFILES:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        raw = open(path, "rb").read()
        vals = {}
        for line in raw.decode("utf-8", "replace").splitlines():
            m = re.match(r"^(\S+)\s+(\S+)?\s+([-+0-9.eE,]+)\s*$", line)
            if m and not line.startswith("#"):
                try:
                    vals[m.group(1)] = (m.group(2) or "", float(m.group(3).replace(",", "")))
                except ValueError:
                    pass

        def in_bytes(key):
            if key not in vals:
                return None
            unit, v = vals[key]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

        rd, wr = in_bytes("dram__bytes_read.sum"), in_bytes("dram__bytes_write.sum")
        dur = vals.get("gpu__time_duration.sum")
        dur_ms = None if dur is None else dur[1] * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(dur[0], 1)
        pct = vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        return {"path": rel, "sha256_16": hashlib.sha256(raw).hexdigest()[:16],
                "traffic": None if rd is None or wr is None else rd + wr,
                "duration_ms": dur_ms, "fp64_pipe_pct": None if pct is None else pct[1]}
    return None


# ---- BASELINE.json configs[2..4] at full size through one multi-device plan ---------------------------------

def big_configs(n_dev: int, which=("cfg3", "cfg5", "cfg4")):
    from mosmicrosim_b200 import backend
    lh = syn.synthetic_lifehistory("An. gambiae")
    dm = lh.das_densmod
    densmod = dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2)
    specs = {
        # Africa 1/24 deg: ~1.5 M land cells, 10 years hourly (+ 1 burn-in year), monthly means of the 6 variables
        "cfg3": dict(baseline="configs[2]", cells=1_500_000, tile=1_500_000, years=10, habitats=[syn.habitat_cone()],
                     agg="month", names=list(backend.VAR_NAMES[:6]), wet=True),
        # global land 1/24 deg: ~10 M cells, 5 years hourly, DAILY means of adult and aquatic only; the raster is
        # walked in 8 tiles of 1.25 M cells through one plan (292 GB of daily summaries do not fit one host buffer)
        "cfg4": dict(baseline="configs[3]", cells=10_000_000, tile=1_250_000, years=5, habitats=[syn.habitat_cone()],
                     agg="day", names=["adult", "aquatic"], wet=False),
        # 64 water-body / leakage settings x 250k cells x 1 year
        "cfg5": dict(baseline="configs[4]", cells=250_000, tile=250_000, years=1, habitats=syn.habitat_ensemble(64),
                     agg="month", names=["adult", "aquatic"], wet=True),
    }
    results = {}
    # the other ranks of a torchrun launch are idle by now: this process may use every host core
    os.environ["MMS_HOST_WORKERS"] = str(max(1, min(12, host_threads() // max(1, n_dev))))
    for name in which:
        sp = specs[name]
        try:
            t_gen = time.perf_counter()
            start, end = syn.month_table(FIRST_YEAR, sp["years"])
            dp = syn.year_dates(FIRST_YEAR, sp["years"])
            mid_day, day_predict = dates.datenums(start, end, dp)
            pod, ps, _ = dates.period_index(dp, sp["agg"])
            C = sp["tile"]
            monthly, altitude = syn.synthetic_forcing_large(C, start, seed=syn.SEED + 7)
            t_gen = time.perf_counter() - t_gen
            t0 = time.perf_counter()
            plan = backend.ChainPlan(C, mid_day=mid_day, day_predict=day_predict, habitats=sp["habitats"],
                                     densmod=densmod, tables=lh.tables(), period_of_day=pod, n_periods=len(ps),
                                     burnin_years_water=0, burnin_years_pop=1, mean_vars=sp["names"],
                                     want_wet_hours=sp["wet"], math_mode="fast", devices=list(range(n_dev)))
            create_s = time.perf_counter() - t0
            out = plan.allocate_outputs()                      # ordinary (pageable) numpy memory
            n_tiles = sp["cells"] // C
            h2d = monthly.nbytes + altitude.nbytes
            d2h = sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray))
            plan.run(monthly, altitude, out)                   # warm-up: first touch of the output pages, bounce buffers
            secs, kms = [], []
            for _ in range(1 if n_tiles > 1 else 2):
                t0 = time.perf_counter()
                for _t in range(n_tiles):
                    plan.run(monthly, altitude, out)
                secs.append(time.perf_counter() - t0)
                kms.append(plan.kernel_ms() * n_tiles)
            sec = float(np.mean(secs))
            # the same call on buffers from mms_host_alloc (page-locked: the DMA runs on them in place, no bounce copy)
            pinned = None
            if os.environ.get("MMS_BENCH_PINNED_CONFIGS", "1") != "0":
                try:
                    t0 = time.perf_counter()
                    pm, pa = backend.PinnedArray(monthly.shape), backend.PinnedArray(altitude.shape)
                    pm.array[...] = monthly
                    pa.array[...] = altitude
                    pout = plan.allocate_outputs(pinned=True)
                    t_alloc = time.perf_counter() - t0
                    plan.run(pm.array, pa.array, pout)
                    t0 = time.perf_counter()
                    for _t in range(n_tiles):
                        plan.run(pm.array, pa.array, pout)
                    pinned = {"seconds_e2e": time.perf_counter() - t0, "allocation_and_fill_s": t_alloc}
                    del pm, pa, pout
                except Exception as e:
                    pinned = {"error": f"{type(e).__name__}: {e}"}
            S = len(sp["habitats"])
            H = 24 * len(dp)
            info = plan.info()
            finite = float(np.isfinite(out["mean_state"][0, 0]).mean())
            plan.close()
            results[name] = {
                "baseline_config": sp["baseline"], "cells": sp["cells"], "years": sp["years"], "hours": H,
                "settings": S, "summaries": f"{sp['agg']} means of {len(sp['names'])} variables"
                                            + (" + wet-hour counts" if sp["wet"] else ""),
                "n_devices": n_dev, "tiles": n_tiles, "seconds_e2e": sec,
                "kernel_seconds_max_over_devices": float(np.mean(kms)) * 1e-3,
                "cell_years_per_s_e2e": sp["cells"] * (H / 8760.0) / sec,
                "setting_cell_years_per_s_e2e": sp["cells"] * S * (H / 8760.0) / sec,
                "h2d_bytes": int(h2d * n_tiles), "d2h_bytes": int(d2h * n_tiles),
                "host_memory": "pageable (numpy) in and out, staged through the plan's pinned bounce buffers",
                "pinned_host_memory": pinned,
                "host_workers_per_device": int(os.environ["MMS_HOST_WORKERS"]),
                "plan": info, "plan_create_s": create_s, "input_generation_s": t_gen,
                "adult_mean_finite_fraction": finite,
                "path": "one mms_chain_create(devices = 0..N-1) plan, mms_chain_run per tile, one host output buffer",
            }
            del monthly, altitude, out
        except Exception as e:                                  # the headline line must still be printed
            results[name] = {"error": f"{type(e).__name__}: {e}"}
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the full-size configs[2..4] runs")
    ap.add_argument("--configs", default="cfg3,cfg5,cfg4")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(args)))
        return 0

    from mosmicrosim_b200 import backend
    dist = None
    cpu_group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl")
        try:
            cpu_group = dist.new_group(backend="gloo")  # host-side waiting (an NCCL barrier spins on the GPUs)
        except Exception:
            cpu_group = None

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if backend.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (libmmsgpu has no CPU fallback)")
    w = workload(rank)
    dev = local_rank
    # host threads of this rank's plan: the box's cores are shared by the ranks (the library reads LOCAL_WORLD_SIZE
    # too; stated here so that the line records it)
    host_workers = max(1, min(12, host_threads() // max(1, world)))
    os.environ["MMS_HOST_WORKERS"] = str(host_workers)
    peak_fp64 = backend.measure_fp64_peak(dev)          # TFLOP/s, DFMA loop, this GPU, now
    hbm_copy = backend.measure_hbm_copy(dev)
    dinfo = backend.device_info(dev)

    plan = backend.ChainPlan(N_CELLS, mid_day=w["mid_day"], day_predict=w["day_predict"], habitats=w["habitats"],
                             densmod=w["densmod"], tables=w["lh"].tables(), period_of_day=w["period_of_day"],
                             n_periods=w["n_periods"], burnin_years_water=0, burnin_years_pop=1,
                             mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode="fast",
                             devices=[dev])
    # pinned host buffers for the end-to-end path
    pin_m = backend.PinnedArray(w["monthly"].shape)
    pin_m.array[...] = w["monthly"]
    pin_a = backend.PinnedArray(w["altitude"].shape)
    pin_a.array[...] = w["altitude"]
    out = plan.allocate_outputs(pinned=True)
    out_pageable = plan.allocate_outputs()
    h2d = pin_m.nbytes + pin_a.nbytes
    d2h = sum(v.nbytes for k, v in out.items() if isinstance(v, np.ndarray))

    # ---- kernel-only arm: inputs resident in HBM ------------------------------------------------
    plan.upload(pin_m.array, pin_a.array)
    plan.sync()
    for _ in range(args.warmup):
        plan.launch()
        plan.sync()
    launches0 = plan.launch_count()
    sampler = ClockSampler(dev)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        backend.flush_l2(dev)                      # untimed; the events below bracket the kernel only
        plan.launch()
        plan.sync()
        kernel_ms.append(plan.kernel_ms())
    barrier()
    wall = time.perf_counter() - wall0
    gpu_launches = plan.launch_count() - launches0
    total_ms = max_over_ranks(float(np.sum(kernel_ms)))
    ms_per_step = total_ms / args.steps
    cell_years = N_CELLS * YEARS * world
    value = cell_years / (ms_per_step * 1e-3)

    # ---- end-to-end arm: one pipelined C-ABI call per step, host buffers in and out ----------------
    def e2e_arm(m, a, o):
        for _ in range(min(args.warmup, 2)):
            plan.run(m, a, o)
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            plan.run(m, a, o)                      # returns when the results are in the host buffers
        barrier()
        return max_over_ranks((time.perf_counter() - e0) / args.steps)

    e2e_s = e2e_arm(pin_m.array, pin_a.array, out)
    e2e_kernel_ms = plan.kernel_ms()
    e2e_pageable_s = e2e_arm(w["monthly"], w["altitude"], out_pageable)
    clocks = sampler.stop()

    # ---- roofline of the fused kernel (FP64-pipe bound: ~800 flop per 0.2 algorithmic bytes) -------
    steps_per_cell = 8760 * (1 + 1) * YEARS          # burn-in year + recorded year, all executed
    flop_per_launch = float(N_CELLS) * steps_per_cell * (FLOP_SHARED + FLOP_CONE + FLOP_CYLINDER)
    own_ms = float(np.mean(kernel_ms))
    achieved = flop_per_launch / (own_ms * 1e-3) / 1e12
    alg_bytes = float(h2d + d2h)
    sm_mhz = clocks.get("sm_mhz") or dinfo["sm_clock_mhz_max"]
    peak_nominal = dinfo["n_sm"] * 64 * 2 * sm_mhz * 1e6 / 1e12      # 64 FP64 FMA / clk / SM at the clock under load
    prof = ncu_profile()
    fp64_pipe_frac = None
    if prof and prof["fp64_pipe_pct"] is not None and prof["duration_ms"]:
        # the FP64 instructions a launch executes do not depend on the run: the share of FP64-pipe cycles they
        # fill scales with the profiled launch's duration over this run's
        fp64_pipe_frac = prof["fp64_pipe_pct"] / 100.0 * prof["duration_ms"] / own_ms
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s",
                "frac": achieved / peak_fp64,
                "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal,
                "peak_nominal_formula": f"{dinfo['n_sm']} SMs x 64 FMA/clk x 2 x {sm_mhz:.0f} MHz",
                "fp64_pipe_frac": fp64_pipe_frac,
                "fp64_pipe_frac_note": "executed FP64 warp instructions / FP64 issue slots (ncu sm__pipe_fp64_cycles_active "
                                       "of the committed capture, rescaled by its launch duration over this run's); the "
                                       "convention `frac` prices pow / exp / log at libm cost (130 / 28 / 36 flop)",
                "traffic": prof["traffic"] if prof else None,
                "traffic_source": {k: prof[k] for k in ("path", "sha256_16")} if prof else None,
                "traffic_unit": "dram bytes read + written per launch (ncu --set full capture named in traffic_source)",
                "peak_source": "measured live: dependent-free DFMA loop on this GPU (mms_measure_fp64_peak); "
                               "FP64 is not in MEASURED_PEAKS.json",
                "flop_per_cell_hour": FLOP_SHARED + FLOP_CONE + FLOP_CYLINDER,
                "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (own_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm_copy, "note": "forcing in + summaries out; the kernel is not HBM bound"}}

    line = {"metric": "cell-years/sec (hourly steps)", "value": value, "unit": "cell-years/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG,
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(gpu_launches),
            "wall_ms_per_step_incl_l2_flush": wall / args.steps * 1e3,
            "e2e": {"value": cell_years / e2e_s, "unit": "cell-years/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "kernel_ms_inside": e2e_kernel_ms, "plan": plan.info(),
                    "path": "mms_chain_run (chunked: H2D of chunk k+1 / kernel of chunk k / D2H of chunk k-1 overlap) "
                            "on pinned host buffers",
                    "pageable": {"value": cell_years / e2e_pageable_s, "ms_per_step": e2e_pageable_s * 1e3,
                                 "path": "the same call on ordinary numpy memory (pinned bounce buffers inside the library)"}}}
    if rank == 0 and world == 1:
        cb = cpu_reference(w, CPU_SAMPLE_CELLS)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "value_per_core", "kind", "sample")}
    plan.close()
    del pin_m, pin_a, out, out_pageable
    os.environ.pop("MMS_HOST_WORKERS", None)      # the multi-device plans below size their own teams

    # ---- the north-star configurations at full size, one multi-device plan over the N GPUs (rank 0) ----
    want_configs = not args.no_configs and os.environ.get("MMS_BENCH_CONFIGS", "1") != "0"
    if want_configs:
        if dist is not None:
            dist.barrier(group=cpu_group)          # every rank has released its GPU
        if rank == 0:
            n_dev = min(world, backend.device_count())
            line["configs"] = big_configs(n_dev, tuple(c for c in args.configs.split(",") if c))
        if dist is not None:
            dist.barrier(group=cpu_group)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
