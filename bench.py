#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend for the mosmicrosim chain.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): cell-years / s, hourly steps, fp64.  Workload at N = 1 is
BASELINE.json configs[1]: a synthetic 100k-cell tile, 1 simulated year at hourly
steps, two water bodies per cell (the reference's cone and a cylinder with
leakage), monthly summaries of the reference's six vector variables.  With N > 1
every rank (one process per GPU) owns its own 100k-cell tile -- cells never
communicate, so there is no collective on the data path ("scaling": "weak").

A step = one pass of the fused chain over the tile: 8760 burn-in hours (executed,
not credited) + 8760 recorded hours per cell.  `value` is timed with CUDA events
on the library's own stream with the inputs resident in HBM; `e2e` goes through
the C-ABI plan calls with pinned HOST buffers (H2D forcing, kernel, D2H summaries
inside the timed region).  `--impl reference` times the CPU restatement of the
reference (oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
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
CPU_SAMPLE_CELLS = 40_000      # ~15 s of host work on 16 cores

# ---- algorithmic work per cell-hour (DESIGN.md "Roofline accounting") ----
# arithmetic counted off the R expressions, transcendental weights exp = 28, log = 36, pow = 130
FLOP_SHARED = 309          # microclimate 15, evaporation 19 + exp, splines 27, bilinear 22,
                           # emergence exp + 2, ds^(1/24) pow, logit(das0) log + 2
FLOP_CONE = 321            # cone step 11 + pow, densmod 8 + exp, das^(1/24) pow, population 14
FLOP_CYLINDER = 191        # as the cone without V^0.666667


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


def cpu_reference(w, n_cells: int, steps: int = 1, warmup: int = 0):
    """The reference's algorithm on the host cores (oracle port, OpenMP over cells)."""
    import oracle
    tabs = w["lh"].tables()
    hab = [oracle.habitat(**h) for h in w["habitats"]]
    dm = oracle.densmod(w["densmod"]["type_"], w["densmod"]["dd_effect"], w["densmod"]["surface_area_cm2"])
    monthly = np.ascontiguousarray(w["monthly"][:, :, :n_cells])
    alt = w["altitude"][:n_cells]
    # every host core this process may use, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    try:
        n_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        n_threads = os.cpu_count() or 1
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.run_chain(monthly, w["mid_day"], w["day_predict"], alt, hab, dm, tabs, w["period_of_day"],
                         w["n_periods"], burnin_years_water=0, burnin_years_pop=1, want_hourly=False,
                         n_threads=n_threads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": n_cells * YEARS / sec, "unit": "cell-years/s", "cores": n_threads,
            "kind": "port", "seconds_per_step": sec,
            "sample": f"first {n_cells} cells of the same tile, same settings and outputs, "
                      f"{n_threads} OpenMP threads over cells (R itself is not installed on this image)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        w = workload(0)
        timed, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)     # bounded: each step is 10-30 s of CPU work
        cb = cpu_reference(w, CPU_SAMPLE_CELLS, steps=timed, warmup=warm)
        line = {"impl": "reference", "metric": "cell-years/sec (hourly steps)", "value": cb["value"],
                "unit": "cell-years/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps, "warmup": args.warmup,
                "timed_steps": timed, "untimed_steps": warm,
                "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "cell-years/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    from mosmicrosim_b200 import backend
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl")

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
    peak_fp64 = backend.measure_fp64_peak(dev)          # TFLOP/s, DFMA loop, this GPU, now
    hbm_copy = backend.measure_hbm_copy(dev)

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

    # ---- end-to-end arm: pinned host -> device -> kernel -> host, every step ----------------------
    for _ in range(min(args.warmup, 2)):
        plan.upload(pin_m.array, pin_a.array)
        plan.launch()
        plan.download(out)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        plan.upload(pin_m.array, pin_a.array)
        plan.launch()
        plan.download(out)                         # synchronises
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - e0) / args.steps)
    clocks = sampler.stop()

    # ---- roofline of the fused kernel (FP64-pipe bound: ~800 flop per 0.2 algorithmic bytes) -------
    steps_per_cell = 8760 * (1 + 1) * YEARS          # burn-in year + recorded year, all executed
    flop_per_launch = float(N_CELLS) * steps_per_cell * (FLOP_SHARED + FLOP_CONE + FLOP_CYLINDER)
    own_ms = float(np.mean(kernel_ms))
    achieved = flop_per_launch / (own_ms * 1e-3) / 1e12
    alg_bytes = float(h2d + d2h)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s",
                "frac": achieved / peak_fp64,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one launch,
                # from the committed capture profiles/r01_ncu_chain_ws.txt (91.4 MB + 54.9 MB)
                "traffic": 147.5e6, "traffic_unit": "bytes per launch (ncu, profiles/r01_ncu_chain_ws.txt)",
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
                    "path": "mms_chain_upload + mms_chain_launch + mms_chain_download on pinned host buffers"}}
    if rank == 0 and world == 1:
        cb = cpu_reference(w, CPU_SAMPLE_CELLS)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    plan.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
