"""Runs one of BASELINE.json's larger configurations (one GPU's shard of it) through the C ABI
and prints the timings (development / evidence aid; bench.py is the contract)."""
import sys, time, json
sys.path.insert(0, ".."); sys.path.insert(0, ".")
import numpy as np
from mosmicrosim_b200 import backend, dates, synthetic as syn

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == "cfg3":      # Africa 1/24 deg: ~1.5 M land cells, 10 years hourly, monthly summaries
    total_cells, years, habitats, agg, names = 1_500_000, 10, [syn.habitat_cone()], "month", list(backend.VAR_NAMES[:6])
elif which == "cfg5":    # 64 water-body / leakage settings x 250k cells x 1 year
    total_cells, years, habitats, agg, names = 250_000, 1, syn.habitat_ensemble(64), "month", ["adult", "aquatic"]
elif which == "cfg4":    # global land 1/24 deg: ~10 M cells, 5 years hourly, DAILY summaries only (two variables)
    total_cells, years, habitats, agg, names = 10_000_000, 5, [syn.habitat_cone()], "day", ["adult", "aquatic"]
else:
    raise SystemExit("cfg3 | cfg4 | cfg5")
C = total_cells // gpus
lh = syn.synthetic_lifehistory("An. gambiae")
start, end = syn.month_table(2001, years)
dp = syn.year_dates(2001, years)
mid_day, day_predict = dates.datenums(start, end, dp)
monthly, altitude = syn.synthetic_forcing(C, start)
pod, ps, _ = dates.period_index(dp, agg)
dm = lh.das_densmod
t0 = time.perf_counter()
plan = backend.ChainPlan(C, mid_day=mid_day, day_predict=day_predict, habitats=habitats,
                         densmod=dict(type_=dm.type, dd_effect=dm.dd_effect, surface_area_cm2=dm.surface_area_cm2),
                         tables=lh.tables(), period_of_day=pod, n_periods=len(ps), burnin_years_pop=1,
                         mean_vars=names, want_wet_hours=True, devices=[0])
out = plan.allocate_outputs()
t1 = time.perf_counter()
plan.upload(monthly, altitude); plan.launch(); plan.download(out)
t2 = time.perf_counter()
plan.launch(); plan.sync()
ms = plan.kernel_ms()
H = 24 * len(dp)
print(json.dumps({"config": which, "shard": f"1/{gpus}", "cells": C, "hours": H, "settings": len(habitats),
                  "kernel_ms": ms, "e2e_s_pageable_host": t2 - t1, "create_s": t1 - t0,
                  "cell_years_per_s_kernel": C * len(habitats) * (H / 8760) / (ms * 1e-3),
                  "adult_mean_finite_fraction": float(np.isfinite(out["mean_state"][0, 0]).mean())}))
plan.close()
