"""Generates tests/golden/chain_small.npz from the numpy restatement of the
reference (oracle/restate_numpy.py) -- NOT from the C oracle, so that the golden
vectors pin the C oracle (and, through it, the GPU path) to an independently
written restatement.  R itself is unavailable, so there are no reference-run
vectors; see DESIGN.md "Oracle".

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import Case                      # noqa: E402
from mosmicrosim_b200 import synthetic as syn  # noqa: E402
from oracle import restate_numpy as rn         # noqa: E402

n_cells = 6
case = Case(n_cells=n_cells, habitats=[syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)])
dm = dict(type=case.densmod["type_"], dd_effect=case.densmod["dd_effect"],
          surface_area_cm2=case.densmod["surface_area_cm2"])
out = rn.run_chain(case.monthly, case.mid_day, case.day_predict, case.altitude, case.habitats, dm, case.tables,
                   case.period_of_day, case.P)
steps = np.unique(np.concatenate([np.arange(0, 48), np.arange(0, case.H, 97), np.arange(case.H - 48, case.H)]))
np.savez_compressed(os.path.join(HERE, "chain_small.npz"), n_cells=n_cells, monthly=case.monthly,
                    altitude=case.altitude, mean=out["mean"], wet_hours=out["wet_hours"], hourly_steps=steps,
                    hourly_sample=out["hourly"][:, :, steps, :])
print("wrote chain_small.npz", out["mean"].shape, out["hourly"][:, :, steps, :].shape)

# ---- TerraClimate front end and the one-pixel / matrix twins (second golden file; chain_small.npz is not rewritten
# unless this script is run, and it is deterministic) ------------------------------------------------------------
from mosmicrosim_b200 import dates  # noqa: E402

start, end = syn.month_table(2001, 1)
month_id, n_days = dates.terraclimate_month_info(start, end)
tc, clear_sky, _ = syn.synthetic_terraclimate(5, start)
tc[0, 2, 0], tc[1, 2, 0] = -3.0, -14.0                                  # the over-ice branch of VAPPRS
monthly_tc = rn.process_terraclimate_pixel_vars(tc, clear_sky, month_id, n_days)
h = out["hourly"][0]                                                    # cone habitat of the chain above
v = {name: h[k] for k, name in enumerate(rn.VAR_NAMES)}
n = 600                                                                 # a short series: whole-series burn-in twins
sub = {k: a[:n, 1] for k, a in v.items()}
hab = case.habitats[0]
w = rn.simulate_ephemeral_habitat_vectorised(*[np.tile(sub[k], 3)[:, None] for k in
                                               ("rainfall", "air_temperature", "humidity", "windspeed")],
                                             case.altitude[1:2], hab, 0)
p = rn.simulate_population_vectorised(*[np.tile(sub[k], 3)[:, None] for k in
                                        ("larval_habitat_area", "mdr", "efd", "das_zerodensity", "ds")],
                                      dm, 100.0, 100.0, 0)
np.savez_compressed(os.path.join(HERE, "front_end_and_twins.npz"), terraclimate=tc, clear_sky=clear_sky,
                    month_id=month_id, n_days=n_days, monthly=monthly_tc,
                    series=np.stack([sub[k] for k in rn.VAR_NAMES]),
                    altitude=case.altitude[1], water_surface_area_burnin2=w["water_surface_area"][-n:, 0],
                    aquatic_burnin2=p["aquatic"][-n:, 0], adult_burnin2=p["adult"][-n:, 0])
print("wrote front_end_and_twins.npz", monthly_tc.shape, n)
