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
