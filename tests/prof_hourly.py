"""Timing driver (development aid): the fused kernel writing full hourly series (56 B per cell-hour
for the seven reference variables) next to the same launch with monthly means only."""
import sys
sys.path.insert(0, ".."); sys.path.insert(0, ".")
from helpers import Case
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
names = ["adult", "aquatic", "larval_habitat_area", "ds", "mdr", "efd", "das_zerodensity"]   # R/pipeline_functions.R:879-891
case = Case(n_cells=C, habitats=[syn.habitat_cone()], out_of_range_fraction=0.01)
for label, kw in (("monthly means only", dict(mean_vars=names[:2] + names[3:])),
                  ("full hourly series ", dict(hourly_vars=names))):
    with backend.ChainPlan(C, **case.gpu_kwargs(math_mode=mode, **kw)) as plan:
        plan.upload(case.monthly, case.altitude)
        for i in range(3):
            plan.launch(); plan.sync()
        ms = plan.kernel_ms()
        out_bytes = 56.0 * case.H * C if "hourly_vars" in kw else 0.0
        print(f"{label}: kernel {ms:8.3f} ms   {C * case.H / ms / 1e6:8.1f} G cell-hours/s"
              + (f"   {out_bytes / ms / 1e6:8.1f} GB/s of hourly output (56 B per cell-hour)" if out_bytes else ""))
