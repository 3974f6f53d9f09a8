"""Profiling driver (development aid): one upload, N launches of the fused kernel."""
import sys
sys.path.insert(0, ".."); sys.path.insert(0, ".")
from helpers import Case
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
years = int(sys.argv[5]) if len(sys.argv) > 5 else 1
habs = [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)][:S] if S <= 2 else syn.habitat_ensemble(S)
case = Case(n_cells=C, habitats=habs, out_of_range_fraction=0.01, years=years)
with backend.ChainPlan(C, **case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode=mode)) as plan:
    plan.upload(case.monthly, case.altitude)
    for i in range(n):
        plan.launch(); plan.sync()
        print(mode, "kernel ms", plan.kernel_ms())
