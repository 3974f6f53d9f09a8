"""Ad-hoc timing of the fused kernel (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".."); sys.path.insert(0, ".")
from helpers import Case
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
habs = [syn.habitat_cone(), syn.habitat_cylinder(leak_per_hour=0.002)][:S] if S <= 2 else syn.habitat_ensemble(S)
case = Case(n_cells=C, habitats=habs, out_of_range_fraction=0.01)
print("fp64 peak TF", backend.measure_fp64_peak(0), "hbm GB/s", backend.measure_hbm_copy(0))
for mode in ("fast", "strict"):
    with backend.ChainPlan(C, **case.gpu_kwargs(mean_vars=list(backend.VAR_NAMES[:6]), want_wet_hours=True, math_mode=mode)) as plan:
        t = time.time(); plan.upload(case.monthly, case.altitude); plan.sync(); print("upload s", time.time() - t)
        for i in range(3):
            plan.launch(); plan.sync()
            ms = plan.kernel_ms()
            print(mode, "kernel ms", ms, "cell-years/s", C * case.H / 8760 / (ms * 1e-3))
