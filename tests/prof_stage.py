"""Timing driver for the stage-level twins (development aid): achieved HBM GB/s of the kernels."""
import sys
sys.path.insert(0, ".."); sys.path.insert(0, ".")
import numpy as np
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8760
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
rng = np.random.default_rng(1)
shape = (H, C)
mk = lambda lo, hi: np.asfortranarray(rng.uniform(lo, hi, size=shape))
rain, airt, rh, wind = mk(0, 0.4), mk(12, 36), mk(25, 100), mk(0.1, 5)
alt = rng.uniform(0, 2500, C)
for rep in range(3):
    w = backend.water_vectorised(rain, airt, rh, wind, alt, syn.habitat_cone(), math_mode=mode)
    ms = backend.stage_kernel_ms()
    print(f"water twin  {mode}: kernel {ms:8.3f} ms  {40.0 * H * C / ms / 1e6:8.1f} GB/s algorithmic (40 B/cell-hour)")
area = w["water_surface_area"]
mdr, efd, das, ds = mk(0, 0.16), mk(0, 10), mk(0.2, 0.95), mk(0.3, 0.9)
dm = backend.make_densmod("logit", syn.DD_EFFECT_AG, syn.MURIU_AREA_CM2)
for rep in range(3):
    backend.population_vectorised(area, mdr, efd, das, ds, dm, burnin_years=1 if H >= 8760 else 0, math_mode=mode)
    ms = backend.stage_kernel_ms()
    steps = H + (8760 if H >= 8760 else 0)
    print(f"population twin {mode}: kernel {ms:8.3f} ms  {56.0 * H * C / ms / 1e6:8.1f} GB/s algorithmic "
          f"(56 B/cell-hour over the {H} recorded steps; {steps} steps executed)")
