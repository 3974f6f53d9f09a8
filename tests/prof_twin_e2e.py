"""End-to-end time of the stage-level twins (development / evidence aid): 100k pixels x 8760 hours of R-layout
matrices from pageable and from pinned host memory, against the kernel time and the PCIe time of the same bytes."""
import sys, time, json
sys.path.insert(0, ".."); sys.path.insert(0, ".")
import numpy as np
from mosmicrosim_b200 import backend, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
H = 8760
rng = np.random.default_rng(5)
base = 2048                                   # a block of distinct pixels, repeated


def mat(lo, hi, pinned):
    blk = rng.uniform(lo, hi, size=(base, H))
    if pinned:
        pa = backend.PinnedArray((C, H))
        a = pa.array
    else:
        pa, a = None, np.empty((C, H))
    for c0 in range(0, C, base):
        n = min(base, C - c0)
        a[c0:c0 + n] = blk[:n]
    return pa, a.T                            # [H, C] column-major view = R's matrix


res = {"pixels": C, "hours": H}
for pinned in (False, True):
    keep = [mat(0, 0.4, pinned), mat(12, 36, pinned), mat(25, 100, pinned), mat(0.1, 5, pinned)]
    alt = rng.uniform(0, 2500, C)
    hab = backend.make_habitat(**syn.habitat_cone())
    import ctypes as ct
    area = backend.PinnedArray((C, H)) if pinned else None
    out = area.array.T if pinned else np.empty((H, C), order="F")
    L = backend.lib()
    dp = backend._dp

    def call():
        backend.check(L.mms_water_vectorised(*[m[1].ctypes.data_as(dp) for m in keep], alt.ctypes.data_as(dp), H, C,
                                             ct.byref(hab), 0, 0, 0, out.ctypes.data_as(dp), None, None))

    call()                                    # warm-up: first touch of the output pages, bounce buffers
    t0 = time.perf_counter(); call(); t1 = time.perf_counter()
    key = "pinned" if pinned else "pageable"
    res[key + "_e2e_s"] = t1 - t0
    res[key + "_kernel_ms"] = backend.stage_kernel_ms()
    res["bytes_in"] = 4 * C * H * 8 + C * 8
    res["bytes_out"] = C * H * 8
    del keep, out, area
print(json.dumps(res))
