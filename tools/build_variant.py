"""Kernel experiments: builds variants of libmmsgpu.so that differ in the -D switches of ONE kernel translation unit
(default: csrc/chain_ws_tu.cu, the pair form of the fused kernel) and links each with the standard objects of
mosmicrosim_b200/build/.  The variants go to mosmicrosim_b200/variants/libmmsgpu_<name>.so (git-ignored; they travel
to the GPU box) and are selected at run time with MMS_LIB=<path> (mosmicrosim_b200/backend.py).

    python tools/build_variant.py name1="-DA=1 -DB=2" name2="-DA=0" [--unit=chain_ws_tu.cu:chain_ws_fast.o]
    python tools/build_variant.py --all name="-DMMS_CHAIN_MAX_WARPS=20"     # every nvcc unit with the flags
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mosmicrosim_b200 import build as b  # noqa: E402

unit_src, unit_obj = "chain_ws_tu.cu", "chain_ws_fast.o"
specs = []
build_all = False
for a in sys.argv[1:]:
    if a == "--all":
        build_all = True
    elif a.startswith("--unit"):
        unit_src, unit_obj = a.split("=", 1)[1].split(":")
    else:
        name, flags = a.split("=", 1)
        specs.append((name, flags.split()))
b.build()                                   # the standard objects
out_dir = os.path.join(b.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
units = [(s, o, e) for s, o, e in b.UNITS if (s.endswith(".cu") if build_all else o == unit_obj)]
procs = []
for name, flags in specs:
    for src, uobj, extra in units:
        obj = os.path.join(b.BUILD, f"variant_{name}_{uobj}")
        cmd = [b.NVCC] + b.COMMON + extra + flags + ["-Xptxas=-v", "-c", os.path.join(b.CSRC, src), "-o", obj]
        procs.append((name, uobj, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
failed = set()
for name, uobj, obj, p in procs:
    out, _ = p.communicate()
    open(os.path.join(b.BUILD, f"variant_{name}_{uobj}.log"), "w").write(out)
    if p.returncode != 0:
        print(name, uobj, "FAILED\n", out[-3000:])
        failed.add(name)
for name, flags in specs:
    if name in failed:
        continue
    mine = {uobj: obj for n, uobj, obj, _ in procs if n == name}
    lib = os.path.join(out_dir, f"libmmsgpu_{name}.so")
    objs = [mine.get(o, os.path.join(b.BUILD, o)) for _, o, _ in b.UNITS]
    r = subprocess.run([b.NVCC, "-shared", "-o", lib] + b.ARCH + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(name, "ok" if r.returncode == 0 else "LINK FAILED " + r.stdout, lib)
