"""Builds libmmsgpu.so in-tree for sm_100a with nvcc (cross-compiles without a GPU).

    python -m mosmicrosim_b200.build [--force]

The kernel translation unit csrc/chain_tu.cu is compiled twice: "fast"
(FMA contraction on) and "strict" (-fmad=false, literal operation order);
csrc/chain_ws_tu.cu holds the warp-pair form of the fused kernel (fast only).
All units compile in parallel (about two minutes).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmmsgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + ARCH
COMMON += os.environ.get("MMS_NVCC_EXTRA", "").split()     # development aid: -D switches for kernel experiments

UNITS = [
    # (source, object, extra flags)
    # the monolithic form gains from conflict-free exp / log tables (100k cells x 1 setting: 37.6 -> 36.4 ms), the pair
    # and fan-out forms of chain_ws_tu.cu do not (49.6 -> 49.8, 89.8 -> 91.7 ms): mms_fastmath.cuh
    ("chain_tu.cu", "chain_fast.o", ["-DMMS_STRICT=0", "-DMMS_MATH_TABLE_COPIES=1"]),
    ("chain_ws_tu.cu", "chain_ws_fast.o", ["-DMMS_STRICT=0"]),
    ("chain_team_tu.cu", "chain_team_fast.o", ["-DMMS_STRICT=0"]),
    ("chain_tu.cu", "chain_strict.o", ["-DMMS_STRICT=1", "-fmad=false"]),
    ("mmsgpu_api.cu", "mmsgpu_api.o", []),
    ("api_common.cu", "api_common.o", []),
    ("api_twins.cu", "api_twins.o", []),
    ("host_pipeline.cu", "host_pipeline.o", []),
    ("host_copy.cpp", "host_copy.o", []),       # g++ only (AVX2 streaming stores behind a run-time check)
    ("measure.cu", "measure.o", []),
]


def _sources():
    out = [os.path.join(HERE, "..", "include", "mmsgpu.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    procs = []
    headers = [os.path.join(HERE, "..", "include", "mmsgpu.h")]
    headers += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    newest_header = max(os.path.getmtime(h) for h in headers)
    flags_tag = " ".join(COMMON)
    for src, obj, extra in UNITS:
        o = os.path.join(BUILD, obj)
        tag = o + ".flags"
        # a unit is rebuilt when its source, any header or the flags changed
        if (not force and os.path.exists(o) and os.path.exists(tag) and open(tag).read() == flags_tag + " ".join(extra)
                and os.path.getmtime(o) > max(newest_header, os.path.getmtime(os.path.join(CSRC, src)))):
            continue
        with open(tag, "w") as f:
            f.write(flags_tag + " ".join(extra))
        cmd = [NVCC] + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", o]
        if src.endswith(".cpp"):
            cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O3", "-fPIC", "-fvisibility=hidden", "-c",
                   os.path.join(CSRC, src), "-o", o]
        if verbose:
            if not src.endswith(".cpp"):
                cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    link = [NVCC, "-shared", "-o", LIB] + ARCH + [os.path.join(BUILD, o) for _, o, _ in UNITS] + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
