#!/usr/bin/env python
"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with
nvdisasm -g line info of the same kernel and aggregate executed warp
instructions by CUDA source line and by opcode.

usage: sass_profile.py src.csv sass_all.txt <mangled kernel name> [n_warp_units]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass_txt, kname = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else None

# --- ncu rows in address order
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
ci = {k: hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples")}
ncu = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    ncu.append((r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])))

# --- nvdisasm listing of that kernel: instructions with the last seen line annotation
lines = open(sass_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
ins = []
cur = ("?", 0)
inl = ""
for l in lines[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
if len(ins) != len(ncu):
    print(f"warning: {len(ins)} disassembled vs {len(ncu)} profiled instructions", file=sys.stderr)

by_line = defaultdict(lambda: [0, 0])
by_op = defaultdict(int)
tot = 0
for (txt, loc), (ntxt, n, smp) in zip(ins, ncu):
    by_line[loc][0] += n
    by_line[loc][1] += smp
    op = ntxt.split()[0] if not ntxt.startswith("@") else ntxt.split()[1]
    by_op[op.split(".")[0]] += n
    tot += n
scale = (1.0 / units) if units else None
print(f"total warp instructions {tot}" + (f"  = {tot * scale:.1f} per unit" if scale else ""))
print("--- by opcode")
for op, n in sorted(by_op.items(), key=lambda kv: -kv[1])[:28]:
    print(f"{n / tot * 100:5.1f}%  {n * scale if scale else n:10.1f}  {op}")
print("--- by source line")
tots = sum(v[1] for v in by_line.values())
for loc, (n, smp) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{n / tot * 100:5.1f}%  {n * scale if scale else n:10.1f}  samples {smp / max(tots,1) * 100:4.1f}%  {loc[0]}:{loc[1]}")
