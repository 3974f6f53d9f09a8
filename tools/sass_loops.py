"""Innermost loops of one kernel's SASS (cuobjdump -sass -fun NAME file.o > x.sass; python tools/sass_loops.py x.sass):
instructions, FP64-pipe instructions, local (spill) loads / stores, barriers and shared-memory accesses of every loop
whose body holds no other backward branch -- the hour loops of the fused kernels' roles."""
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
back = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            back.append((addr_index[tgt], i))
for lo, hi in back:
    if any(l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in back):
        continue
    body = [t for _, t in ins[lo:hi + 1]]
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    ops = [op(t) for t in body]
    fp64 = sum(o.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")) or o.startswith("MUFU.RCP64") for o in ops)
    print(f"loop {ins[lo][0]:#07x}..{ins[hi][0]:#07x}: {len(body):4d} instr, fp64 {fp64:3d}, "
          f"LDL {sum(o.startswith('LDL') for o in ops)}, STL {sum(o.startswith('STL') for o in ops)}, "
          f"BAR {sum(o.startswith('BAR') for o in ops)}, LDS {sum(o.startswith('LDS') for o in ops)}, "
          f"STS {sum(o.startswith('STS') for o in ops)}, LDG {sum(o.startswith(('LDG', 'LD.')) for o in ops)}, "
          f"LDC {sum(o.startswith(('LDC', 'LDCU')) for o in ops)}, FSEL/SEL {sum(o.startswith(('FSEL', 'SEL')) for o in ops)}")
