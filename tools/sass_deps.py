"""Producer distance of the FP64 operations in the big loops of a kernel's SASS: for every DFMA / DMUL / DADD, how many
instructions back the most recent producer of one of its sources sits (1 = the instruction just before: a stall of the
full FP64 latency unless another warp fills in).  cuobjdump -sass -fun NAME file.o > x.sass; python tools/sass_deps.py x.sass [min_len]"""
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 150
idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tg = int(m.group(1), 16)
        if tg <= a and tg in idx:
            loops.append((idx[tg], i))


def reg(tok):
    m = re.match(r"-?\|?R(\d+)", tok)
    return int(m.group(1)) if m else None


for lo, hi in loops:
    if any(l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops) or hi - lo < min_len:
        continue
    last, hist, n = {}, {}, 0
    for k, (a, t) in enumerate(ins[lo:hi + 1]):
        t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t2.split()[0]
        args = [x.strip() for x in t2[len(op):].split(",")]
        if op.startswith(("DFMA", "DMUL", "DADD")):
            n += 1
            src = [reg(x) for x in args[1:]]
            # a 64-bit operand Rn names the pair Rn, Rn+1
            d = min([k - last[r] for s in src if s is not None for r in (s, s + 1) if r in last] or [99])
            hist[min(d, 8)] = hist.get(min(d, 8), 0) + 1
        d0 = reg(args[0]) if args and re.match(r"R\d+", args[0]) else None
        if d0 is not None:
            last[d0] = k
            if op.startswith(("D", "LDS.64", "LDS.128", "IMAD.WIDE", "I2F.F64", "F2F.F64", "MUFU.RCP64")) or ".64" in op:
                last[d0 + 1] = k
    tot = sum(hist.values())
    print(f"loop {ins[lo][0]:#07x}: {hi - lo + 1:4d} instr, {n:3d} FP64 ops; producer distance (8 = 8 or more): "
          + "  ".join(f"{d}:{100 * c // tot}%" for d, c in sorted(hist.items())))
