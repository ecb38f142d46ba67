#!/usr/bin/env python
"""Find the innermost loops of a kernel in `cuobjdump -sass` output and classify their FP64 instructions.

    cuobjdump -sass -fun <mangled kernel> lib.so > k.sass ; python scripts/sass_loops.py k.sass [min_len]

For every backward branch (loop) longer than min_len instructions: instruction count, FP64-pipe instructions (DFMA / DMUL /
DADD / DSETP), how many of them read three DISTINCT vector-register pairs that are not covered by a .reuse hit (those
occupy the FP64 pipe for 3 cycles instead of 2 on sm_100, scripts/microbench_dfma_operands*.cu), MUFU count, LDS count.
"""
import re
import sys

path = sys.argv[1]
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ins = []
for line in open(path):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2idx and i - addr2idx[tgt] >= min_len:
            loops.append((addr2idx[tgt], i))
# innermost only
inner = [l for l in loops if not any(o != l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
reg = re.compile(r"\bR(\d+)\b(\.reuse)?")
for s, e in inner:
    body = ins[s:e + 1]
    n = len(body)
    fp = [t for _, t in body if re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD|DSETP)", t)]
    three = 0
    reuse = 0
    for t in fp:
        ops = t.split(None, 1)[1] if not t.startswith("@") else t.split(None, 2)[2]
        parts = [p.strip() for p in ops.split(",")]
        srcs = parts[1:]
        if t.lstrip("@!UP0123456789 ").startswith("DSETP"):
            srcs = parts[2:]
        regs = set()
        for p in srcs:
            m = re.match(r"[-|~]*R(\d+)(\.reuse)?", p.strip("|-"))
            if m and "RZ" not in p:
                if m.group(2):
                    reuse += 1
                regs.add(int(m.group(1)))
        if len(regs) >= 3:
            three += 1
    mufu = sum(1 for _, t in body if "MUFU" in t)
    lds = sum(1 for _, t in body if re.search(r"\bLDS", t))
    print("loop 0x%05x..0x%05x: %4d instr, %4d FP64 (%3d with 3 distinct regs, %3d .reuse operands), %2d MUFU, %2d LDS" % (
        body[0][0], body[-1][0], n, len(fp), three, reuse, mufu, lds))
