#!/usr/bin/env python
"""Per-function cost of one kernel from an ncu source-page CSV: share of samples, calls, instructions and cycles per call, stalls.
    python scripts/ncu_funcs.py src.csv funcs.txt <kernel cycles> [warps per SM] [SMs]"""
import csv
import subprocess
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ins, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
funcs = []
for l in open(sys.argv[2]):
    p = l.split()
    d = subprocess.run(["c++filt", p[2].split("$")[-1]], capture_output=True, text=True).stdout.strip().split("(")[0]
    funcs.append((int(p[0], 16), int(p[0], 16) + int(p[1], 16), d))
cycles = float(sys.argv[3])
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 16
sms = int(sys.argv[5]) if len(sys.argv) > 5 else 148
first = int(rows[2][ia], 16)
agg, tot = {}, 0.0
for r in rows[2:]:
    off = int(r[ia], 16) - first
    f = "<main>"
    for a, b, d in funcs:
        if a <= off < b:
            f = d
            break
    g = agg.setdefault(f, {"smp": 0.0, "ex": 0.0, "first": None, **{s: 0.0 for s in stalls}})
    g["smp"] += float(r[ins] or 0)
    g["ex"] += float(r[iex] or 0)
    tot += float(r[ins] or 0)
    if g["first"] is None:
        g["first"] = float(r[iex] or 0)
    for s in stalls:
        g[s] += float(r[hdr.index(s)] or 0)
cps = warps * sms * cycles / tot
print("total samples %d, %.0f warp-cycles per sample" % (tot, cps))
for f, g in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:10]:
    calls = g["first"] or 1
    t = sum(g[s] for s in stalls) or 1
    print("%-34s %5.1f%%  calls %.3g  instr/call %.0f  cycles/call %.0f (%.1f/instr) | %s" % (
        f[:34], 100 * g["smp"] / tot, calls, g["ex"] / calls, g["smp"] * cps / calls, g["smp"] * cps / max(g["ex"], 1),
        ", ".join("%s %.0f%%" % (s[6:], 100 * g[s] / t) for s in sorted(stalls, key=lambda s: -g[s])[:6])))
