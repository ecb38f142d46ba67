#!/usr/bin/env python
"""Hottest source lines of a kernel from `ncu -i rep --page source --print-source cuda,sass --csv > cs.csv`:
    python scripts/ncu_lines.py cs.csv [top N] [file substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
filt = sys.argv[3] if len(sys.argv) > 3 else ""
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    g = lambda name: float(r[hdr.index(name)] or 0)
    out.append((g("# Samples"), cur, r[0], r[1].strip()[:90], g("Instructions Executed"), g("stall_wait"), g("stall_long_sb"), g("stall_short_sb"),
                g("stall_math"), g("stall_barrier"), g("stall_no_inst"), g("stall_sleep")))
tot = sum(o[0] for o in out)
print("total samples %d" % tot)
out.sort(key=lambda o: -o[0])
print("%6s %5s  %-22s %-90s %10s | wait long short math barrier no_inst sleep" % ("smp", "%", "file:line", "source", "warp-instr"))
for o in [o for o in out if filt in o[1]][:top]:
    print("%6d %5.2f  %-22s %-90s %10d | %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f" % (
        o[0], 100 * o[0] / tot, (o[1] + ":" + o[2])[-22:], o[3], o[4], *[100 * v / max(o[0], 1) for v in o[5:]]))
