#!/usr/bin/env python
"""Code size of every device function of one kernel: python scripts/func_sizes.py lib.so k_march"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for l in out.splitlines():
    p = l.split()
    if len(p) >= 7 and p[0].startswith("0x") and pat in p[-1]:
        try:
            size = int(p[2], 16)
        except ValueError:
            continue
        name = p[-1].split("$")[-1]
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        rows.append((size, d[:110]))
for s, n in sorted(rows):
    if s:
        print("%8d  %s" % (s, n))
