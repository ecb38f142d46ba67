#!/usr/bin/env python
"""Print the SASS of one device function inside a kernel: python scripts/func_sass.py lib.so <kernel substr> <function substr>"""
import re, subprocess, sys
lib, kpat, fpat = sys.argv[1], sys.argv[2], sys.argv[3]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rng = None; kname = None
for l in elf.splitlines():
    p = l.split()
    if len(p) >= 7 and p[0].startswith("0x") and kpat in p[-1]:
        name = p[-1]
        if "$" in name and fpat in name.split("$")[-1]:
            rng = (int(p[1], 16), int(p[1], 16) + int(p[2], 16))
        if "$" not in name and not name.startswith("."):
            kname = name
sass = subprocess.run(["cuobjdump", "-sass", "-fun", kname, lib], capture_output=True, text=True).stdout
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and rng[0] <= int(m.group(1), 16) < rng[1]:
        print("%05x  %s" % (int(m.group(1), 16), m.group(2).strip()))
