#!/usr/bin/env python
"""Opcode histogram of one device function inside a kernel:
   python scripts/func_opcodes.py lib.so <kernel mangled substring> <function substring>"""
import collections
import re
import subprocess
import sys

lib, kpat, fpat = sys.argv[1], sys.argv[2], sys.argv[3]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rng = None
kname = None
for l in elf.splitlines():
    p = l.split()
    if len(p) >= 7 and p[0].startswith("0x") and kpat in p[-1]:
        name = p[-1]
        if "$" in name and fpat in name.split("$")[-1]:
            rng = (int(p[1], 16), int(p[1], 16) + int(p[2], 16))
        if "$" not in name and not name.startswith("."):
            kname = name
sass = subprocess.run(["cuobjdump", "-sass", "-fun", kname, lib], capture_output=True, text=True).stdout
hist = collections.Counter()
n = 0
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if rng[0] <= a < rng[1]:
        t = m.group(2).strip()
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0].split(".")[0]] += 1
        n += 1
print("%s: %d instructions (%d bytes)" % (fpat, n, rng[1] - rng[0]))
for k, v in hist.most_common(25):
    print("  %-10s %5d" % (k, v))
