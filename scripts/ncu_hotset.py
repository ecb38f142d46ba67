#!/usr/bin/env python
"""Hot instruction set of a kernel from an ncu source-page CSV: which device functions the instructions executed at (at
least) the per-pass rate belong to, and stall samples per function.
    ncu -i rep --page source --csv > src.csv ; cuobjdump -elf lib.so | grep <kernel> | grep '\\$' | awk '{print $2,$3,$NF}' > funcs.txt
    python scripts/ncu_hotset.py src.csv funcs.txt"""
import collections
import csv
import subprocess
import sys

import numpy as np

funcs = []
for l in open(sys.argv[2]):
    p = l.split()
    if len(p) != 3 or not p[0].startswith("0x"):
        continue
    d = subprocess.run(["c++filt", p[2].split("$")[-1]], capture_output=True, text=True).stdout.strip().split("(")[0][:44]
    funcs.append((int(p[0], 16), int(p[0], 16) + int(p[1], 16), d))
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: hdr.index(n) for n in ("Address", "# Samples", "Instructions Executed", "stall_no_inst", "stall_barrier", "stall_long_sb",
                                  "stall_wait", "stall_sleep", "stall_short_sb", "stall_math")}
first = None
data = []


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


for r in rows[2:]:
    try:
        addr = int(r[col["Address"]], 16)
    except ValueError:
        continue
    if first is None:
        first = addr
    off = addr - first
    f = "<main body>"
    for a, b, d in funcs:
        if a <= off < b:
            f = d
            break
    data.append((f, {k: num(r[i]) for k, i in col.items() if k != "Address"}))
ex = np.array([d["Instructions Executed"] for _, d in data])
loop_rate = np.median(np.sort(ex)[-150:])
per_pass = loop_rate / 32 / 0.7
tot_s = sum(d["# Samples"] for _, d in data)
print("per-pass execution rate ~ %.3g; total samples %d" % (per_pass, tot_s))
for thr, name in ((0.5, "per-pass set"), (0.12, "per-point set")):
    c = collections.Counter()
    for (f, d) in data:
        if d["Instructions Executed"] >= thr * per_pass:
            c[f] += 1
    print("%s: %d instructions = %.1f KB: %s" % (name, sum(c.values()), sum(c.values()) * 16 / 1024.0,
                                                ", ".join("%s %d" % (k, v) for k, v in c.most_common(6))))
agg = collections.defaultdict(collections.Counter)
for f, d in data:
    for k, v in d.items():
        agg[f][k] += v
for f, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:8]:
    s = max(v["# Samples"], 1.0)
    print("%-44s %5.1f%% of samples | no_inst %2.0f%% barrier %2.0f%% long_sb %2.0f%% wait %2.0f%% sleep %2.0f%% short_sb %2.0f%% math %2.0f%%" % (
        f, 100 * s / tot_s, *[100 * v[k] / s for k in ("stall_no_inst", "stall_barrier", "stall_long_sb", "stall_wait", "stall_sleep",
                                                        "stall_short_sb", "stall_math")]))
