#!/bin/bash
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool synccheck --print-limit 3 python scripts/march_small.py 12 5 solo > gpurun_out/synccheck_solo_r02.txt 2>&1
tail -5 gpurun_out/synccheck_solo_r02.txt | cut -c1-200
grep "hazard\|and .* access at\|Race reported\|RACECHECK" gpurun_out/racecheck_r02.txt 2>/dev/null | head -5
