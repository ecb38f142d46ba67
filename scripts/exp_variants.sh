#!/bin/bash
# Development aid: time cfg5 (device-resident, no CPU baseline, no e2e) for a list of "ENV=... ENV=..." settings.
# usage: scripts/exp_variants.sh "PNJL_WS_PRIO=0" "PNJL_WS_PRIO=1" "PNJL_LIB=path/to/variant.so" ...
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -o '"value": [0-9.]*\|"frac": [0-9.]*\|phases.*' | tr '\n' ' '
  echo
done
