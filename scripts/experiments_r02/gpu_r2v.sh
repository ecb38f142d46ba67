#!/bin/bash
for e in "PNJL_WS_SLOTS=56" "PNJL_WS_SLOTS=64" "PNJL_WS_SLOTS=48" "PNJL_WS_CTRL=3 PNJL_WS_WORKERS=13 PNJL_WS_SLOTS=60" "PNJL_WS_CTRL=3 PNJL_WS_WORKERS=13 PNJL_WS_SLOTS=78" "PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12 PNJL_WS_SLOTS=72"; do
  echo "-- $e"; env $e timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1 | sed 's/ | passes.*//'
done
echo "== auto 1/2"; timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 2 --rank 1 2>&1 | tail -1 | sed 's/ | passes.*//'
