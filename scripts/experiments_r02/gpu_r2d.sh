#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -8
for r in "8 3" "4 1" "2 1"; do set -- $r
  PNJL_MARCH_LOCKSTEP=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks $1 --rank $2 --schedule 0 2>&1 | tail -1
  PNJL_MARCH_LOCKSTEP=0 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks $1 --rank $2 --schedule 0 2>&1 | tail -1
  timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks $1 --rank $2 --schedule 2 2>&1 | tail -1
  PNJL_WS_WSOLVE=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks $1 --rank $2 --schedule 2 2>&1 | tail -1
done
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1
PNJL_MARCH_LOCKSTEP=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 0 2>&1 | tail -1
PNJL_MARCH_LOCKSTEP=1 timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 --ranks 8 --rank 2 2>&1 | tail -1
PNJL_MARCH_LOCKSTEP=0 timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 --ranks 8 --rank 2 2>&1 | tail -1
PNJL_MARCH_LOCKSTEP=1 PNJL_MARCH_PARTS=4 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 2>&1 | tail -1
PNJL_MARCH_LOCKSTEP=1 PNJL_MARCH_PARTS=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 2>&1 | tail -1
