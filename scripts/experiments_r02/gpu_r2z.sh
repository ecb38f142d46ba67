#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -3
for w in "cfg3 --schedule 2 --n-t 60000" "cfg3 --schedule 3 --n-t 60000" "cfg5 --schedule 3" "cfg5 --schedule 2" "cfg5 --ranks 2 --rank 1 --schedule 3" "cfg5 --ranks 4 --rank 1 --schedule 0" "cfg5 --ranks 8 --rank 3 --schedule 0" "cfg4 --schedule 0"; do
  echo "== $w"; timeout 300 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
