#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py tests/test_gpu_api.py -m gpu -x -q --timeout 600 2>&1 | tail -3
for w in "cfg5 --schedule 3" "cfg5 --schedule 2" "cfg5 --ranks 2 --rank 1" "cfg5 --ranks 4 --rank 1" "cfg5 --ranks 8 --rank 3" "cfg4"; do
  echo "== $w"; timeout 300 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes.*//'
done
