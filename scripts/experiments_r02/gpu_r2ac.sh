#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py tests/test_gpu_api.py -m gpu -x -q --timeout 600 2>&1 | tail -3
for b in 1 0; do for w in "cfg5 --schedule 3" "cfg5 --ranks 2 --rank 1" "cfg5 --ranks 8 --rank 3" "cfg4"; do
  echo "== boot $b: $w"; PNJL_MARCH_BOOT=$b timeout 300 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes.*//'
done; done
echo "== WS"; timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'
