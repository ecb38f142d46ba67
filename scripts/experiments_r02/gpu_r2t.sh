#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "organisations or quantum or collapse" 2>&1 | tail -3
for w in "cfg5 --ranks 8 --rank 3 --schedule 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 2" "cfg5 --ranks 4 --rank 1 --schedule 3" "cfg5 --ranks 4 --rank 1 --schedule 3 --parts 1" "cfg5 --ranks 2 --rank 1 --schedule 3" "cfg5 --schedule 3" "cfg4 --schedule 3"; do
  echo "== $w"; timeout 200 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
