#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "organisations or quantum or collapse" 2>&1 | tail -4
for w in "cfg5 --ranks 8 --rank 3 --schedule 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 5" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 2" "cfg5 --ranks 4 --rank 1 --schedule 3" "cfg5 --ranks 4 --rank 1 --schedule 3 --parts 3"; do
  echo "== $w"; timeout 200 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
PNJL_MARCH_PARTS=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "full_size_config5" 2>&1 | tail -3
