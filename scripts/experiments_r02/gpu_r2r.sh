#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "organisations or quantum or collapse" 2>&1 | tail -6
for w in "cfg5 --ranks 8 --rank 3 --schedule 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --duo 0" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 8" "cfg5 --ranks 4 --rank 1 --schedule 3" "cfg5 --ranks 4 --rank 1 --schedule 3 --parts 4" "cfg5 --ranks 2 --rank 1 --schedule 3 --parts 4" "cfg5 --schedule 3 --parts 4"; do
  echo "== $w"; timeout 200 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march_duo -c 1 -f -o gpurun_out/prof_duo python scripts/dev_bench.py --workload cfg5 --schedule 3 --ranks 8 --rank 3 --reps 1 > gpurun_out/ncu_duo.log 2>&1
tail -1 gpurun_out/ncu_duo.log
