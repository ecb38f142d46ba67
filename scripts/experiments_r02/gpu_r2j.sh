#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py -m gpu -x -q --timeout 600 -k "organisations or golden or configs or config or fine_mesh or cross_first or collapse or quantum" 2>&1 | tail -6
for w in "cfg5 --schedule 3" "cfg5 --schedule 3 --parts 2" "cfg5 --ranks 8 --rank 3 --schedule 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 4" "cfg5 --ranks 4 --rank 1 --schedule 3" "cfg5 --ranks 4 --rank 1 --schedule 3 --parts 2"  "cfg5 --ranks 2 --rank 1 --schedule 3" "cfg5 --ranks 2 --rank 1 --schedule 3 --parts 2" "cfg4 --schedule 3" "cfg4 --schedule 3 --parts 4" "cfg4 --ranks 8 --rank 2 --schedule 3" "cfg3 --schedule 3 --n-t 20000"; do
  echo "== $w"; timeout 300 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march$ -c 1 -f -o gpurun_out/prof_march4 python scripts/dev_bench.py --workload cfg5 --schedule 3 --ranks 8 --rank 3 --reps 1 > gpurun_out/ncu_march4.log 2>&1
tail -1 gpurun_out/ncu_march4.log
