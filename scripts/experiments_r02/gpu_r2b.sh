#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "organisations or golden or cross_first or fine_mesh" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q --timeout 400 -s 2>&1 | tail -12
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 0 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 --ranks 8 --rank 2 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march$ -c 1 -f -o gpurun_out/prof_march python scripts/dev_bench.py --workload cfg5 --schedule 0 --reps 1 --n-t 256 > gpurun_out/ncu_march.log 2>&1
tail -3 gpurun_out/ncu_march.log
