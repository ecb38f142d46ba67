#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_tmu_scan.py tests/test_dual_branch.py -m gpu -x -q --timeout 400 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q --timeout 400 -s 2>&1 | tail -12
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 2 2>&1 | tail -2
PNJL_WS_PARTS=4 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 2 2>&1 | tail -2
PNJL_WS_PARTS=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 2 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 4 --rank 1 --schedule 2 2>&1 | tail -2
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 2 --rank 1 --schedule 2 2>&1 | tail -2
PNJL_WS_WORKERS=15 PNJL_WS_CTRL=1 PNJL_WS_SLOTS=32 timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -2
PNJL_WS_SPW=3 timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_ws -c 1 -f -o gpurun_out/prof_wsplus python scripts/dev_bench.py --workload cfg5 --schedule 2 --reps 1 --n-t 256 > gpurun_out/ncu_wsplus.log 2>&1
tail -3 gpurun_out/ncu_wsplus.log
