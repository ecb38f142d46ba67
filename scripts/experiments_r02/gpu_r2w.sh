#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_march$' -c 1 -f -o gpurun_out/prof_march_full python scripts/dev_bench.py --workload cfg5 --schedule 3 --reps 1 > gpurun_out/ncu_march_full.log 2>&1
tail -1 gpurun_out/ncu_march_full.log
