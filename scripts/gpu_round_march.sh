#!/bin/bash
# second call of a round (gpurun_out is limited to 64 MiB per call): full captures of the line-march kernel
set -x
mkdir -p gpurun_out
# line-march kernel: a 1/8 share of config 5 (what a rank of an 8-GPU run carries; teams of four warps) and config 4 (one warp per line)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_march$ -c 1 -f -o gpurun_out/prof_march_share8 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --reps 1 > gpurun_out/ncu_march_share8.log 2>&1
tail -1 gpurun_out/ncu_march_share8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_march$ -c 1 -f -o gpurun_out/prof_march_cfg4 python scripts/dev_bench.py --workload cfg4 --reps 1 > gpurun_out/ncu_march_cfg4.log 2>&1
tail -1 gpurun_out/ncu_march_cfg4.log
