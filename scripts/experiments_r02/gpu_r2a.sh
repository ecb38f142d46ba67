#!/bin/bash
# Round 2, first GPU call: correctness of the line-march kernel, then A/B timings against the warp-specialised kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "organisations or golden or cross_first or fine_mesh or edge" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q --timeout 400 -s 2>&1 | tail -30
for s in 0 2; do
  timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule $s 2>&1 | tail -3
done
for s in 0 2; do
  timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule $s 2>&1 | tail -3
  timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 2 --rank 1 --schedule $s 2>&1 | tail -3
done
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 --parts 1 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 --parts 4 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 0 --quantum 8 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 0 --quantum 128 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 2 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 --ranks 8 --rank 2 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 2 --ranks 8 --rank 2 2>&1 | tail -3
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 --ranks 8 --rank 2 --parts 2 2>&1 | tail -3
