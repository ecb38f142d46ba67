#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py -m gpu -x -q --timeout 600 -k "organisations or golden or configs or config or fine_mesh or cross_first or collapse or quantum" 2>&1 | tail -4
echo "== cfg3 WS"; timeout 300 python scripts/dev_bench.py --workload cfg3 --schedule 2 --n-t 60000 2>&1 | tail -1
echo "== cfg5 WS N=1"; timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'
for w in 16 12 8; do for q in 0 32; do echo "== cfg4 march warps $w quantum $q"; PNJL_MARCH_WARPS=$w PNJL_MARCH_Q=$q timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 3 2>&1 | tail -1 | sed 's/ | passes.*//'; done; done
for w in 12 8; do echo "== cfg5 march warps $w"; PNJL_MARCH_WARPS=$w PNJL_MARCH_Q=32 timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 3 2>&1 | tail -1 | sed 's/ | passes.*//'; done
echo "== cfg4 WS"; timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'
echo "== cfg4 1/8 WS"; timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 2 --ranks 8 --rank 2 2>&1 | tail -1 | sed 's/ | lanes.*//'
