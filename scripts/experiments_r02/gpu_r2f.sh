#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python scripts/dev_bench.py --workload cfg3 --ranks 8 --schedule 2 2>&1 | tail -1
timeout 300 python scripts/dev_bench.py --workload cfg3 --ranks 8 --schedule 3 2>&1 | tail -1
timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 0 2>&1 | tail -1
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 2>&1 | tail -1
PNJL_MARCH_PARTS=1 timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 3 2>&1 | tail -1
timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 4 --rank 1 --schedule 0 2>&1 | tail -1
timeout 300 python scripts/dev_bench.py --workload cfg4 --schedule 0 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -c 4000 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
