#!/bin/bash
# N=1 experiments on the warp-specialised kernel + the ncu evidence of HEAD for profiles/
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python scripts/dev_bench.py --workload cfg5 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'; }
run PNJL_X=0
run PNJL_WS_WORKERS=15 PNJL_WS_CTRL=1 PNJL_WS_SLOTS=32
run PNJL_WS_SPW=3
run PNJL_WS_SPW=5
run PNJL_WS_WSOLVE=1
run PNJL_PREDICT_TOL=0
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_ws -s 3 -c 1 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-flush > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 600 gpurun_out/bench_cfg4.json
timeout 600 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 600 gpurun_out/bench_cfg3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1200 gpurun_out/bench_reference.json
