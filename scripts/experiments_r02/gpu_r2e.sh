#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'; }
run PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12
run PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12 PNJL_WS_PARTS=4
run PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9
run PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9 PNJL_WS_PARTS=4
run PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9 PNJL_WS_PARTS=1
run PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12 PNJL_WS_WSOLVE=1
run PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9 PNJL_WS_WSOLVE=1
run PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9 PNJL_WS_WSOLVE=1 PNJL_WS_PARTS=4
run2() { echo "== $*"; env "$@" timeout 300 python scripts/dev_bench.py --workload cfg5 --ranks 4 --rank 1 --schedule 2 2>&1 | tail -1 | sed 's/ | lanes.*//'; }
run2 PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12
run2 PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12 PNJL_WS_PARTS=2
run2 PNJL_WS_CTRL=7 PNJL_WS_WORKERS=9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march$ -c 1 -f -o gpurun_out/prof_march_slab env PNJL_MARCH_LOCKSTEP=0 python scripts/dev_bench.py --workload cfg5 --ranks 8 --rank 3 --schedule 0 --reps 1 --n-t 256 > gpurun_out/ncu_march_slab.log 2>&1
tail -2 gpurun_out/ncu_march_slab.log
