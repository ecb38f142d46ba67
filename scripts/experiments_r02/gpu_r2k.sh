#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "organisations or quantum" 2>&1 | tail -3
for w in "cfg5 --ranks 8 --rank 3 --schedule 3" "cfg5 --ranks 8 --rank 3 --schedule 3 --parts 4" "cfg5 --ranks 4 --rank 1 --schedule 3 --parts 2" "cfg4 --schedule 3" "cfg4 --ranks 8 --rank 2 --schedule 3"; do
  echo "== $w"; timeout 300 python scripts/dev_bench.py --workload $w 2>&1 | tail -1 | sed 's/ | lanes/\n   lanes/'
done
echo "== cfg3 WS variants"
for e in "PNJL_WS_CTRL=2" "PNJL_WS_CTRL=3 PNJL_WS_WORKERS=13" "PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12" "PNJL_WS_CTRL=4 PNJL_WS_WORKERS=12 PNJL_WS_SLOTS=64" "PNJL_WS_CTRL=2 PNJL_WS_SLOTS=64" "PNJL_WS_CTRL=6 PNJL_WS_WORKERS=10 PNJL_WS_SLOTS=96" "PNJL_WS_WSOLVE=1"; do
  echo "-- $e"; env $e timeout 300 python scripts/dev_bench.py --workload cfg3 --schedule 2 --n-t 60000 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_ws -c 1 -f -o gpurun_out/prof_ws_cfg3 python scripts/dev_bench.py --workload cfg3 --schedule 2 --reps 1 --n-t 30000 > gpurun_out/ncu_ws_cfg3.log 2>&1
tail -1 gpurun_out/ncu_ws_cfg3.log
