#!/bin/bash
# One GPU round: parity tests, smoke, headline bench, ncu launch list and one full capture of the top kernel —
# both ncu passes on the SAME command as the bench (the full cfg5 grid), so that profile and bench line describe one launch.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ -z "$SKIP_TESTS" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -5
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -c 3000 gpurun_out/bench_cfg5.json; tail -5 gpurun_out/bench_cfg5.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1500 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_ws -s 3 -c 1 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-flush > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 1200 gpurun_out/bench_cfg4.json
python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 1200 gpurun_out/bench_cfg3.json
ls -la gpurun_out
