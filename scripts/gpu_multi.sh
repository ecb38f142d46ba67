#!/bin/bash
# 8-GPU call: strong scaling of config 5 at N = 8, 4, 2 (fixed grid), config 4 at N = 8, the 2-GPU peer-gather test
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { # N, port, extra args, output stem
  local n=$1 port=$2 stem=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 3 --warmup 3 "$@" > gpurun_out/$stem.json 2> gpurun_out/$stem.err
  echo "== $stem rc=$?"; tail -c 1800 gpurun_out/$stem.json; tail -2 gpurun_out/$stem.err | cut -c1-300
}
run 8 29521 bench_cfg5_n8
run 8 29522 bench_cfg4_n8 --workload cfg4 --no-weak
run 4 29523 bench_cfg5_n4 --no-weak
run 2 29524 bench_cfg5_n2 --no-weak
timeout 600 python -m pytest tests/test_peer_gather.py -m gpu -x -q --timeout 500 2>&1 | tail -2
