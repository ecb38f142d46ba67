#!/bin/bash
# 2-GPU call: peer-gather test, the new analytic-derivative GPU test, strong-scaling bench at N=2 (peer and NCCL gather)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_peer_gather.py tests/test_thermo_derivatives.py -m gpu -x -q --timeout 600 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --gather nccl --no-weak --no-e2e > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; tail -c 700 gpurun_out/bench_n2_nccl.json
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --layout slab --no-weak --no-e2e > gpurun_out/bench_n2_slab.json 2> gpurun_out/bench_n2_slab.err; tail -c 700 gpurun_out/bench_n2_slab.json
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --workload cfg4 --no-weak > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err; tail -c 700 gpurun_out/bench_cfg4_n2.json
