#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 3 --warmup 3 --configs none > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300; cut -c1-300 gpurun_out/r2_bench_n$N.json
