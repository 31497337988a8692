#!/bin/bash
# config-5 family, 64 organizations, strong scaling over N GPUs: 100 000 x 50 000, 20 M ratings, 1 local epoch
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 800 python scripts/scaled_round.py 100000 50000 20000000 64 1 1 > gpurun_out/r2_scale64_n1.json 2> gpurun_out/r2_scale64_n1.err
else
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/scaled_round.py 100000 50000 20000000 64 $N 1 > gpurun_out/r2_scale64_n$N.json 2> gpurun_out/r2_scale64_n$N.err
fi
echo "rc=$?"; tail -3 gpurun_out/r2_scale64_n$N.err | cut -c1-300; cut -c1-700 gpurun_out/r2_scale64_n$N.json
