#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/check_sharded_dropin.py > gpurun_out/r2_sharded_dropin_n2.log 2>&1
echo "sharded check rc=$?"; grep -E "identical_F|Error|error" gpurun_out/r2_sharded_dropin_n2.log | cut -c1-400 | tail -4
timeout 300 $TR --master-port 29534 scripts/check_sharded_dropin.py Douban_user_explicit_ae_0_genre_assist_optim-0.3_constant tiny-Douban >> gpurun_out/r2_sharded_dropin_n2.log 2>&1
echo "sharded check 2 rc=$?"; grep -E "identical_F" gpurun_out/r2_sharded_dropin_n2.log | cut -c1-300 | tail -2
timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --configs none > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > /dev/null 2>&1; echo "ref arm (rank0 only path) rc=$?"
