#!/bin/bash
# bulk-copy gather kernels: parity, then A/B against the register-load form
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_sharding_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t6.log
cat gpurun_out/r2_t6.log
out=gpurun_out/r2_ab6.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab6.err; }
W=1; run DMT_GATHER=bulk; run DMT_GATHER=ldg; run DMT_GATHER=bulk DMT_BULK_BLOCKS=296; run DMT_GATHER=bulk DMT_BULK_BLOCKS=444
W=8; run DMT_GATHER=bulk; run DMT_GATHER=ldg
W=2; run DMT_GATHER=bulk
cat $out; tail -5 gpurun_out/r2_ab6.err
