#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_sharding_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t9.log
cat gpurun_out/r2_t9.log
timeout 1000 python -m pytest tests/test_reference_driver_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_t9b.log
tail -30 gpurun_out/r2_t9b.log
out=gpurun_out/r2_ab9.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab9.err; }
W=1; run DMT_ROWS=tma; run DMT_ROWS=stream
W=8; run DMT_ROWS=tma; run DMT_ROWS=stream
echo "# douban tma" >> $out; DMT_ROWS=tma timeout 300 python scripts/config_block.py douban >> $out 2>> gpurun_out/r2_ab9.err
cat $out | cut -c1-700; tail -5 gpurun_out/r2_ab9.err
