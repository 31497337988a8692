#!/bin/bash
# weight-streamed (bulk copy) row kernels: parity, then A/B against the register-streamed form
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_sharding_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t7.log
cat gpurun_out/r2_t7.log
out=gpurun_out/r2_ab7.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab7.err; }
W=1; run DMT_ROWS=tma; run DMT_ROWS=stream
W=8; run DMT_ROWS=tma; run DMT_ROWS=stream
W=4; run DMT_ROWS=tma
echo "# douban tma" >> $out; DMT_ROWS=tma timeout 300 python scripts/config_block.py douban >> $out 2>> gpurun_out/r2_ab7.err
echo "# douban stream" >> $out; DMT_ROWS=stream timeout 300 python scripts/config_block.py douban >> $out 2>> gpurun_out/r2_ab7.err
echo "# amazon tma" >> $out; DMT_ROWS=tma timeout 300 python scripts/config_block.py amazon >> $out 2>> gpurun_out/r2_ab7.err
cat $out | cut -c1-1500; tail -5 gpurun_out/r2_ab7.err
