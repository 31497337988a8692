#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "benchmark_shape and gather and not bulk and not classic" 2>&1 | tail -3
DMT_STREAM_ROWS=4 timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "benchmark_shape and gather and not bulk and not classic" 2>&1 | tail -3
out=gpurun_out/r2_ab14.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab14.err; }
W=8; run DMT_STREAM_ROWS=8; run DMT_STREAM_ROWS=4; run DMT_STREAM_ROWS=16
W=1; run DMT_STREAM_ROWS=4; run DMT_STREAM_ROWS=16
W=4; run DMT_STREAM_ROWS=4
cat $out | cut -c1-420; tail -5 gpurun_out/r2_ab14.err
