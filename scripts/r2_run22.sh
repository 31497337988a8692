#!/bin/bash
set -u
mkdir -p gpurun_out
out=gpurun_out/r2_ab22.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab22.err; }
W=6; run DMT_STREAM_ROWS=4; run DMT_STREAM_ROWS=8; run DMT_STREAM_ROWS=4 DMT_DEC_BLOCKS=148; run DMT_STREAM_ROWS=4 DMT_FANOUT=0
cat $out | cut -c1-330; tail -3 gpurun_out/r2_ab22.err
