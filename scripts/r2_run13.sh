#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2_t13.log
cat gpurun_out/r2_t13.log | cut -c1-300
out=gpurun_out/r2_ab13.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab13.err; }
W=8; run DMT_PDL=0; run DMT_PDL=1
W=4; run DMT_PDL=0; run DMT_PDL=1
W=1; run DMT_PDL=0; run DMT_PDL=1
cat $out | cut -c1-420; tail -5 gpurun_out/r2_ab13.err
timeout 300 python scripts/profile_nmf.py nmf 2>&1 | head -3 > gpurun_out/r2_profile_nmf2.txt; head -3 gpurun_out/r2_profile_nmf2.txt
