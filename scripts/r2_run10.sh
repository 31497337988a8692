#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
timeout 300 python scripts/scaled_round.py 20000 10000 2000000 64 8 1 > gpurun_out/r2_scaled_small.json 2> gpurun_out/r2_scaled_small.err
echo "scaled small rc=$?"; tail -3 gpurun_out/r2_scaled_small.err; cut -c1-400 gpurun_out/r2_scaled_small.json
timeout 900 python scripts/scaled_round.py 400000 200000 100000000 64 8 1 > gpurun_out/r2_scaled_round.json 2> gpurun_out/r2_scaled_round.err
echo "scaled rc=$?"; tail -3 gpurun_out/r2_scaled_round.err; cut -c1-600 gpurun_out/r2_scaled_round.json
bash scripts/r2_prof.sh
