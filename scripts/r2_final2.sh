#!/bin/bash
# final state of round 2 on one B200: smoke, the whole -m gpu suite, the default bench line
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err | cut -c1-200
