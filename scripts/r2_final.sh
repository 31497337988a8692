#!/bin/bash
# round-2 final validation on one B200: smoke, the whole -m gpu suite, the bench line, the reference arm, ncu evidence
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_ref.json
bash scripts/r2_prof.sh
