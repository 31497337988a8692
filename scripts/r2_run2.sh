#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_sharding_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t4.log
python scripts/ab_round.py 3 20 6 > gpurun_out/r2_ab2_fused_w6.json 2> gpurun_out/r2_ab2_fused_w6.err
python scripts/ab_round.py 3 20 1 > gpurun_out/r2_ab2_fused_w1.json 2> gpurun_out/r2_ab2_fused_w1.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
cat gpurun_out/r2_t4.log gpurun_out/r2_ab2_fused_w6.json gpurun_out/r2_ab2_fused_w1.json
