#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_dropin_gpu.py tests/test_metrics_privacy_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_t11.log
cat gpurun_out/r2_t11.log
echo "# amazon device lbfgs" > gpurun_out/r2_ab11.jsonl
timeout 300 python scripts/config_block.py amazon 3 >> gpurun_out/r2_ab11.jsonl 2>> gpurun_out/r2_ab11.err
echo "# amazon host lbfgs" >> gpurun_out/r2_ab11.jsonl
DMT_LBFGS=host timeout 300 python scripts/config_block.py amazon 3 >> gpurun_out/r2_ab11.jsonl 2>> gpurun_out/r2_ab11.err
cut -c1-700 gpurun_out/r2_ab11.jsonl; tail -5 gpurun_out/r2_ab11.err
timeout 300 python scripts/profile_nmf.py nmf > gpurun_out/r2_profile_nmf.txt 2>&1
head -60 gpurun_out/r2_profile_nmf.txt | cut -c1-220
