#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_dropin_gpu.py -m gpu -q -k "lbfgs or joint" 2>&1 | tail -60 > gpurun_out/r2_t12.log
cat gpurun_out/r2_t12.log | cut -c1-300
