#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_t5.log
cat gpurun_out/r2_t5.log
