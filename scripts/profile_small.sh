#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/profile_small.py hbm > gpurun_out/small_hbm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ae_decoder_fwd -s 3 -c 2 -f -o gpurun_out/prof_decoder_hbm python scripts/profile_small.py hbm > gpurun_out/ncu_small_hbm.log 2>&1
tail -2 gpurun_out/ncu_small_hbm.log
python scripts/profile_small.py org > gpurun_out/small_org.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"sgemm_kernel|ae_encoder_fwd_kernel|segment_finish|ae_decoder_finish|sqnorm|colsum" -s 60 -c 14 -f -o gpurun_out/prof_step_small python scripts/profile_small.py org > gpurun_out/ncu_small_org.log 2>&1
tail -2 gpurun_out/ncu_small_org.log
ls -la gpurun_out/*.ncu-rep
