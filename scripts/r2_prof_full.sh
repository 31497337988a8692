#!/bin/bash
# ncu --set full capture of every kernel of the fused step on the launch configuration of the bench's timed rounds
set -u
mkdir -p gpurun_out
python scripts/profile_small.py org > gpurun_out/r2_small_org.log 2>&1 || { echo "small run failed"; tail -5 gpurun_out/r2_small_org.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ae_fwd_rows|ae_dec_chunks|ae_bwd_rows|ae_seg_chunks|ae_grad_phase|norm_prepare|adam_shadow" -s 42 -c 14 -f -o gpurun_out/r2_prof_fused python scripts/profile_small.py org > gpurun_out/r2_ncu_small_org.log 2>&1
tail -2 gpurun_out/r2_ncu_small_org.log
ncu -i gpurun_out/r2_prof_fused.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_fused.csv
timeout 300 python scripts/profile_nmf.py mf 2>&1 | head -3
