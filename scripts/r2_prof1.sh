#!/bin/bash
set -u
mkdir -p gpurun_out
for w in 6 2; do
  python scripts/ab_round.py 3 20 $w > gpurun_out/r2_ab_fused_w$w.json 2> gpurun_out/r2_ab_fused_w$w.err
  DMT_STEP=classic python scripts/ab_round.py 3 20 $w > gpurun_out/r2_ab_classic_w$w.json 2> gpurun_out/r2_ab_classic_w$w.err
done
python scripts/profile_small.py org > gpurun_out/r2_small_org.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ae_fwd_rows|ae_dec_chunks|ae_bwd_phase|ae_grad_phase|norm_prepare|adam_shadow" -s 36 -c 12 -f -o gpurun_out/r2_prof_fused python scripts/profile_small.py org > gpurun_out/r2_ncu_small_org.log 2>&1
tail -2 gpurun_out/r2_ncu_small_org.log
cat gpurun_out/r2_ab_*.json
