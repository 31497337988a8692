#!/bin/bash
# ncu evidence for round 2 (run under gpurun AFTER the same commands exited 0 without ncu; see profiles/README.md):
#   gpurun_out/r2_launches_bench.csv   launch list of the bench command (gpu__time_duration.sum per launch)
#   gpurun_out/r2_prof_fused.ncu-rep   --set full capture of every kernel of the fused step (one ML1M-shape organization)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --local-epochs 2 --configs none"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
python scripts/profile_small.py org > gpurun_out/r2_small_org.log 2>&1 || { echo "small run failed"; tail -5 gpurun_out/r2_small_org.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ae_fwd_rows|ae_dec_chunks|ae_bwd_rows|ae_seg_chunks|ae_grad_phase|norm_prepare|adam_shadow" -s 42 -c 14 -f -o gpurun_out/r2_prof_fused python scripts/profile_small.py org > gpurun_out/r2_ncu_small_org.log 2>&1
tail -2 gpurun_out/r2_ncu_small_org.log
ncu -i gpurun_out/r2_prof_fused.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_fused.csv
