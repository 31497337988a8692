#!/bin/bash
# 4-warp decoder + prefetching segment kernel: parity, A/B; ncu of the row kernels
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_sharding_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t8.log
cat gpurun_out/r2_t8.log
out=gpurun_out/r2_ab8.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab8.err; }
W=1; run DMT_DEC_FORM=4w; run DMT_DEC_FORM=8w; run DMT_DEC_FORM=4w DMT_DEC_BLOCKS=148; run DMT_DEC_FORM=4w DMT_DEC_BLOCKS=74
W=8; run DMT_DEC_FORM=4w; run DMT_DEC_FORM=8w
cat $out | cut -c1-600; tail -5 gpurun_out/r2_ab8.err
python scripts/profile_small.py org > gpurun_out/r2_small_org.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rows_tma|ae_dec_chunks4|ae_seg_chunks_kernel|ae_grad_phase" -s 30 -c 10 -f -o gpurun_out/r2_prof_rows python scripts/profile_small.py org > gpurun_out/r2_ncu_rows.log 2>&1
tail -2 gpurun_out/r2_ncu_rows.log
