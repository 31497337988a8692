#!/bin/bash
# round 2 checkpoint: GPU tests, the bench line, the launch list and full ncu captures of the fused step kernels
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --local-epochs 2 --configs none"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
python scripts/profile_small.py org > gpurun_out/r2_small_org.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ae_fwd_rows|ae_dec_chunks|ae_bwd_rows|ae_seg_chunks|ae_grad_phase|norm_prepare|adam_shadow" -s 42 -c 14 -f -o gpurun_out/r2_prof_fused python scripts/profile_small.py org > gpurun_out/r2_ncu_small_org.log 2>&1
tail -2 gpurun_out/r2_ncu_small_org.log
ls -la gpurun_out/
