#!/bin/bash
# ncu evidence for the bench command (run under gpurun; see profiles/README.md)
set -u
CMD="python bench.py --steps 1 --warmup 1 --local-epochs 2"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 13000 -c 8000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
for spec in "ae_decoder_chunk:40:2:decoder_chunk" "segment_chunks:80:2:segment_chunks" "adam_kernel:40:2:adam" "sgemm_kernel:240:3:sgemm" "ae_decoder_fwd:72:2:decoder_hbm_case"; do
  IFS=: read -r pat skip cnt name <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
ls -la gpurun_out/
