#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharding_gpu.py tests/test_metrics_privacy_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -3
out=gpurun_out/r2_ab18.jsonl; : > $out
run() { echo "# W=$W $*" >> $out; env "$@" timeout 300 python scripts/ab_round.py 3 20 $W >> $out 2>> gpurun_out/r2_ab18.err; }
W=1; run A=1
W=8; run A=1
cat $out | cut -c1-200; tail -3 gpurun_out/r2_ab18.err
timeout 600 python bench.py --configs none > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_quick.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('e2e_with_state_dicts',{}).get('ms_per_step'))
PY
