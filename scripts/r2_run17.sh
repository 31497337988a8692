#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropin_gpu.py tests/test_reference_driver_gpu.py tests/test_sharding_gpu.py tests/test_metrics_privacy_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --configs none > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_quick.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('e2e_with_state_dicts',{}).get('ms_per_step'), 'd2h', d['e2e']['d2h_bytes_per_step'])
PY
