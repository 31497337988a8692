#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/check_sharded_dropin.py > gpurun_out/r2_sharded_dropin_n2.log 2>&1
echo "sharded check rc=$?"; grep -E "identical_F" gpurun_out/r2_sharded_dropin_n2.log | cut -c1-200 | tail -2
timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --configs none > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n2.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['n_gpus_used'], d['clocks'])
PY
