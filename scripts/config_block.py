"""One BASELINE.json configuration block of the bench line alone (no CPU leg):
python scripts/config_block.py douban|amazon [rounds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200  # noqa: F401
from dmtcdr_b200 import native

native.load()
torch.cuda.set_device(0)
key = sys.argv[1] if len(sys.argv) > 1 else "douban"
n_rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for k, control, data_name in bench.OTHER_CONFIGS:
    if k == key:
        out = bench.run_config_block(control, data_name, "cuda:0", n_rounds=n_rounds, cpu_leg=False)
        out.pop("step_classes", None)
        print(json.dumps(out))
