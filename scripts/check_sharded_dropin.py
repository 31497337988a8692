"""Organization sharding BEHIND the drop-in API, checked on real ranks (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_sharded_dropin.py

Every rank runs the same experiment through Assist / Organization (device-RNG mode: permutations, initial parameters and
dropout are seeded per organization and round); the drop-in classes shard the organizations and all-gather their outputs
over NCCL. Afterwards every rank repeats the experiment unsharded (DMT_SHARD=0) and the global predictions F_t of every
round must be IDENTICAL — same kernels, same seeds, same summation order."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmtcdr_b200  # noqa: E402,F401
from dmtcdr_b200 import dist as D  # noqa: E402
from dmtcdr_b200 import runner, synth  # noqa: E402


def main():
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    control = sys.argv[1] if len(sys.argv) > 1 else "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_optim_0.5"
    data_name = sys.argv[2] if len(sys.argv) > 2 else "tiny-Amazon"
    data = synth.make_rating_data(data_name, seed=0)
    kw = dict(seed=0, local_epochs=2, rounds=2, rng="device", keep_objects=True)
    from dmtcdr_b200.config import cfg
    sharded = runner.run_assist_experiment(data, control, **kw)
    owners = [o.__dict__.get("_owner_rank") for o in sharded["organization"]]
    trained_here = [i for i, o in enumerate(sharded["organization"]) if o.model_state_dict[1] is not None]
    os.environ["DMT_SHARD"] = "0"
    single = runner.run_assist_experiment(data, control, **kw)
    os.environ["DMT_SHARD"] = "1"
    ok = True
    for t in range(len(single["F"])):
        for k in ("train", "test"):
            same = np.array_equal(sharded["F"][t][k], single["F"][t][k])
            ok = ok and same
    assert world == 1 or sorted(set(owners)) == list(range(min(world, len(owners)))), owners
    assert world == 1 or all(owners[i] == rank for i in trained_here), (owners, trained_here)
    assert ok, "sharded drop-in run differs from the single-process run"
    assert single["metrics"] == sharded["metrics"]
    D.barrier()
    print(json.dumps({"rank": rank, "world": world, "owners": owners, "trained_here": trained_here,
                      "identical_F": ok, "metrics_last": sharded["metrics"][2]}), flush=True)


if __name__ == "__main__":
    main()
