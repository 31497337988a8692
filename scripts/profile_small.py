"""Small programs for `ncu --set full` captures (each finishes in seconds without ncu).
  python scripts/profile_small.py hbm   -> the decoder kernel on the HBM-bound shape (bench.hbm_bound_case)
  python scripts/profile_small.py org   -> one organization, one local epoch at ML1M shape (every step kernel)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop

mode = sys.argv[1] if len(sys.argv) > 1 else "hbm"
if mode == "hbm":
    hbm, _ = bench.measured_peaks()
    print(bench.hbm_bound_case("cuda:0", hbm))
else:
    data, dataset, data_split, mats, cfg = bench.build_problem()
    R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=1, rank=0, world=18,
                               device="cuda:0")
    # one organization is enough for a capture, but it runs with the launch configuration of the bench's timed rounds
    # (18 organizations on the GPU): 8-row tiles, the reduced decoder grid
    for eng in R.eng.values():
        eng.h.set_row_tile(roundloop.row_tile_for(18))
        eng.h.set_decoder_blocks(roundloop.decoder_blocks_for(18))
    R.round0()
    R.run_round(1)
    R.sync()
    print("ok")
