"""torch.profiler view of the module-level path (config 2: alone NMF through the drop-in models): where a step's time
goes, host ops and device kernels. python scripts/profile_nmf.py [nmf|mf]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import dmtcdr_b200
from dmtcdr_b200 import native, runner, synth
from dmtcdr_b200.config import make_cfg

native.load()
which = sys.argv[1] if len(sys.argv) > 1 else "nmf"
control = "ML1M_item_implicit_nmf_0_random-8_alone" if which == "nmf" else "ML1M_user_explicit_mf_0_genre_joint"
cfg = make_cfg(control, device="cuda", seed=0)
models, _, _ = dmtcdr_b200.use_dropin()
data = synth.make_rating_data("ML1M", seed=0)
torch.manual_seed(0)
dataset = runner.fetch_dataset(data)
runner.process_dataset(dataset)
if which == "nmf":
    split = runner.split_dataset(dataset)
    ds = runner.make_split_dataset(dataset, split)[0]["train"]
    model = models.nmf(ds.num_users["data"], ds.num_items["data"]).cuda()
else:
    ds = dataset["train"]
    model = models.mf().cuda()
model.train(True)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4)
perm = torch.randperm(len(ds)).numpy()
batches = [{k: v.cuda() for k, v in runner.pair_batch(ds, perm[s:s + 500]).items()} for s in range(0, len(ds), 500)]
batches = [b for b in batches if len(b[cfg["data_mode"]]) > 0]


def step(b):
    opt.zero_grad()
    out = model(b)
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
    opt.step()


for b in batches:
    step(b)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for b in batches:
    step(b)
torch.cuda.synchronize()
print("steps", len(batches), "ms/step", 1e3 * (time.perf_counter() - t0) / len(batches), "ratings/batch",
      sum(len(b["rating"]) for b in batches) / len(batches))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for b in batches:
        step(b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))
