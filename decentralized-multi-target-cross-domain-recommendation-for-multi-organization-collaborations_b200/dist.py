"""Organization -> rank sharding and the per-round exchange of organization outputs.

The reference passes the K prediction matrices around as Python lists inside one process
(src/train_recsys_assist.py:166-172). Here ``assign_orgs`` spreads the organizations over the ranks (balanced by
count, then by work); rank r's organizations occupy rows [r*c, (r+1)*c) of the organization-major matrix O[split]
([world*c x nnz], c = most organizations any rank holds, unused rows are padding), so after ``predict`` ONE in-place
all-gather per split publishes everybody's rows.
NCCL over NVLink on the GPU box; gloo on CPU for the world_size-2 tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def assign_orgs(costs, world):
    """Balanced organization -> rank map (every rank computes the same one from the same inputs): longest-processing-
    time-first on ``costs`` with the rank's organization COUNT as the first criterion, so no rank stays empty while
    another holds two (contiguous blocks of ceil(K/world) left ranks idle: 18 organizations on 8 ranks -> 3,3,3,3,3,3,0,0).
    Returns (orgs_of_rank: list of ascending lists, chunk = max organizations per rank, org_row: row of the
    rank-blocked matrix O_full [world*chunk x nnz] that holds each organization, rank r's rows being
    [r*chunk, (r+1)*chunk) — the layout one in-place all-gather fills)."""
    K = len(costs)
    order = sorted(range(K), key=lambda k: (-float(costs[k]), k))
    load = [0.0] * world
    count = [0] * world
    mine = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda q: (count[q], load[q], q))
        mine[r].append(k)
        load[r] += float(costs[k])
        count[r] += 1
    mine = [sorted(m) for m in mine]
    chunk = max(1, max(len(m) for m in mine))
    org_row = [0] * K
    for r, m in enumerate(mine):
        for j, k in enumerate(m):
            org_row[k] = r * chunk + j
    return mine, chunk, org_row


def shard_context():
    """(rank, world) of the organization sharding behind the drop-in API: active when torch.distributed is initialised
    with more than one rank (the reference's driver launched under torchrun, one process per GPU), else (0, 1)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 \
            and os.environ.get("DMT_SHARD", "1") != "0":
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def org_block(K, world, rank):
    """Organizations owned by ``rank``: a contiguous block of ceil(K/world) ids (possibly empty at the tail)."""
    c = -(-K // world)
    return list(range(rank * c, min(K, (rank + 1) * c))), c


def exchange_outputs(O_full, chunk, rank, world, group=None):
    """In-place all-gather of every split's organization-major matrix: rank r contributes rows [r*chunk, (r+1)*chunk)."""
    if world == 1:
        return
    for k, O in O_full.items():
        if O.shape[0] != chunk * world:
            raise ValueError("O[{}] must have world*chunk rows for the in-place all-gather".format(k))
        mine = O[rank * chunk:(rank + 1) * chunk]
        dist.all_gather_into_tensor(O, mine, group=group)


def max_over_ranks(value, device):
    """Max of a python float over ranks (device-timed numbers are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
