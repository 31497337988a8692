"""Evaluation with the reference's batching semantics (parity instrument, not a hot path).

reference src/metrics/metrics.py:8-11 (RMSE per block), :63-84 (NDCG@10 on a densified block: unobserved scores
-inf, unobserved gains 0), src/logger.py:35-55 (n-weighted running mean of the per-block values).
"""
from collections import defaultdict
from numbers import Number

import torch

from .config import cfg


def rmse(output, target):
    return float(((output - target) ** 2).mean().sqrt())


def ndcg(output, target, user, item, topk=10):
    rows, cols = (user, item) if cfg['data_mode'] == 'user' else (item, user)
    _, ri = torch.unique(rows, return_inverse=True)
    _, ci = torch.unique(cols, return_inverse=True)
    nr, nc = int(ri.max()) + 1, int(ci.max()) + 1
    score = torch.full((nr, nc), -float('inf'), device=output.device)
    gain = torch.zeros(nr, nc, device=output.device)
    score[ri, ci] = output
    gain[ri, ci] = target
    k = min(topk, nc)
    disc = 1.0 / torch.log2(torch.arange(1, k + 1, dtype=torch.float32, device=output.device) + 1)
    dcg = (gain.gather(1, score.topk(k, dim=-1).indices) * disc).sum(-1)
    idcg = (gain.topk(k, dim=-1).values * disc).sum(-1)
    return float(torch.nan_to_num(dcg / idcg, nan=0.0, posinf=0.0, neginf=0.0).mean())


class Metric:
    def __init__(self, metric_name):
        self.metric_name = metric_name
        explicit = cfg['target_mode'] == 'explicit'
        self.pivot = float('inf') if explicit else -float('inf')
        self.pivot_name = 'RMSE' if explicit else 'NDCG'
        self.pivot_direction = 'down' if explicit else 'up'
        self.metric = {
            'Loss': lambda i, o: float(o['loss']),
            'RMSE': lambda i, o: rmse(o['target_rating'], i['target_rating']),
            'NDCG': lambda i, o: ndcg(o['target_rating'], i['target_rating'], i['target_user'], i['target_item']),
        }

    def evaluate(self, metric_names, input, output):
        return {n: self.metric[n](input, output) for n in metric_names}

    def compare(self, val):
        return self.pivot > val if self.pivot_direction == 'down' else self.pivot < val

    def update(self, val):
        self.pivot = val


class Logger:
    """n-weighted running means per tag (the part of reference src/logger.py the hot path writes to)."""

    def __init__(self):
        self.counter = defaultdict(int)
        self.mean = defaultdict(int)
        self.history = defaultdict(list)

    def safe(self, write):
        if not write:
            for name in self.mean:
                self.history[name].append(self.mean[name])

    def reset(self):
        self.counter = defaultdict(int)
        self.mean = defaultdict(int)

    def append(self, result, tag, n=1, mean=True):
        if not mean:
            return
        for k, v in result.items():
            if isinstance(v, Number):
                name = '{}/{}'.format(tag, k)
                self.counter[name] += n
                self.mean[name] = ((self.counter[name] - n) * self.mean[name] + n * v) / self.counter[name]
