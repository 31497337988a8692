"""Helpers to read the golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch
from scipy.sparse import csr_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Fixture:
    def __init__(self, case):
        self.case = case
        self.z = np.load(os.path.join(GOLDEN, case + ".npz"), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files

    def group(self, prefix, as_torch=True):
        out = {}
        p = prefix + "/"
        for k in self.z.files:
            if k.startswith(p):
                v = self.z[k]
                out[k[len(p):]] = torch.from_numpy(np.array(v)) if as_torch else v
        return out

    def csr(self, name):
        return csr_matrix((self.z[name + "/data"], self.z[name + "/indices"], self.z[name + "/indptr"]),
                          shape=tuple(self.z[name + "/shape"]))

    def json(self, key):
        return json.loads(str(self.z[key]))


def cases(kind):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(kind + "_") and f.endswith(".npz"))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max())) if a.size else 0.0
