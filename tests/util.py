"""Shared helpers for the test-suite (test infrastructure)."""
import glob
import os

import numpy as np
import torch

from eagcn_b200.data import MolBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


class Golden:
    def __init__(self, case):
        z = np.load(os.path.join(GOLDEN, case + ".npz"), allow_pickle=False)
        self.z = z
        self.batch = MolBatch(adj=z["in.adj"].astype(np.float32), afm=z["in.afm"], codes=z["in.codes"],
                              sizes=z["in.sizes"], channels=tuple(int(c) for c in z["in.channels"]))
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
        self.out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out.")}
        self.cot = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cot.")}
        self.grad = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
        self.post = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("post.")}
        self.meta = {k[5:]: z[k] for k in z.files if k.startswith("meta.")}

    def dense(self):
        return [torch.from_numpy(a) for a in self.batch.dense()]

    def codes_i64(self):
        """per-view int64 codes, -1 off-graph (oracle convention)."""
        c = torch.from_numpy(self.batch.codes.astype(np.int64))
        c = torch.where(c == 255, torch.full_like(c, -1), c)
        return [c[:, v] for v in range(c.shape[1])]


def rel_err(x, ref):
    denom = float(ref.abs().max())
    err = float((x.detach().double() - ref.detach().double()).abs().max())
    return err if denom == 0.0 else err / denom
