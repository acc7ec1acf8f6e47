"""Synthetic molecular-batch generator and collate layouts.

Produces batches in exactly the layout the reference's collate functions hand to the
model (reference eagcn_pytorch/utils.py:528-566 / :594-633): all ``float32``, zero-padded
to the *batch* maximum atom count N:

    adj      [B, N, N]        0/1, symmetric, zero diagonal   (neural_fp.py:85,109-110)
    afm      [B, N, 24]       atom features, 0 on padding      (neural_fp.py:312-333)
    TypeAtt  [B, Kb, N, N]    one-hot over bond element-pair types on bonded pairs
    OrderAtt [B, 4, N, N]     one-hot bond order                (neural_fp.py:214)
    AromAtt / ConjAtt / RingAtt [B, 2, N, N]                    (neural_fp.py:215-217)
    size     [B] int64

and, for the packed data boundary, the same batch as uint8 edge
codes ``codes [B, 5, N, N]`` (255 = no bond) from which the dense one-hot tensors can be
re-expanded bit-exactly.

Everything here is numpy and deterministic in ``seed``; nothing needs a GPU.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

VIEW_CHANNELS_TAIL = (4, 2, 2, 2)  # reference layers.py:270-273 hard-codes C_2..C_5
N_AFEAT = 24                        # reference utils.py:530 / neural_fp.py:312-333
NO_EDGE = 255

# per-dataset layer widths and Kb (train.py:61-114; Kb: check_model.py:20-35, 30 = placeholder)
DATASETS = {
    "tox21":    dict(kb=30, sgc1=80,  sgc2=140, den=(256, 64), nclass=12),
    "hiv":      dict(kb=30, sgc1=100, sgc2=250, den=(512, 128), nclass=1),
    "lipo":     dict(kb=18, sgc1=60,  sgc2=100, den=(128, 64), nclass=1),
    "freesolv": dict(kb=17, sgc1=40,  sgc2=60,  den=(128, 64), nclass=1),
}

_HIST = None


def size_histogram(dataset: str):
    """(sizes, probabilities) of heavy-atom counts for ``dataset`` (tools/make_size_hist.py)."""
    global _HIST
    if _HIST is None:
        with open(os.path.join(os.path.dirname(__file__), "size_hist.json")) as f:
            _HIST = json.load(f)
    h = _HIST[dataset]
    sizes = np.array([int(k) for k in h], dtype=np.int64)
    cnt = np.array([h[k] for k in h], dtype=np.float64)
    return sizes, cnt / cnt.sum()


def view_channels(kb: int, n_views: int = 5):
    """C_v per view: (Kb, 4, 2, 2, 2), cycled for K != 5."""
    base = (kb,) + VIEW_CHANNELS_TAIL
    return tuple(base[v % 5] for v in range(n_views))


@dataclass
class MolBatch:
    adj: np.ndarray          # [B,N,N] f32
    afm: np.ndarray          # [B,N,F] f32
    codes: np.ndarray        # [B,V,N,N] u8, NO_EDGE off-graph
    sizes: np.ndarray        # [B] i64
    channels: tuple          # C_v per view

    @property
    def B(self):
        return self.adj.shape[0]

    @property
    def N(self):
        return self.adj.shape[1]

    def rel(self, v: int) -> np.ndarray:
        """Dense one-hot relation tensor of view ``v``: [B, C_v, N, N] float32."""
        return expand_onehot(self.codes[:, v], self.channels[v])

    def dense(self):
        """(adj, afm, rel_1..rel_V) -- the tensors GraphConv_Layer.forward takes."""
        return (self.adj, self.afm) + tuple(self.rel(v) for v in range(len(self.channels)))

    def dense_bytes(self) -> int:
        B, N = self.B, self.N
        return 4 * (B * N * N * (1 + sum(self.channels)) + B * N * self.afm.shape[2])

    def packed_bytes(self) -> int:
        return self.codes.nbytes + self.afm.nbytes + self.sizes.nbytes


def expand_onehot(code: np.ndarray, C: int) -> np.ndarray:
    """uint8 codes [B,N,N] -> one-hot float32 [B,C,N,N]; code >= C (e.g. NO_EDGE) -> all zero."""
    B, N, _ = code.shape
    out = np.zeros((B, C, N, N), dtype=np.float32)
    b, i, j = np.nonzero(code < C)
    out[b, code[b, i, j], i, j] = 1.0
    return out


def _one_graph(rng: np.random.Generator, n: int):
    """chain 0-1-...-(n-1) plus floor(n/6) random chords -> list of undirected edges."""
    edges = {(k, k + 1) for k in range(n - 1)}
    for _ in range(n // 6):
        a, b = rng.integers(0, n, size=2)
        if a != b:
            edges.add((min(a, b), max(a, b)))
    return sorted(edges)


def make_batch(batch: int, dataset: str = "tox21", seed: int = 0, kb: int | None = None,
               n_views: int = 5, fixed_n: int | None = None, n_afeat: int = N_AFEAT,
               pad_to: int | None = None) -> MolBatch:
    """Deterministic synthetic batch.

    dataset-shaped sizes unless ``fixed_n`` is given (sweep config).  ``pad_to`` forces the
    padded width N (>= batch max) -- used to emulate padding to a *global* batch maximum.
    """
    rng = np.random.default_rng(seed)
    if kb is None:
        kb = DATASETS[dataset]["kb"]
    chans = view_channels(kb, n_views)
    if fixed_n is not None:
        sizes = np.full(batch, fixed_n, dtype=np.int64)
    else:
        s, p = size_histogram(dataset)
        sizes = rng.choice(s, size=batch, replace=True, p=p).astype(np.int64)
    N = int(sizes.max())
    if pad_to is not None:
        assert pad_to >= N
        N = pad_to
    adj = np.zeros((batch, N, N), dtype=np.float32)
    afm = np.zeros((batch, N, n_afeat), dtype=np.float32)
    codes = np.full((batch, n_views, N, N), NO_EDGE, dtype=np.uint8)
    for b in range(batch):
        n = int(sizes[b])
        ed = _one_graph(rng, n)
        if ed:
            e = np.array(ed, dtype=np.int64)
            adj[b, e[:, 0], e[:, 1]] = 1.0
            adj[b, e[:, 1], e[:, 0]] = 1.0
            for v in range(n_views):
                c = rng.integers(0, chans[v], size=len(ed)).astype(np.uint8)
                codes[b, v, e[:, 0], e[:, 1]] = c
                codes[b, v, e[:, 1], e[:, 0]] = c
        afm[b, :n] = rng.random((n, n_afeat), dtype=np.float32)
    return MolBatch(adj=adj, afm=afm, codes=codes, sizes=sizes, channels=chans)


def shard(batch: MolBatch, rank: int, world: int) -> MolBatch:
    """Contiguous shard of the molecule batch for data parallelism.

    The padded width N of the *global* batch is kept so padded-row BatchNorm accounting can
    follow the single-process semantics when ``bn_sync='global'``.
    """
    B = batch.B
    per = (B + world - 1) // world
    lo, hi = rank * per, min(B, (rank + 1) * per)
    return MolBatch(adj=batch.adj[lo:hi], afm=batch.afm[lo:hi], codes=batch.codes[lo:hi],
                    sizes=batch.sizes[lo:hi], channels=batch.channels)
