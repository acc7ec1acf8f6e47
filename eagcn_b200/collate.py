"""Packed data boundary, host side: what the reference's collate functions would emit if they
did not expand every bond relation into dense one-hot fp32 planes.

The reference pads each molecule's ``adj`` [n,n], ``afm`` [n,24] and five one-hot relation tensors [C_v,n,n] to the
batch maximum and stacks them (utils.py:504-573 ``mol_collate_func_reg``, :575-640 ``mol_collate_func_class``): for a
Tox21 batch of 256 that is ~365 MB of float32 of which 97 % is padding and the rest is one 1.0 per bonded pair and view.
``mol_collate_func_packed`` takes the SAME per-molecule tuples (the items of the reference's ``MolDataset``,
utils.py:478-502) and emits one uint8 code per atom pair and view instead -- 13 MB -- in the layout
``GraphPlan.from_codes`` / ``EAGCN.forward(plan, afm, size=...)`` consume:

    code[b, v, i, j] = c          the single channel with rel_v[b, c, i, j] == 1     (bonded pair)
                     = C_v        bonded pair whose relation vector is all zero       (the 1x1 conv then scores 0)
                     = 255        no bond (adj == 0)

``codes_from_dense`` converts already collated dense tensors (numpy or CPU torch) the same way.  Both validate what the
packed form assumes (0/1 adjacency, at most one 1.0 per bonded pair and view) and raise ``ValueError`` otherwise; a
relation value on a non-bonded pair is ignored, exactly as the reference multiplies it by ``adj == 0``
(layers.py:83).  Pure numpy: this runs in DataLoader workers like the reference's collate."""
from __future__ import annotations

import numpy as np

NO_EDGE = 255


def _codes_one(adj: np.ndarray, rel: np.ndarray) -> np.ndarray:
    """adj [..., n, n], rel [..., C, n, n] -> uint8 codes [..., n, n]."""
    C = rel.shape[-3]
    if C > 254:
        raise ValueError(f"a relation tensor has {C} channels; the packed form holds at most 254")
    bonded = adj != 0
    if not np.all((adj == 0) | (adj == 1)):
        raise ValueError("adjacency holds values other than 0.0 / 1.0")
    nz = rel != 0
    on_bond = nz & np.expand_dims(bonded, -3)
    if np.any(on_bond & (rel != 1)):
        raise ValueError("a relation tensor is not 0/1 on a bonded pair")
    cnt = on_bond.sum(axis=-3)
    if np.any(cnt > 1):
        raise ValueError("a relation tensor is not one-hot on a bonded pair")
    code = np.where(cnt == 1, on_bond.argmax(axis=-3), C)
    return np.where(bonded, code, NO_EDGE).astype(np.uint8)


def codes_from_dense(adj, rels):
    """Collated dense tensors -> codes uint8 [B, V, N, N] and the channel counts (C_1..C_V).

    adj [B,N,N]; rels: sequence of V one-hot tensors [B,C_v,N,N] (TypeAtt, OrderAtt, AromAtt, ConjAtt, RingAtt)."""
    a = np.asarray(adj)
    planes = [np.asarray(r) for r in rels]
    for r in planes:
        if r.ndim != 4 or r.shape[0] != a.shape[0] or r.shape[2:] != a.shape[1:]:
            raise ValueError(f"relation tensor {r.shape} does not match adjacency {a.shape}")
    codes = np.stack([_codes_one(a, r) for r in planes], axis=1)
    return codes, tuple(int(r.shape[1]) for r in planes)


def mol_collate_func_packed(batch):
    """Drop-in for ``mol_collate_func_reg`` / ``mol_collate_func_class`` (utils.py:504-640) on the same items
    ``(adj, afm, TypeAtt, orderAtt, aromAtt, conjAtt, ringAtt, label, smile, subtype, index)``.

    Returns ``dict(codes u8 [B,5,N,N], channels, afm f32 [B,N,F], labels, subtype f32 [B,N,1], size i64 [B],
    index)`` padded to the largest molecule of the batch like the reference (utils.py:524,590); the dense one-hot planes
    are never materialised.  ``labels`` is ``np.array(label_list)`` as in the reference."""
    sizes = np.array([d[0].shape[0] for d in batch], dtype=np.int64)
    N, B = int(sizes.max()), len(batch)
    F = batch[0][1].shape[1]
    channels = tuple(int(batch[0][2 + v].shape[0]) for v in range(5))
    codes = np.full((B, 5, N, N), NO_EDGE, dtype=np.uint8)
    afm = np.zeros((B, N, F), dtype=np.float32)
    subtype = np.zeros((B, N, 1), dtype=np.float32)
    for b, d in enumerate(batch):
        n = int(sizes[b])
        a = np.asarray(d[0])
        for v in range(5):
            r = np.asarray(d[2 + v])
            if r.shape[0] != channels[v]:
                raise ValueError("molecules of one batch disagree on the channel count of a relation tensor")
            codes[b, v, :n, :n] = _codes_one(a, r)
        afm[b, :n] = d[1]
        subtype[b, :n] = d[9]
    return {"codes": codes, "channels": channels, "afm": afm, "labels": np.array([d[7] for d in batch]),
            "subtype": subtype, "size": sizes, "index": np.array([d[10] for d in batch])}
