"""CPU: the packed collate (eagcn_b200/collate.py) against the dense layout of the reference's collate functions
(utils.py:504-640) -- bit-exact indexing, validation, and (build container) the live reference collate."""
import numpy as np
import pytest
import torch

from eagcn_b200 import collate as C
from eagcn_b200.data import make_batch


def _items(batch):
    """Per-molecule tuples in the order of MolDataset.__getitem__ (utils.py:478-502), cut to each molecule's size."""
    dense = batch.dense()
    items = []
    for b in range(batch.B):
        n = int(batch.sizes[b])
        items.append((dense[0][b, :n, :n].copy(), dense[1][b, :n].copy(),
                      *[dense[2 + v][b, :, :n, :n].copy() for v in range(5)],
                      np.float32(b % 2), "smiles%d" % b, np.ones((n, 1), np.float32) * b, b))
    return items


@pytest.mark.parametrize("dataset,B", [("freesolv", 9), ("tox21", 12)])
def test_codes_from_dense_bit_exact(dataset, B):
    batch = make_batch(B, dataset, seed=3)
    codes, channels = C.codes_from_dense(batch.adj, batch.dense()[2:])
    assert channels == tuple(batch.channels)
    assert codes.dtype == np.uint8 and np.array_equal(codes, batch.codes)
    codes_t, _ = C.codes_from_dense(torch.from_numpy(batch.adj), [torch.from_numpy(r) for r in batch.dense()[2:]])
    assert np.array_equal(codes_t, batch.codes)


def test_packed_collate_equals_dense_collate():
    batch = make_batch(10, "lipo", seed=4)
    out = C.mol_collate_func_packed(_items(batch))
    assert np.array_equal(out["codes"], batch.codes)             # same padding (batch maximum), same codes
    assert np.array_equal(out["afm"], batch.afm)
    assert np.array_equal(out["size"], batch.sizes)
    assert out["channels"] == tuple(batch.channels)
    assert out["subtype"].shape == (10, batch.N, 1) and out["labels"].shape == (10,)
    assert out["codes"].nbytes * 20 < sum(a.nbytes for a in batch.dense())


def test_all_zero_relation_vector_and_off_graph_values():
    batch = make_batch(4, "freesolv", seed=5)
    adj, rels = batch.adj, [r.copy() for r in batch.dense()[2:]]
    b, i, j = [int(x[0]) for x in np.nonzero(adj)]
    rels[1][b, :, i, j] = 0.0                                    # bonded pair without a bond-order channel
    zb, zi, zj = [int(x[0]) for x in np.nonzero(adj == 0)]
    rels[2][zb, 1, zi, zj] = 7.0                                 # garbage on a non-bonded pair: multiplied by adj == 0
    codes, channels = C.codes_from_dense(adj, rels)
    assert codes[b, 1, i, j] == channels[1]
    assert codes[zb, 2, zi, zj] == C.NO_EDGE
    ref = batch.codes.copy()
    ref[b, 1, i, j] = channels[1]
    assert np.array_equal(codes, ref)


def test_malformed_inputs_rejected():
    batch = make_batch(3, "freesolv", seed=6)
    adj, rels = batch.adj, [r.copy() for r in batch.dense()[2:]]
    b, i, j = [int(x[0]) for x in np.nonzero(adj)]
    bad = [r.copy() for r in rels]
    bad[0][b, :, i, j] = 0.0
    bad[0][b, 0, i, j] = bad[0][b, 1, i, j] = 1.0                # two channels set
    with pytest.raises(ValueError):
        C.codes_from_dense(adj, bad)
    bad = [r.copy() for r in rels]
    bad[3][b, 0, i, j] = 0.5                                     # not 0/1
    bad[3][b, 1, i, j] = 0.0
    with pytest.raises(ValueError):
        C.codes_from_dense(adj, bad)
    a2 = adj.copy()
    a2[b, i, j] = 2.0
    with pytest.raises(ValueError):
        C.codes_from_dense(a2, rels)
    with pytest.raises(ValueError):
        C.codes_from_dense(adj, [rels[0][:, :, :-1]] + rels[1:])


@pytest.mark.reference
def test_against_live_reference_collate():
    """The reference's own mol_collate_func_reg on the same items: its dense tensors expand from our codes bit-exactly."""
    from oracle import ref_loader
    from eagcn_b200.data import expand_onehot
    _, _, U = ref_loader.load()
    batch = make_batch(7, "tox21", seed=8)
    items = _items(batch)
    ref = U.mol_collate_func_reg(items)
    out = C.mol_collate_func_packed(items)
    assert np.array_equal(ref[1].cpu().numpy(), out["afm"])
    assert np.array_equal(ref[9].cpu().numpy(), out["size"])
    adj_ref = ref[0].cpu().numpy()
    assert np.array_equal((out["codes"][:, 0] != C.NO_EDGE).astype(np.float32), adj_ref)
    for v in range(5):
        assert np.array_equal(expand_onehot(out["codes"][:, v], out["channels"][v]), ref[2 + v].cpu().numpy())
