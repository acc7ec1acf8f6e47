"""CPU, world_size 2, gloo: the host-side data-parallel logic -- molecule sharding, the single
flat-buffer gradient all-reduce, and the global-batch BatchNorm statistic combination -- checked against the
single-process oracle on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eagcn_b200 import parallel as PAR
from eagcn_b200.data import make_batch, shard
from oracle import eagcn_oracle as O
from tests.util import Golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


# ---------------------------------------------------------------- flat gradient bucket
def _flat_grad(rank, world):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    frozen = torch.nn.Parameter(torch.zeros(4))                    # never receives a gradient
    params = list(model.parameters()) + [frozen]
    X = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    lo, hi = rank * 4, rank * 4 + 4

    class M(torch.nn.Module):
        def parameters(self, recurse=True):
            return iter(params)
    holder = M()
    bucket = PAR.FlatGradBucket.from_probe(holder, lambda: model(X[lo:hi]).sum().backward())
    assert len(bucket.params) == 4 and bucket.flat.numel() == sum(p.numel() for p in model.parameters())
    bucket.zero()
    model(X[lo:hi]).sum().backward()                               # accumulates in place into the views
    bucket.all_reduce(average=False)
    return bucket.flat.clone().numpy()


def test_flat_grad_allreduce_equals_full_batch():
    out = _spawn(_flat_grad)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    X = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    model(X).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    np.testing.assert_allclose(out[0], ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(out[0], out[1])


# ---------------------------------------------------------------- global-batch BatchNorm statistics
def _bn_partials(rank, world):
    g = Golden("layer_wide")
    full = g.batch
    sh = shard(full, rank, world)                                   # keeps the global padded width N
    assert sh.N == full.N
    M, Npad = PAR.global_population(sh.B, sh.N)
    assert (M, Npad) == (full.B * full.N, full.N)
    sd = O.clone_sd({("layer1." + k): v for k, v in g.sd.items()})
    dense = [torch.from_numpy(a) for a in sh.dense()]
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    out = O.layer_forward(sd, "layer1.", dense[0], dense[1], codes, True)
    m = O.row_mask(dense[0]).unsqueeze(2)
    res = {}
    for v in range(5):
        b = sd[f"layer1.block{v + 1}.graph_conv.bias"]
        D = (out["Y"][v] - b) * m                                   # what the CUDA agg kernel accumulates
        sums = torch.stack([D.sum((0, 1)), (D * D).sum((0, 1))]).double()
        PAR.make_stat_allreduce()(sums)
        mean, var, invstd = PAR.combine_bn_partials(sums[0], sums[1], b.double(), float(M))
        res[v] = (mean.numpy(), var.numpy())
    return res


def test_global_bn_statistics_match_single_process():
    out = _spawn(_bn_partials)
    g = Golden("layer_wide")
    sd = O.clone_sd({("layer1." + k): v for k, v in g.sd.items()})
    dense = g.dense()
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    ref = O.layer_forward(sd, "layer1.", dense[0], dense[1], codes, True)
    for v in range(5):
        Y = ref["Y"][v].reshape(-1, ref["Y"][v].shape[2])
        mean, var = Y.mean(0).numpy(), Y.var(0, unbiased=False).numpy()
        for r in (0, 1):
            np.testing.assert_allclose(out[r][v][0], mean, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out[r][v][1], var, rtol=1e-4, atol=1e-7)


def test_shard_partitions_batch():
    b = make_batch(10, "freesolv", seed=3)
    parts = [shard(b, r, 4) for r in range(4)]
    assert sum(p.B for p in parts) == 10
    assert np.array_equal(np.concatenate([p.adj for p in parts]), b.adj)
    assert all(p.N == b.N for p in parts)


# ---------------------------------------------------------------- gradient arena + two-piece all-reduce
def _arena_allreduce(rank, world):
    from eagcn_b200.functional import GradArena
    arena = GradArena(200, "cpu")
    ar = PAR.ArenaAllReduce(arena, overlap=True)
    out = []
    for step in range(2):                                          # offsets repeat every step
        ar.begin()
        a = arena.take(10); a.fill_(float(rank + 1))               # "head + upper layers"
        b = arena.take(33); b.copy_(torch.arange(33.0) * (rank + 1))
        assert a.data_ptr() == arena.flat.data_ptr() and b.data_ptr() == arena.flat[16:].data_ptr()   # 64-byte aligned slices
        ar.flush_async()                                           # first piece goes out while "layer 1" still computes
        c = arena.take(5); c.fill_(10.0 * (rank + 1))
        ar.finish()
        assert arena.holds(a) and arena.holds(c) and not arena.holds(torch.zeros(3))
        out.append((a.clone().numpy(), b.clone().numpy(), c.clone().numpy()))
    return out


def test_arena_allreduce_is_replica_mean():
    out = _spawn(_arena_allreduce)
    for r in (0, 1):
        for a, b, c in out[r]:
            np.testing.assert_allclose(a, np.full(10, 1.5))
            np.testing.assert_allclose(b, np.arange(33.0) * 1.5)
            np.testing.assert_allclose(c, np.full(5, 15.0))


def test_grad_arena_measure_and_overflow():
    from eagcn_b200.functional import GradArena
    from eagcn_b200._lib import EagcnError
    n = GradArena.measure(lambda: (GradArena.empty(10, "cpu"), GradArena.empty(20, "cpu")), "cpu")
    assert n == 16 + 32 and GradArena.current is None
    arena = GradArena(n, "cpu")
    GradArena.current = arena
    try:
        x, y = GradArena.empty(10, "cpu"), GradArena.empty(20, "cpu")
        assert arena.holds(x) and arena.holds(y) and arena.off == 48
        with pytest.raises(EagcnError):
            GradArena.empty(1, "cpu")
    finally:
        GradArena.current = None


def test_shard_balanced_same_molecules_even_atoms():
    world = 4
    batches = [make_batch(32, "tox21", seed=10 + r) for r in range(world)]
    parts = [PAR.shard_balanced(batches, r, world) for r in range(world)]
    assert all(p.B == 32 for p in parts)
    all_sizes = np.sort(np.concatenate([b.sizes for b in batches]))
    assert np.array_equal(np.sort(np.concatenate([p.sizes for p in parts])), all_sizes)      # same multiset of molecules
    atoms = [int(p.sizes.sum()) for p in parts]
    contiguous = [int(b.sizes.sum()) for b in batches]
    assert max(atoms) - min(atoms) <= max(8, (max(contiguous) - min(contiguous)) // 4)
    # a molecule keeps its graph, features and codes
    p0 = parts[0]
    n = int(p0.sizes[0])
    found = False
    for b in batches:
        for m in range(b.B):
            if int(b.sizes[m]) == n and np.array_equal(b.adj[m, :n, :n], p0.adj[0, :n, :n]) and \
                    np.array_equal(b.afm[m, :n], p0.afm[0, :n]) and np.array_equal(b.codes[m, :, :n, :n], p0.codes[0, :, :n, :n]):
                found = True
    assert found
    # one rank: the identity
    one = PAR.shard_balanced(batches[:1], 0, 1)
    assert np.array_equal(one.adj, batches[0].adj) and np.array_equal(one.codes, batches[0].codes)


def test_deferred_join_needs_a_single_gradient_producer():
    """functional.Overlap.fresh: the join of a side-stream weight-gradient branch may be deferred to the end of the backward
    pass only when autograd adopts the returned tensors untouched.  A parameter with TWO gradient producers (the attention
    weights of a layer whose dense attention is also returned -- 'pool' read-out, models.py:104-106 -- or tied weights) is
    summed by the engine when the second gradient arrives, so it must not be deferred.  Pure bookkeeping: CPU tensors."""
    import torch
    from eagcn_b200.functional import Overlap
    a = torch.nn.Parameter(torch.zeros(3))
    b = torch.nn.Parameter(torch.zeros(3))
    c = torch.nn.Parameter(torch.zeros(3))
    frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
    Overlap._uses.clear(); Overlap._bwd_seen = False
    with torch.no_grad():
        assert not Overlap.fresh((a,))                       # never registered: not deferred
        Overlap.note_use((a, b, frozen, None))
        Overlap.note_use((b,))                               # b has a second producer (e.g. attention_dense)
        assert Overlap.fresh((a, frozen, None))
        assert not Overlap.fresh((a, b)) and not Overlap.fresh((b,))
        a.grad = torch.zeros(3)                              # a gradient is already there: the engine will accumulate
        assert not Overlap.fresh((a,))
        a.grad = None
        Overlap.note_backward()                              # a backward pass ran ...
        Overlap.note_use((b, c))                             # ... the next forward starts counting afresh
        assert Overlap.fresh((b, c)) and not Overlap.fresh((a,))
        Overlap.note_use((c,))                               # two forward passes before one backward: both count
        assert not Overlap.fresh((c,)) and Overlap.fresh((b,))
    assert not Overlap.fresh((b,))                           # grad mode on (create_graph): never deferred
    Overlap._uses.clear(); Overlap._bwd_seen = False
