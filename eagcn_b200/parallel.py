"""Data parallelism over the molecule batch: one process per GPU, parameters
replicated, molecules sharded, ONE flat-buffer gradient all-reduce per step (NCCL over NVLink/NVSwitch
on the GPU box, gloo in the CPU tests), optional global-batch BatchNorm statistics.

The reference has no distributed code at all; its semantics are single-process
big-batch.  ``bn_sync='global'`` reproduces them across ranks: every BatchNorm's per-channel
(sum, sum-of-squares) partials -- and the two backward sums -- are all-reduced, and the population is
the padded size of the *global* batch (``global_population``).  ``bn_sync='local'`` keeps per-replica
statistics (plain data parallelism, what the throughput runs use).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameters that receive gradients share one flat fp32 buffer; ``p.grad`` are views into it.

    Parameters whose gradient is structurally None (layerN.self_r for non-'pool' read-outs,
    ave_A.weight, AFM_BatchNorm.weight/bias -- SURVEY.md 7) are left out, exactly as Adam skips them.
    """

    def __init__(self, params, has_grad=None):
        params = [p for p in params if p.requires_grad]
        if has_grad is not None:
            params = [p for p, h in zip(params, has_grad) if h]
        self.params = params
        n = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @classmethod
    def from_probe(cls, model, run_backward):
        """Run one backward (``run_backward()``) to find which parameters get gradients."""
        for p in model.parameters():
            p.grad = None
        run_backward()
        ps = [p for p in model.parameters() if p.requires_grad]
        has = [p.grad is not None for p in ps]
        return cls(ps, has)

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average=True):
        """One collective over the whole gradient (sum, then 1/world for the mean-over-replicas)."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / world)

    @property
    def nbytes(self):
        return self.flat.numel() * 4


class ArenaAllReduce:
    """The step's ONE gradient exchange over a ``functional.GradArena`` -- no packing copy, the mean over replicas taken
    by the collective itself (ReduceOp.AVG) -- issued on its own stream in (at most) two pieces so that it overlaps the
    backward pass: everything the arena holds when the gradient of ``early`` tensors arrives (``watch``: typically the
    first layer's output, i.e. the head's and the upper layers' gradients, >95 % of the bytes) is reduced while the first
    layer's backward still runs; ``finish()`` reduces the rest and makes the current stream wait.  Captured into a CUDA
    graph the communication stream becomes a parallel branch of the step (NCCL collectives are capturable; the process
    group must be created with TORCH_NCCL_ASYNC_ERROR_HANDLING=0, see bench.py)."""

    def __init__(self, arena, group=None, overlap=True):
        self.arena, self.group, self.overlap = arena, group, overlap
        self.comm = torch.cuda.Stream(device=arena.device) if arena.device.type == "cuda" else None
        self.done = 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def begin(self):
        self.arena.reset()
        self.done = 0

    def _reduce(self, lo, hi):
        if hi <= lo or self.world == 1:
            return
        t = self.arena.flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                                        # gloo (CPU tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.mul_(1.0 / self.world)

    def flush_async(self):
        """Reduce what the arena holds so far on the communication stream (ordered after everything enqueued on the
        current stream and on the side stream of functional.Overlap)."""
        if not self.overlap or self.world == 1:
            return
        from .functional import Overlap
        dev = self.arena.device
        hi = self.arena.off
        if self.comm is None:
            self._reduce(self.done, hi)
        else:
            self.comm.wait_stream(torch.cuda.current_stream(dev))
            if Overlap.enabled:
                self.comm.wait_stream(Overlap.side(dev))
            with torch.cuda.stream(self.comm):
                self._reduce(self.done, hi)
        self.done = hi

    def watch(self, tensor):
        """When the gradient of ``tensor`` has been computed, reduce everything produced so far (tensor hook)."""
        if tensor.requires_grad and self.overlap and self.world > 1:
            tensor.register_hook(lambda g: (self.flush_async(), None)[1])

    def finish(self):
        """Reduce the remainder and join: afterwards every ``p.grad`` (views of the arena) holds the replica mean."""
        if self.world == 1:
            return
        from .functional import Overlap
        dev = self.arena.device
        if self.comm is None:
            self._reduce(self.done, self.arena.off)
        else:
            if Overlap.enabled:
                Overlap.join_pending(dev)            # weight gradients still running on the side stream
            self.comm.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.comm):
                self._reduce(self.done, self.arena.off)
            torch.cuda.current_stream(dev).wait_stream(self.comm)
        self.done = self.arena.off


def shard_balanced(batches, rank, world):
    """Deal the molecules of a global batch (list of MolBatch, one per rank's worth) to ``world`` ranks so that every
    rank gets the same number of molecules and nearly the same number of atoms: molecules sorted by size, dealt in
    snake order.  Synchronous data parallelism runs at the pace of the slowest rank every step; with molecule sizes
    spread from 2 to 132 atoms, contiguous shards differ by several per cent in rows and edges.  The union over ranks
    is the same global batch, so the replica-mean gradient is unchanged."""
    import numpy as np
    from .data import MolBatch, NO_EDGE
    mols = []
    for bi, b in enumerate(batches):
        for m in range(b.B):
            mols.append((int(b.sizes[m]), bi, m))
    order = sorted(range(len(mols)), key=lambda i: (-mols[i][0], mols[i][1], mols[i][2]))
    mine = []
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world)
        r = k if rnd % 2 == 0 else world - 1 - k
        if r == rank:
            mine.append(mols[i])
    mine.sort(key=lambda t: (t[1], t[2]))               # keep the generator's order within a rank
    N = max(t[0] for t in mine)
    B = len(mine)
    V = batches[0].codes.shape[1]
    F = batches[0].afm.shape[2]
    adj = np.zeros((B, N, N), np.float32)
    afm = np.zeros((B, N, F), np.float32)
    codes = np.full((B, V, N, N), NO_EDGE, np.uint8)
    sizes = np.zeros(B, np.int64)
    for j, (n, bi, m) in enumerate(mine):
        src = batches[bi]
        adj[j, :n, :n] = src.adj[m, :n, :n]
        afm[j, :n] = src.afm[m, :n]
        codes[j, :, :n, :n] = src.codes[m, :, :n, :n]
        sizes[j] = n
    return MolBatch(adj=adj, afm=afm, codes=codes, sizes=sizes, channels=batches[0].channels)


def make_stat_allreduce(group=None):
    """Callable for GraphConv_Layer.stat_allreduce: sums the fp64 [2, C] BatchNorm partials over ranks."""
    def _ar(t):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return _ar


def set_bn_sync(model, mode: str, group=None):
    """mode 'local' | 'global' for every GraphConv_Layer of ``model``."""
    from .layers import GraphConv_Layer
    if mode not in ("local", "global"):
        raise ValueError(mode)
    hook = make_stat_allreduce(group) if mode == "global" else None
    for m in model.modules():
        if isinstance(m, GraphConv_Layer):
            m.stat_allreduce = hook
    # the read-out head's three BatchNorm1d (models.py:112,115,119): SyncBatchNorm over the same parameters / buffers
    if hasattr(model, "head_bn"):
        if mode == "global":
            if model.head_bn != "sync":
                model._head_bn_local = model.head_bn
            model.head_bn = "sync"
        elif model.head_bn == "sync":
            model.head_bn = getattr(model, "_head_bn_local", "cuda")


def global_population(local_B: int, n_pad: int, group=None, device=None):
    """(M_total, N_pad) of the global padded batch: sum_r B_r * max_r N_r and max_r N_r.

    Two scalars, reduced on the host side of the step (sizes come from the data loader, so no device
    sync is involved)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_B * n_pad, n_pad
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    t = torch.tensor([local_B, n_pad], dtype=torch.int64, device=dev)
    tb = t.clone()
    dist.all_reduce(tb[:1], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t[1:], op=dist.ReduceOp.MAX, group=group)
    B, N = int(tb[0]), int(t[1])
    return B * N, N


def combine_bn_partials(sum_d, sum_d2, bias, M, eps=1e-5):
    """Host-side statement of the statistic the CUDA path derives from (all-reduced) partials:
    sum_d = sum over active rows of (Y - b), sum_d2 of (Y - b)^2; the M - T padded rows hold Y = b.
    Returns (mean, biased var, invstd).  Used by the gloo tests as the N-rank == 1-rank check."""
    m1 = sum_d / M
    var = (sum_d2 / M - m1 * m1).clamp_min(0)
    return bias + m1, var, 1.0 / torch.sqrt(var + eps)
