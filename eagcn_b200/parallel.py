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
