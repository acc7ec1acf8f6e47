"""2-GPU check of SURVEY.md 8(e): N-rank data parallel with bn_sync='global' == single-process big batch.

    torchrun --nproc-per-node 2 tools/dp_parity.py

Every rank builds the same global batch, runs (a) its contiguous shard under global-batch BatchNorm with the
flat-gradient all-reduce, (b) the whole batch alone; prints max relative deviations of outputs and gradients.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from eagcn_b200 import models as EM, parallel as PAR
from eagcn_b200.data import make_batch, shard
from eagcn_b200.plan import GraphPlan


def run(model, mb, dev, m_total=None, n_pad=None):
    dense = [torch.from_numpy(a).to(dev) for a in mb.dense()]
    plan = GraphPlan.build(dense[0], dense[2:]).check()
    if m_total is not None:
        plan.m_total, plan.n_pad = m_total, n_pad
    out, _, _ = model(plan, dense[1], size=torch.from_numpy(mb.sizes).to(dev))
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    res = check(dev, rank, world)
    if rank == 0:
        print(f"DP-{world} global-BN vs single process: atoms {res['atoms']:.2e}  layer grads {res['layer_grads']:.2e}  "
              f"running stats {res['running_stats']:.2e}")
        assert res["ok"], "data-parallel parity failed"
        print("DP PARITY OK")
    dist.destroy_process_group()


def check(dev, rank, world):
    """N-rank data parallel (contiguous shards, global-batch BatchNorm, ONE flat gradient all-reduce) against the single
    process on the concatenated batch, on this process group.  Returns max relative deviations (max over ranks)."""
    full = make_batch(32 * world, "tox21", seed=7)
    torch.manual_seed(0)
    model = EM.EAGCNStack(30, 24, [(16,) * 5, (24,) * 5], 32, 16, 3, dropout=0.0).to(dev)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model.train()

    # (b) single process, whole batch (head BatchNorms see the whole batch)
    for p in model.parameters(): p.grad = None
    out_full = run(model, full, dev)
    out_full.sum().backward()
    g_full = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    rs_full = {k: v.clone() for k, v in model.state_dict().items() if "running" in k and "layer" in k}

    # (a) data parallel: shard + global-batch statistics in the graph-conv layers.  The dense head's
    # nn.BatchNorm1d would need SyncBatchNorm for exact equality, so compare the atom representations
    # (pure hot-path output) and the hot-path parameter gradients driven by a sum-of-atoms loss.
    model.load_state_dict(sd0)
    PAR.set_bn_sync(model, "global")
    mb = shard(full, rank, world)
    M, Npad = PAR.global_population(mb.B, mb.N, device=dev)
    lo = rank * ((full.B + world - 1) // world)

    def atoms_loss(m, batch, m_total=None, n_pad=None):
        from eagcn_b200 import functional as EF
        from eagcn_b200.layers import PackedRows
        dense = [torch.from_numpy(a).to(dev) for a in batch.dense()]
        plan = GraphPlan.build(dense[0], dense[2:]).check()
        if m_total is not None:
            plan.m_total, plan.n_pad = m_total, n_pad
        h = PackedRows(EF.gather_rows(plan, dense[1]), plan)
        for layer in m.conv_layers:
            h, _ = layer(plan, h)
        return h.dense()

    for p in model.parameters(): p.grad = None
    x_dp = atoms_loss(model, mb, M, Npad)
    w = torch.linspace(0.5, 1.5, x_dp.shape[2], device=dev)
    (x_dp * w).sum().backward()
    params = [p for n, p in model.named_parameters() if n.startswith("layer") and p.grad is not None]
    bucket_flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(bucket_flat)                                  # ONE flat gradient all-reduce (sum over shards)
    rs_dp = {k: v.clone() for k, v in model.state_dict().items() if "running" in k and "layer" in k}

    model.load_state_dict(sd0)
    PAR.set_bn_sync(model, "local")
    for p in model.parameters(): p.grad = None
    x_full = atoms_loss(model, full)
    (x_full * w).sum().backward()
    flat_full = torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters() if n.startswith("layer") and p.grad is not None])
    rs_one = {k: v.clone() for k, v in model.state_dict().items() if "running" in k and "layer" in k}

    ex = float((x_dp - x_full[lo:lo + mb.B]).abs().max() / x_full.abs().max())
    eg = float((bucket_flat - flat_full).abs().max() / flat_full.abs().max())
    er = max(float((rs_dp[k] - rs_one[k]).abs().max() / rs_one[k].abs().max().clamp_min(1e-12)) for k in rs_one)
    res = torch.tensor([ex, eg, er], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    worst = []
    if rank == 0:
        names = [n for n, p in model.named_parameters() if n.startswith("layer") and p.grad is not None]
        off = 0
        for n, p in zip(names, params):
            k = p.numel()
            e = float((bucket_flat[off:off + k] - flat_full[off:off + k]).abs().max() / flat_full.abs().max())
            if e > 5e-5:
                worst.append((n, e))
            off += k
    PAR.set_bn_sync(model, "local")
    a, g, r = float(res[0]), float(res[1]), float(res[2])
    return {"ranks": world, "global_batch": full.B, "bn_sync": "global", "atoms": a, "layer_grads": g, "running_stats": r,
            "ok": bool(a <= 1e-5 and g <= 5e-5 and r <= 1e-5), "mismatches": worst[:4],
            "what": "N-rank shards + global-batch BatchNorm + one flat gradient all-reduce vs the single process on the "
                    "concatenated batch (max relative deviation over ranks)"}


if __name__ == "__main__":
    main()
