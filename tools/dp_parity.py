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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    res = check(dev, rank, world)
    if rank == 0:
        print(f"DP-{world} global-BN vs single process: atoms {res['atoms']:.2e}  outputs {res['outputs']:.2e}  all grads "
              f"{res['all_grads']:.2e}  running stats {res['running_stats']:.2e}  {res['mismatches']}")
        assert res["ok"], "data-parallel parity failed"
        print("DP PARITY OK")
    dist.destroy_process_group()


def check(dev, rank, world):
    """N-rank data parallel (contiguous shards, global-batch BatchNorm in the graph-conv layers AND the read-out head, ONE
    flat gradient all-reduce) against the single process on the concatenated batch, on this process group: the FULL model
    (layers + read-out + head), loss = sum of the outputs.  Returns max relative deviations (max over ranks)."""
    full = make_batch(32 * world, "tox21", seed=7)
    torch.manual_seed(0)
    model = EM.EAGCNStack(30, 24, [(16,) * 5, (24,) * 5], 32, 16, 3, dropout=0.0).to(dev)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model.train()

    def fwd_bwd(mb, m_total=None, n_pad=None):
        for p in model.parameters():
            p.grad = None
        dense = [torch.from_numpy(a).to(dev) for a in mb.dense()]
        plan = GraphPlan.build(dense[0], dense[2:]).check()
        if m_total is not None:
            plan.m_total, plan.n_pad = m_total, n_pad
        out, atoms, _ = model(plan, dense[1], size=torch.from_numpy(mb.sizes).to(dev))
        out.sum().backward()
        named = [(n, p) for n, p in model.named_parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for _, p in named])
        rs = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
        return out.detach(), atoms.materialize().to(dev), flat, rs, named

    # (b) single process, whole batch
    out_full, atoms_full, flat_full, rs_one, named = fwd_bwd(full)

    # (a) data parallel: shard + global-batch statistics everywhere + ONE flat gradient all-reduce (sum over shards)
    model.load_state_dict(sd0)
    PAR.set_bn_sync(model, "global")
    mb = shard(full, rank, world)
    M, Npad = PAR.global_population(mb.B, mb.N, device=dev)
    lo = rank * ((full.B + world - 1) // world)
    out_dp, atoms_dp, flat_dp, rs_dp, _ = fwd_bwd(mb, M, Npad)
    dist.all_reduce(flat_dp)
    PAR.set_bn_sync(model, "local")
    model.load_state_dict(sd0)

    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    eo = rel(out_dp, out_full[lo:lo + mb.B])
    ex = rel(atoms_dp, atoms_full[lo:lo + mb.B])
    eg = rel(flat_dp, flat_full)
    # running means that are analytically 0 (bn_den1: its input is a linear map of Graph_BN's zero-mean output) hold
    # rounding noise ~1e-8: measure running statistics against a floor of 1e-2 (running variances start at 1)
    ers = {k: float((rs_dp[k] - rs_one[k]).abs().max() / rs_one[k].abs().max().clamp_min(1e-2)) for k in rs_one}
    er = max(ers.values())
    worst_rs = sorted(ers.items(), key=lambda kv: -kv[1])[:3]
    res = torch.tensor([ex, eg, er, eo], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    worst = []
    if rank == 0:
        off = 0
        for n, p in named:
            k = p.numel()
            e = float((flat_dp[off:off + k] - flat_full[off:off + k]).abs().max() / flat_full.abs().max())
            if e > 5e-5:
                worst.append((n, e))
            off += k
    a, g, r, o = (float(x) for x in res)
    return {"ranks": world, "global_batch": full.B, "bn_sync": "global (layers + head)", "atoms": a, "outputs": o,
            "all_grads": g, "layer_grads": g, "running_stats": r,
            "ok": bool(a <= 1e-5 and o <= 5e-5 and g <= 5e-5 and r <= 1e-5), "mismatches": worst[:4],
            "worst_running_stats": [(k, float(v)) for k, v in worst_rs],
            "what": "N-rank shards + global-batch BatchNorm (graph-conv layers and the head's three BatchNorm1d) + one flat "
                    "gradient all-reduce vs the single process on the concatenated batch: full model, loss = sum of outputs "
                    "(max relative deviation over ranks)"}


if __name__ == "__main__":
    main()
