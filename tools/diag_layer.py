"""GPU diagnostic: one layer config under each GEMM engine vs the oracle, with error maps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import eagcn_oracle as O
from eagcn_b200 import functional as EF, layers as EL
from eagcn_b200.data import make_batch

dev = torch.device("cuda", 0)


def rng(v):
    if not v: return "-"
    out, s, p = [], v[0], v[0]
    for x in v[1:]:
        if x != p + 1: out.append((s, p)); s = x
        p = x
    out.append((s, p))
    return ",".join(f"{a}-{b}" for a, b in out[:10]) + ("..." if len(out) > 10 else "")


def run(B, dataset, fin, fo, seed, engines):
    batch = make_batch(B, dataset=dataset, seed=seed, kb=30, n_afeat=fin)
    torch.manual_seed(seed)
    layer = EL.GraphConv_Layer(fin, 30, *fo, dropout=0.0, structure="Concate").to(dev)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, prm in layer.named_parameters():
            if n.endswith("att.weight"): prm.copy_(torch.randn(prm.shape, generator=g))
            elif n.endswith("graph_conv.weight"): prm.copy_(torch.randn(prm.shape, generator=g) * (1.0 / fin ** 0.5))
            elif n.endswith("bn.weight"): prm.copy_(torch.rand(prm.shape, generator=g) + 0.5)
            elif n.endswith("bn.bias"): prm.copy_(torch.randn(prm.shape, generator=g) * 0.1)
    layer.train()
    dense = [torch.from_numpy(a) for a in batch.dense()]
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    sd = O.clone_sd({("layer1." + k): v for k, v in sd0.items()}, requires_grad=True)
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    afm_ref = dense[1].clone().requires_grad_(True)
    ref = O.layer_forward(sd, "layer1.", dense[0], afm_ref, codes, True)
    R = torch.randn(ref["x"].shape, generator=torch.Generator().manual_seed(1))
    (ref["x"] * R).sum().backward()
    m = dense[0].max(2).values
    print(f"config B={B} {dataset} fin={fin} fo={fo}: N={batch.N} T={int(m.sum())}")
    for eng in engines:
        EF.set_gemm_engine(eng)
        layer.load_state_dict(sd0)
        for p in layer.parameters(): p.grad = None
        ins = [t.to(dev) for t in dense]
        ins[1].requires_grad_(True)
        EL.GraphConv_Layer._plan_cache = None
        x, _ = layer(*ins)
        (x * R.to(dev)).sum().backward()
        torch.cuda.synchronize()
        ex = (x.detach().cpu() - ref["x"]).abs().max() / ref["x"].abs().max()
        ga = ins[1].grad.cpu(); gr = afm_ref.grad
        err = (ga - gr).abs() / gr.abs().max()
        badf = (err.amax((0, 1)) > 1e-4).nonzero().flatten().tolist()
        bi = (err.amax(2) > 1e-4).nonzero()
        badmol = sorted(set(bi[:, 0].tolist()))
        print(f"  [{eng}] x err {float(ex):.2e}  afm.grad err max {float(err.max()):.2e}  bad feature cols [{rng(badf)}] "
              f"bad molecules [{rng(badmol)}] n_bad_rows {len(bi)}")
        if len(bi):
            b, i = bi[0].tolist()
            print(f"     first bad row (b={b}, i={i}, size={int(batch.sizes[b])}): got {ga[b, i, :4].tolist()} ref {gr[b, i, :4].tolist()}")
        for k, prm in layer.named_parameters():
            rg = sd["layer1." + k].grad
            if rg is None: continue
            e = float((prm.grad.cpu() - rg).abs().max()) / max(float(rg.abs().max()), 1e-12)
            if e > 1e-4 and not k.endswith("graph_conv.bias"):
                print(f"     param {k}: rel err {e:.2e}")


if __name__ == "__main__":
    engs = ["ffma", "tcgen05-nt", "tcgen05"]
    run(48, "tox21", 400, (140,) * 5, 48, engs)
    run(64, "tox21", 24, (80,) * 5, 64, engs)
