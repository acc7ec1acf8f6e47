"""Generate golden vectors FROM THE UNMODIFIED REFERENCE classes (run in the build container only).

    python tests/golden/make_golden.py [case ...]

Imports /root/reference/eagcn_pytorch/{layers,models,utils}.py through oracle/ref_loader.py,
runs them (CPU, fp32) on small seeded synthetic batches (eagcn_b200.data.make_batch) and writes
tests/golden/<case>.npz holding inputs (compact: uint8 adjacency + uint8 edge codes), every
state_dict tensor, the forward outputs and the autograd gradients.  These fixtures pin both the
oracle restatement (CPU tests) and the CUDA path (GPU tests) to the reference itself; the
reference ships no golden vectors of its own.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eagcn_b200.data import make_batch  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def randomise(module, seed, scale_bn=True):
    """Reference init (layers.py:32-36,77-79,290-291) + utils.weights_init (utils.py:702-708) under a
    seed, then non-trivial BN affine/running stats so eval mode and d(gamma,beta) are exercised.
    AFM_BatchNorm.weight/.bias and Ave_multi_view are created uninitialised by the reference
    (layers.py:402-404) -- give them finite values."""
    torch.manual_seed(seed)
    for m in module.modules():
        if hasattr(m, "reset_parameters") and m is not module:
            try:
                m.reset_parameters()
            except Exception:
                pass
    _, _, U = ref_loader.load()
    module.apply(U.weights_init)
    g = torch.Generator().manual_seed(seed + 1)
    for name, p in module.named_parameters():
        if not torch.isfinite(p).all() or name.endswith("batch_norm.weight") or name.endswith("batch_norm.bias"):
            p.data = torch.rand(p.shape, generator=g) * 0.5 + 0.5
        if name.endswith("att.weight"):
            p.data = torch.randn(p.shape, generator=g)             # spread the attention logits
        if name.endswith("self_r"):
            p.data = torch.randn(p.shape, generator=g) * 0.5
        if name.endswith("graph_conv.weight") or name.endswith("graph_conv.bias"):
            p.data = torch.randn(p.shape, generator=g) * 0.3
    if scale_bn:
        for name, b in module.named_buffers():
            if name.endswith("running_mean"):
                b.data = torch.randn(b.shape, generator=g) * 0.1
            if name.endswith("running_var"):
                b.data = torch.rand(b.shape, generator=g) * 0.5 + 0.75


def save(case, batch, sd, extra):
    d = {"in.adj": batch.adj.astype(np.uint8), "in.afm": batch.afm, "in.codes": batch.codes,
         "in.sizes": batch.sizes, "in.channels": np.array(batch.channels, dtype=np.int64)}
    for k, v in sd.items():
        d["sd." + k] = v.detach().cpu().numpy()
    for k, v in extra.items():
        d[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(OUT, case + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def layer_case(case, B, seed, kb, fouts, fin=24, training=True, structure="Concate", last=False,
               dataset="freesolv", fixed_n=None):
    L, _, _ = ref_loader.load()
    batch = make_batch(B, dataset=dataset, seed=seed, kb=kb, n_afeat=fin, fixed_n=fixed_n)
    layer = L.GraphConv_Layer(fin, kb, *fouts, dropout=0.0, structure=structure, last=last)
    randomise(layer, seed)
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    layer.train(training)
    ins = [t(a) for a in batch.dense()]
    ins[1].requires_grad_(True)
    x, A = layer(*ins)
    g = torch.Generator().manual_seed(seed + 7)
    Rx = torch.randn(x.shape, generator=g)
    RA = torch.randn(A.shape, generator=g)
    loss = (x * Rx).sum() + (A * RA).sum()
    loss.backward()
    extra = {"out.x": x, "out.A": A, "cot.x": Rx, "cot.A": RA, "grad.afm": ins[1].grad,
             "meta.training": int(training), "meta.last": int(last),
             "meta.structure": np.array(structure), "meta.fouts": np.array(fouts)}
    for name, p in layer.named_parameters():
        if p.grad is not None:
            extra["grad." + name] = p.grad
    for k, v in layer.state_dict().items():   # post-step buffers (running stats after the train forward)
        if "running" in k or "num_batches" in k:
            extra["post." + k] = v
    save(case, batch, sd0, extra)


def model_case(case, B, seed, dataset, kb, sgc1, sgc2, den, nclass, training=True, molfp="sum", structure="Concate"):
    _, M, _ = ref_loader.load()
    batch = make_batch(B, dataset=dataset, seed=seed, kb=kb)
    model = M.EAGCN(kb, 24, *([sgc1] * 5), *([sgc2] * 5), den[0], den[1], nclass, dropout=0.0,
                    structure=structure, molfp_mode=molfp)
    randomise(model, seed)
    for name, p in model.named_parameters():     # head: keep activations O(1)
        if name.startswith("den"):
            p.data = p.data * 3.0
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model.train(training)
    ins = [t(a) for a in batch.dense()]
    out, atom_rep, graph_rep = model(*ins, t(batch.sizes))
    g = torch.Generator().manual_seed(seed + 7)
    R = torch.randn(out.shape, generator=g)
    (out * R).sum().backward()
    extra = {"out.y": out, "out.atom_rep": atom_rep, "out.graph_rep": graph_rep, "cot.y": R,
             "meta.training": int(training), "meta.molfp": np.array(molfp), "meta.structure": np.array(structure),
             "meta.dims": np.array([kb, sgc1, sgc2, den[0], den[1], nclass])}
    for name, p in model.named_parameters():
        if p.grad is not None:
            extra["grad." + name] = p.grad
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            extra["post." + k] = v
    save(case, batch, sd0, extra)


def stack_case(case, B, seed, dataset, kb, widths, den, nclass, training=True):
    """BASELINE.json's "2-layer" / "3-layer" configurations: the reference hard-codes four layers (models.py:50-61), so
    the golden is the reference's own GraphConv_Layer stacked ``len(widths)`` times, followed by the head of
    models.py:108-120 built from the reference's Dense and torch's BatchNorm1d under the reference's attribute names."""
    L, _, _ = ref_loader.load()
    import torch.nn as nn
    import torch.nn.functional as F
    batch = make_batch(B, dataset=dataset, seed=seed, kb=kb)

    class Stack(nn.Module):
        def __init__(self):
            super().__init__()
            fin = 24
            for l, w in enumerate(widths):
                setattr(self, f"layer{l + 1}", L.GraphConv_Layer(fin, kb, *w, dropout=0.0, structure="Concate"))
                fin = sum(w)
            self.den1, self.den2, self.den3 = L.Dense(fin, den[0]), L.Dense(den[0], den[1]), L.Dense(den[1], nclass)
            self.Graph_BN, self.bn_den1, self.bn_den2 = nn.BatchNorm1d(fin), nn.BatchNorm1d(den[0]), nn.BatchNorm1d(den[1])

        def forward(self, adjs, afms, *rels):
            x2 = afms
            for l in range(len(widths)):                                   # models.py:97-100
                x2, _ = getattr(self, f"layer{l + 1}")(adjs, x2, *rels)
            x = self.Graph_BN(torch.sum(x2, 1))                            # models.py:108,112
            x = F.relu(self.bn_den1(self.den1(x)))                         # models.py:114-115 (dropout 0)
            x = self.den2(x)
            g = x
            x = self.den3(F.relu(self.bn_den2(x)))                         # models.py:119-120
            return x, x2, g

    model = Stack()
    randomise(model, seed)
    for name, p in model.named_parameters():     # head: keep activations O(1)
        if name.startswith("den"):
            p.data = p.data * 3.0
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model.train(training)
    ins = [t(a) for a in batch.dense()]
    out, atom_rep, graph_rep = model(*ins)
    g = torch.Generator().manual_seed(seed + 7)
    R = torch.randn(out.shape, generator=g)
    (out * R).sum().backward()
    extra = {"out.y": out, "out.atom_rep": atom_rep, "out.graph_rep": graph_rep, "cot.y": R,
             "meta.training": int(training), "meta.n_layers": len(widths)}
    for name, p in model.named_parameters():
        if p.grad is not None:
            extra["grad." + name] = p.grad
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            extra["post." + k] = v
    save(case, batch, sd0, extra)


def main():
    assert ref_loader.available(), "needs /root/reference (build container)"
    torch.set_num_threads(1)
    only = set(sys.argv[1:])                      # optional: regenerate just the named cases
    global layer_case, model_case
    _lc, _mc = layer_case, model_case
    layer_case = lambda case, **kw: _lc(case, **kw) if (not only or case in only) else None   # noqa: E731
    model_case = lambda case, **kw: _mc(case, **kw) if (not only or case in only) else None   # noqa: E731
    layer_case("layer_train", B=5, seed=11, kb=7, fouts=(6, 5, 4, 3, 2))
    layer_case("layer_eval", B=4, seed=12, kb=5, fouts=(8, 8, 8, 8, 8), training=False)
    layer_case("layer_last", B=3, seed=13, kb=4, fouts=(4, 4, 4, 4, 4), last=True)
    layer_case("layer_wsum", B=3, seed=14, kb=6, fouts=(5, 5, 5, 5, 5), structure="Weighted_sum")
    layer_case("layer_wide", B=6, seed=15, kb=17, fouts=(40, 40, 40, 40, 40), fin=24, dataset="tox21")
    layer_case("layer_fixed", B=2, seed=16, kb=3, fouts=(3, 2, 2, 2, 1), fin=9, fixed_n=33)
    model_case("model_train", B=6, seed=21, dataset="freesolv", kb=5, sgc1=4, sgc2=6, den=(8, 4), nclass=3)
    model_case("model_eval", B=4, seed=22, dataset="freesolv", kb=5, sgc1=4, sgc2=6, den=(8, 4), nclass=1,
               training=False, molfp="ave")
    model_case("model_wsum", B=5, seed=23, dataset="freesolv", kb=5, sgc1=1, sgc2=2, den=(8, 4), nclass=2,
               structure="Weighted_sum")
    model_case("model_pool", B=5, seed=24, dataset="freesolv", kb=5, sgc1=3, sgc2=4, den=(8, 4), nclass=2, molfp="pool")
    if not only or "stack_lipo3" in only:      # config 3: Lipophilicity, 3 layers + regression head (widths / 10)
        stack_case("stack_lipo3", B=6, seed=31, dataset="lipo", kb=18, widths=[(6,) * 5, (10,) * 5, (20,) * 5],
                   den=(12, 6), nclass=1)
    if not only or "stack_hiv2" in only:       # config 4: HIV, 2 layers, widths that are not multiples of 4 (250 -> 25)
        stack_case("stack_hiv2", B=6, seed=32, dataset="hiv", kb=30, widths=[(10,) * 5, (25,) * 5], den=(16, 8), nclass=1)


if __name__ == "__main__":
    main()
