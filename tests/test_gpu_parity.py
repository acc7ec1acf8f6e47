"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libeagcn_sm100.so) via eagcn_b200; the checker is the oracle / the golden vectors produced by the
unmodified reference.  Tolerance: fp32, max|d|/max|ref| <= 1e-5 (north_star); indexing bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import eagcn_oracle as O
from tests.util import Golden, golden_cases, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _to(dev, xs):
    return [x.to(dev) for x in xs]


# ------------------------------------------------------------------ indexing: bit-exact
@pytest.mark.parametrize("case", golden_cases("layer_"))
def test_pack_roundtrip_bit_exact(case):
    from eagcn_b200.plan import GraphPlan
    dev = _cuda()
    g = Golden(case)
    adj, afm, *rels = _to(dev, g.dense())
    plan = GraphPlan.build(adj, rels).check()
    assert plan.n_rows == int((adj.sum(2) > 0).sum())
    assert plan.n_edges == int(adj.sum())
    for v in range(5):
        rel_back, adj_back = plan.unpack_view(v)
        assert torch.equal(rel_back, rels[v])                  # one-hot planes round-trip bit-exactly
        assert torch.equal(adj_back, adj)
    # same plan from the packed uint8 boundary
    codes = torch.from_numpy(g.batch.codes).to(dev)
    plan2 = GraphPlan.from_codes(codes, g.batch.channels).check()
    for name in ("pos_row", "mol_ptr"):
        assert torch.equal(getattr(plan, name), getattr(plan2, name))
    T, E = plan.n_rows, plan.n_edges
    assert (plan2.n_rows, plan2.n_edges) == (T, E)
    assert torch.equal(plan.row_ptr[:T + 1], plan2.row_ptr[:T + 1])
    for name in ("col", "colpos", "rev"):
        assert torch.equal(getattr(plan, name)[:E], getattr(plan2, name)[:E])
    assert torch.equal(plan.code[:, :E], plan2.code[:, :E])
    assert torch.equal(plan.rcode[:, :E], plan2.rcode[:, :E])
    # row mask == max_j adj (layers.py:295)
    assert torch.equal(plan.row_mask(), adj.max(2).values)
    # zero-copy boundary: one-hot planes left in pinned host memory, gathered over PCIe by the packer
    plan3 = GraphPlan.build(adj, [r.cpu().pin_memory() for r in rels]).check()
    assert (plan3.n_rows, plan3.n_edges) == (T, E)
    for name in ("col", "colpos", "rev"):
        assert torch.equal(getattr(plan, name)[:E], getattr(plan3, name)[:E])
    assert torch.equal(plan.code[:, :E], plan3.code[:, :E])
    assert torch.equal(plan.rcode[:, :E], plan3.rcode[:, :E])


@pytest.mark.parametrize("B,fixed_n,pad_to,kind", [(6, 40, 300, "sparse"), (3, 290, 290, "sparse"), (2, 270, 300, "dense")])
def test_pack_wide_padding_bit_exact(B, fixed_n, pad_to, kind):
    """Padded widths beyond 256 columns (the packer fetches adjacency rows 256 columns at a time) and rows with more
    than 256 neighbours: dense planes -> plan -> dense planes is the identity, and equals the uint8-code boundary."""
    from eagcn_b200.plan import GraphPlan
    from eagcn_b200.data import make_batch, DATASETS
    dev = _cuda()
    if kind == "dense":
        batch = _dense_graph_batch(B, fixed_n, DATASETS["tox21"]["kb"], 8, seed=5)
        import numpy as _np
        pad = pad_to - fixed_n
        batch.adj = _np.pad(batch.adj, ((0, 0), (0, pad), (0, pad)))
        batch.afm = _np.pad(batch.afm, ((0, 0), (0, pad), (0, 0)))
        batch.codes = _np.pad(batch.codes, ((0, 0), (0, 0), (0, pad), (0, pad)), constant_values=255)
    else:
        batch = make_batch(B, dataset="tox21", seed=11, fixed_n=fixed_n, pad_to=pad_to, n_afeat=8)
    adj, afm, *rels = _to(dev, [torch.from_numpy(a) for a in batch.dense()])
    plan = GraphPlan.build(adj, rels).check()
    assert plan.n_edges == int(adj.sum())
    for v in range(5):
        rel_back, adj_back = plan.unpack_view(v)
        assert torch.equal(rel_back, rels[v])
        assert torch.equal(adj_back, adj)
    plan2 = GraphPlan.from_codes(torch.from_numpy(batch.codes).to(dev), batch.channels).check()
    E = plan.n_edges
    assert plan2.n_edges == E
    for name in ("col", "colpos", "rev"):
        assert torch.equal(getattr(plan, name)[:E], getattr(plan2, name)[:E])
    assert torch.equal(plan.code[:, :E], plan2.code[:, :E])
    assert torch.equal(plan.rcode[:, :E], plan2.rcode[:, :E])


def test_pack_rejects_malformed():
    from eagcn_b200.plan import GraphPlan
    dev = _cuda()
    g = Golden("layer_train")
    adj, afm, *rels = _to(dev, g.dense())
    b, i, j = [int(t[0]) for t in torch.nonzero(adj, as_tuple=True)]
    bad = rels[1].clone(); bad[b, :, i, j] = 0.5
    with pytest.raises(ValueError, match="one-hot"):
        GraphPlan.build(adj, [rels[0], bad] + rels[2:]).check()
    asym = adj.clone(); asym[b, j, i] = 0.0
    with pytest.raises(ValueError, match="symmetric"):
        GraphPlan.build(asym, rels).check()
    nonbin = adj.clone(); nonbin[b, i, j] = 2.0
    with pytest.raises(ValueError, match="0.0/1.0"):
        GraphPlan.build(nonbin, rels).check()
    with pytest.raises(ValueError, match="capacity"):
        GraphPlan.build(adj, rels, t_cap=128, e_cap=3).check()


def test_cpu_tensors_fail_loudly():
    from eagcn_b200 import layers as EL
    from eagcn_b200._lib import EagcnError
    _cuda()
    g = Golden("layer_train")
    layer = EL.GraphConv_Layer(24, 7, 6, 5, 4, 3, 2, dropout=0.0, structure="Concate")
    with pytest.raises(EagcnError):
        layer(*g.dense())                                     # CPU tensors: no fallback


# ------------------------------------------------------------------ layer vs golden (reference outputs)
def _our_layer(g, dev):
    from eagcn_b200 import layers as EL
    fo = [int(x) for x in g.meta["fouts"]]
    fin = g.batch.afm.shape[2]
    layer = EL.GraphConv_Layer(fin, g.batch.channels[0], *fo, dropout=0.0, structure=str(g.meta["structure"]),
                               last=bool(g.meta["last"])).to(dev)
    missing = layer.load_state_dict(g.sd, strict=True)
    layer.train(bool(g.meta["training"]))
    return layer


@pytest.mark.parametrize("case", golden_cases("layer_"))
def test_layer_forward_backward_vs_golden(case):
    dev = _cuda()
    g = Golden(case)
    layer = _our_layer(g, dev)
    ins = _to(dev, g.dense())
    ins[1].requires_grad_(True)
    x, A = layer(*ins)
    assert rel_err(x.cpu(), g.out["x"]) <= TOL
    assert rel_err(A.cpu(), g.out["A"]) <= TOL
    m = ins[0].max(2).values
    if str(g.meta["structure"]) == "Concate":
        assert float((x.detach() * (1 - m).unsqueeze(2)).abs().max()) == 0.0  # layers.py:313: exact zeros
    else:                                                                     # layers.py:314-316: padded rows NOT masked
        assert float((x.detach() * (1 - m).unsqueeze(2)).abs().max()) > 0.0
    loss = (x * g.cot["x"].to(dev)).sum() + (A * g.cot["A"].to(dev)).sum()
    loss.backward()
    assert rel_err(ins[1].grad.cpu(), g.grad["afm"]) <= 2 * TOL
    scale = max(float(v.abs().max()) for v in g.grad.values())
    named = dict(layer.named_parameters())
    for k, ref in g.grad.items():
        if k == "afm":
            continue
        got = named[k].grad
        assert got is not None, k
        denom = max(float(ref.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and bool(g.meta["training"]):
            denom = scale
        assert float((got.cpu() - ref).abs().max()) / denom <= 5 * TOL, k
    if bool(g.meta["training"]):
        sd = layer.state_dict()
        for k, ref in g.post.items():
            if "num_batches" in k:
                assert int(sd[k]) == int(ref)
            else:
                assert rel_err(sd[k].cpu(), ref) <= TOL, k


def test_weighted_sum_eval_and_dropout():
    """'Weighted_sum' beyond the golden case (training, p = 0): eval mode against the oracle, and the training-mode
    dropout on padded rows keeps the reference's statistics (mean over many draws == the p = 0 value)."""
    from eagcn_b200 import layers as EL
    dev = _cuda()
    g = Golden("layer_wsum")
    layer = _our_layer(g, dev)
    layer.eval()
    ins = _to(dev, g.dense())
    ins[1].requires_grad_(True)
    x, _ = layer(*ins)
    sd = O.clone_sd({("layer1." + k): v for k, v in layer.state_dict().items()}, requires_grad=True)
    dense = g.dense()
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    afm_ref = dense[1].clone().requires_grad_(True)
    ref = O.layer_forward(sd, "layer1.", dense[0], afm_ref, codes, False, structure="Weighted_sum")
    assert rel_err(x.cpu(), ref["x"]) <= TOL
    R = torch.randn(ref["x"].shape, generator=torch.Generator().manual_seed(2))
    (ref["x"] * R).sum().backward()
    (x * R.to(dev)).sum().backward()
    assert rel_err(ins[1].grad.cpu(), afm_ref.grad) <= 2 * TOL
    named = dict(layer.named_parameters())
    scale = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    for k, prm in named.items():
        ref_g = sd["layer1." + k].grad
        if ref_g is None:
            continue
        assert prm.grad is not None, k
        assert float((prm.grad.cpu() - ref_g).abs().max()) / max(float(ref_g.abs().max()), 1e-3 * scale) <= 5 * TOL, k
    # dropout on the un-masked padded rows: unbiased
    layer2 = EL.GraphConv_Layer(layer.node_feature_in, layer.block1.bond_feature_num,
                                *[layer.total_output] * 5, dropout=0.5, structure="Weighted_sum").to(dev)
    layer2.load_state_dict(layer.state_dict())
    layer2.train()
    torch.manual_seed(0)
    with torch.no_grad():
        m = ins[0].max(2).values
        x0, _ = _train_p0(layer2, ins)
        acc = torch.zeros_like(x0)
        n = 600
        for _ in range(n):
            xi, _ = layer2(*[t.detach() for t in ins])
            acc += xi
        pad_mean = (acc / n) * (1 - m).unsqueeze(2)
        pad_ref = x0 * (1 - m).unsqueeze(2)
        assert float((pad_mean - pad_ref).abs().max()) <= 0.3 * float(pad_ref.abs().max()) + 1e-6


def _train_p0(layer, ins):
    p = [b.dropout for b in layer.blocks]
    for b in layer.blocks:
        b.dropout = 0.0
    try:
        return layer(*[t.detach() for t in ins])
    finally:
        for b, q in zip(layer.blocks, p):
            b.dropout = q


# ------------------------------------------------------------------ model vs golden
@pytest.mark.parametrize("case", golden_cases("model_"))
def test_model_vs_golden(case):
    from eagcn_b200 import models as EM
    dev = _cuda()
    g = Golden(case)
    kb, s1, s2, d1, d2, nc = [int(x) for x in g.meta["dims"]]
    structure = str(g.meta["structure"]) if "structure" in g.meta else "Concate"
    model = EM.EAGCN(kb, 24, *([s1] * 5), *([s2] * 5), d1, d2, nc, dropout=0.0, structure=structure,
                     molfp_mode=str(g.meta["molfp"])).to(dev)
    model.load_state_dict(g.sd, strict=True)                  # reference checkpoint, incl. ave / pool1 / pool3 keys
    training = bool(g.meta["training"])
    model.train(training)
    ins = _to(dev, g.dense())
    y, atom_rep, graph_rep = model(*ins, torch.from_numpy(g.batch.sizes).to(dev))
    assert rel_err(atom_rep.materialize(), g.out["atom_rep"]) <= TOL
    assert rel_err(y.cpu(), g.out["y"]) <= 5 * TOL
    assert rel_err(graph_rep.cpu(), g.out["graph_rep"]) <= 5 * TOL
    (y * g.cot["y"].to(dev)).sum().backward()
    scale = max(float(v.abs().max()) for v in g.grad.values())
    named = dict(model.named_parameters())
    for k, ref in g.grad.items():
        got = named[k].grad
        assert got is not None, k
        denom = max(float(ref.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and training:
            denom = scale
        assert float((got.cpu() - ref).abs().max()) / denom <= 2e-4, k


# ------------------------------------------------------------------ layer vs oracle at realistic widths
def _oracle_case(dev, B, dataset, fin, fo, seed, training, p=0.0, kb=None):
    from eagcn_b200 import layers as EL, functional as EF
    from eagcn_b200.data import make_batch, DATASETS
    kb = kb or DATASETS[dataset]["kb"]
    batch = make_batch(B, dataset=dataset, seed=seed, kb=kb, n_afeat=fin)
    torch.manual_seed(seed)
    layer = EL.GraphConv_Layer(fin, kb, *fo, dropout=p, structure="Concate").to(dev)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, prm in layer.named_parameters():
            if n.endswith("att.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g))
            elif n.endswith("graph_conv.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g) * (1.0 / fin ** 0.5))
            elif n.endswith("bn.weight"):
                prm.copy_(torch.rand(prm.shape, generator=g) + 0.5)
            elif n.endswith("bn.bias"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.1)
    layer.train(training)
    return batch, layer


@pytest.mark.parametrize("B,dataset,fin,fo,training", [
    (64, "tox21", 24, (80,) * 5, True),
    (48, "tox21", 400, (140,) * 5, True),
    (32, "lipo", 300, (100,) * 5, False),
    (16, "hiv", 24, (100, 52, 36, 20, 12), True),
])
def test_layer_vs_oracle(B, dataset, fin, fo, training):
    dev = _cuda()
    batch, layer = _oracle_case(dev, B, dataset, fin, fo, seed=B, training=training)
    dense = [torch.from_numpy(a) for a in batch.dense()]
    sd = O.clone_sd({("layer1." + k): v for k, v in layer.state_dict().items()}, requires_grad=True)
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    ins = _to(dev, dense)
    ins[1].requires_grad_(True)
    x, A = layer(*ins)
    # ReLU decisions of the implementation under test (p = 0: x > 0 <=> pre-activation > 0), per view
    xg = x.detach().cpu()
    masks, off = [], 0
    for w in fo:
        masks.append((xg[:, :, off:off + w] > 0).float()); off += w
    afm_ref = dense[1].clone().requires_grad_(True)
    ref = O.layer_forward(sd, "layer1.", dense[0], afm_ref, codes, training, relu_masks=masks)
    assert rel_err(x.cpu(), ref["x"]) <= TOL
    assert rel_err(A.cpu(), ref["A_weight"]) <= TOL
    # ... which may differ from the oracle's own (Z > 0) only at the kink: |Z| below the forward tolerance
    m = dense[0].max(2).values.unsqueeze(2)
    for v, Zv in enumerate(ref["Z"]):
        disagree = ((Zv.detach() > 0).float() != masks[v]) & (m > 0)
        if bool(disagree.any()):
            assert float(Zv.detach().abs()[disagree].max()) <= 1e-4 * float(Zv.detach().abs().max())
            assert int(disagree.sum()) <= 1e-4 * disagree.numel() + 8
    gen = torch.Generator().manual_seed(1)
    R = torch.randn(ref["x"].shape, generator=gen)
    (ref["x"] * R).sum().backward()
    (x * R.to(dev)).sum().backward()
    assert rel_err(ins[1].grad.cpu(), afm_ref.grad) <= 2 * TOL
    named = dict(layer.named_parameters())
    scale = max(float(sd["layer1." + k].grad.abs().max()) for k in named if sd["layer1." + k].grad is not None)
    for k, prm in named.items():
        ref_g = sd["layer1." + k].grad
        if ref_g is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, k
            continue
        denom = max(float(ref_g.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and training:
            denom = scale
        assert float((prm.grad.cpu() - ref_g).abs().max()) / denom <= 5 * TOL, k


def test_dropout_mask_consistency():
    """Training with dropout: the CUDA path's keep mask (its own philox stream -- torch's global RNG of
    layers.py:94 cannot be matched) is exported, fed to the oracle, and forward+backward must agree."""
    from eagcn_b200 import functional as EF
    dev = _cuda()
    p = 0.3
    batch, layer = _oracle_case(dev, 40, "tox21", 24, (16, 12, 8, 8, 4), seed=9, training=True, p=p)
    EF.manual_seed(1234, dev)
    rng_before = EF.RngState.get(dev).state.clone()
    dense = [torch.from_numpy(a) for a in batch.dense()]
    ins = _to(dev, dense)
    ins[1].requires_grad_(True)
    x, _ = layer(*ins)
    from eagcn_b200.layers import GraphConv_Layer
    plan = GraphConv_Layer._plan_cache[2]
    cfg = EF.LayerConfig(fin=24, fo=(16, 12, 8, 8, 4), training=True, p_drop=p, rng_stream=0)
    keep_packed = EF.dropout_keep_mask(plan, cfg, 48, rng_before).float()
    frac = float(keep_packed[:plan.n_rows].mean())
    assert abs(frac - (1 - p)) < 0.02
    keep_dense = plan.scatter(keep_packed).cpu()
    off, keeps = 0, []
    for fo in cfg.fo:
        keeps.append(keep_dense[:, :, off:off + fo]); off += fo
    sd = O.clone_sd({("layer1." + k): v for k, v in layer.state_dict().items()}, requires_grad=True)
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    afm_ref = dense[1].clone().requires_grad_(True)
    # running stats were already updated by the CUDA forward; the oracle only needs batch stats
    ref = O.layer_forward(sd, "layer1.", dense[0], afm_ref, codes, True, p=p, keeps=keeps)
    assert rel_err(x.cpu(), ref["x"]) <= TOL
    gen = torch.Generator().manual_seed(2)
    R = torch.randn(ref["x"].shape, generator=gen)
    (ref["x"] * R).sum().backward()
    (x * R.to(dev)).sum().backward()
    assert rel_err(ins[1].grad.cpu(), afm_ref.grad) <= 2 * TOL
    g_ref = sd["layer1.block1.graph_conv.weight"].grad
    assert rel_err(layer.block1.graph_conv.weight.grad.cpu(), g_ref) <= 5 * TOL
    # a second call draws a different mask
    x2, _ = layer(*[t.detach() for t in ins])
    assert not torch.equal((x2 == 0), (x == 0))


# ------------------------------------------------------------------ size-independent properties at full size
def test_full_size_properties():
    """BASELINE config 2 shape (Tox21, B=256, 24->400->700): properties that need no oracle run."""
    from eagcn_b200 import models as EM
    from eagcn_b200.data import make_batch
    dev = _cuda()
    torch.manual_seed(0)
    model = EM.EAGCNStack(30, 24, [(80,) * 5, (140,) * 5], 256, 64, 12, dropout=0.0).to(dev)
    model.eval()
    batch = make_batch(256, "tox21", seed=0)
    ins = _to(dev, [torch.from_numpy(a) for a in batch.dense()])
    size = torch.from_numpy(batch.sizes).to(dev)
    with torch.no_grad():
        y, atom, _ = model(*ins, size)
        # (1) determinism: bitwise identical on a re-run
        y2, _, _ = model(*ins, size)
        assert torch.equal(y, y2)
        # (2) molecules are independent in eval mode: permuting the batch permutes the outputs
        perm = torch.randperm(256, generator=torch.Generator().manual_seed(0)).to(dev)
        yp, _, _ = model(*[t[perm] for t in ins], size[perm])
        assert rel_err(yp.cpu(), y[perm].cpu()) <= TOL
        # (3) padded rows of the atom representation are exact zeros
        dense = atom.materialize()
        m = ins[0].max(2).values.cpu()
        assert float((dense * (1 - m).unsqueeze(2)).abs().max()) == 0.0
        # (4) padding invariance: a wider zero padding only moves the 1e-9 terms
        wide = make_batch(256, "tox21", seed=0, pad_to=batch.N + 17)
        insw = _to(dev, [torch.from_numpy(a) for a in wide.dense()])
        yw, _, _ = model(*insw, size)
        assert rel_err(yw.cpu(), y.cpu()) <= TOL
    # (5) packed uint8 boundary gives bit-identical results to the dense one-hot boundary
    from eagcn_b200.plan import GraphPlan
    with torch.no_grad():
        plan = GraphPlan.from_codes(torch.from_numpy(batch.codes).to(dev), batch.channels).check()
        yc, _, _ = model(plan, ins[1], size=size)
        assert torch.equal(yc, y)


# ------------------------------------------------------------------ K != 5 views (BASELINE config 5 sweep)
@pytest.mark.parametrize("V,fixed_n,B", [(1, 32, 24), (10, 64, 12), (5, 128, 6)])
def test_view_count_sweep_vs_oracle(V, fixed_n, B):
    """K in {1, 5, 10} views, fixed-N batches: the functional core against K independent oracle blocks
    (the reference hard-codes 5 blocks, layers.py:269-273; the oracle for K != 5 is K GraphConv_blocks + cat)."""
    from eagcn_b200 import functional as EF
    from eagcn_b200.data import make_batch
    from eagcn_b200.plan import GraphPlan
    dev = _cuda()
    fin, fo = 24, tuple([16, 12, 8, 8, 4, 12, 8, 4, 4, 4][:V])
    batch = make_batch(B, "tox21", seed=V, kb=9, n_views=V, fixed_n=fixed_n)
    gen = torch.Generator().manual_seed(V)
    sd = {}
    for v in range(V):
        pre = f"layer1.block{v + 1}."
        C = batch.channels[v]
        sd[pre + "att.weight"] = torch.randn(1, C, 1, 1, generator=gen)
        sd[pre + "self_r"] = torch.randn(1, generator=gen) * 0.3
        sd[pre + "graph_conv.weight"] = torch.randn(fin, fo[v], generator=gen) * 0.2
        sd[pre + "graph_conv.bias"] = torch.randn(fo[v], generator=gen) * 0.1
        sd[pre + "batch_norm.bn.weight"] = torch.rand(fo[v], generator=gen) + 0.5
        sd[pre + "batch_norm.bn.bias"] = torch.randn(fo[v], generator=gen) * 0.1
        sd[pre + "batch_norm.bn.running_mean"] = torch.zeros(fo[v])
        sd[pre + "batch_norm.bn.running_var"] = torch.ones(fo[v])
    dense = [torch.from_numpy(a) for a in batch.dense()]
    ref_sd = O.clone_sd(sd, requires_grad=True)
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    afm_ref = dense[1].clone().requires_grad_(True)
    ref = O.layer_forward(ref_sd, "layer1.", dense[0], afm_ref, codes, True, n_views=V)

    dsd = {k: v.to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    plan = GraphPlan.build(dense[0].to(dev), [r.to(dev) for r in dense[2:]]).check()
    params, buffers = [], []
    for v in range(V):
        pre = f"layer1.block{v + 1}."
        params += [dsd[pre + k] for k in ("att.weight", "self_r", "graph_conv.weight", "graph_conv.bias",
                                         "batch_norm.bn.weight", "batch_norm.bn.bias")]
        buffers += [dsd[pre + "batch_norm.bn.running_mean"], dsd[pre + "batch_norm.bn.running_var"], None]
    afm = dense[1].to(dev).requires_grad_(True)
    cfg = EF.LayerConfig(fin=fin, fo=fo, training=True)
    X = EF.graph_conv_layer(plan, cfg, EF.gather_rows(plan, afm), params, buffers)
    x = EF.scatter_rows(plan, X)
    assert rel_err(x.cpu(), ref["x"]) <= TOL
    gen2 = torch.Generator().manual_seed(5)
    R = torch.randn(ref["x"].shape, generator=gen2)
    (ref["x"] * R).sum().backward()
    (x * R.to(dev)).sum().backward()
    assert rel_err(afm.grad.cpu(), afm_ref.grad) <= 2 * TOL
    scale = max(float(t.grad.abs().max()) for t in ref_sd.values() if t.grad is not None)
    for k, t in ref_sd.items():
        if t.grad is None:
            continue
        denom = max(float(t.grad.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias"):
            denom = scale
        assert float((dsd[k].grad.cpu() - t.grad).abs().max()) / denom <= 5 * TOL, k


# ------------------------------------------------------------------ dense head
@pytest.mark.parametrize("B,F,D1,D2,NC,training,p", [
    (64, 100, 48, 24, 5, True, 0.0), (256, 700, 256, 64, 12, True, 0.0), (37, 70, 33, 17, 1, False, 0.0),
    (128, 300, 128, 64, 3, True, 0.3),
])
def test_head_vs_oracle(B, F, D1, D2, NC, training, p):
    """models.py:112-120 on the library's kernels (fused BatchNorm/ReLU/dropout launches between mm_tile products) vs the
    oracle: the dropout mask is exported from the library's stream, the ReLU decisions are taken from the
    implementation under test."""
    import torch.nn as nn
    from eagcn_b200 import functional as EF
    dev = _cuda()
    gen = torch.Generator().manual_seed(B + F)
    x0 = (torch.randn(B, F, generator=gen) * 2 + 0.5)
    sd = {"den1.weight": torch.randn(F, D1, generator=gen) / F ** 0.5, "den2.weight": torch.randn(D1, D2, generator=gen) / D1 ** 0.5,
          "den3.weight": torch.randn(D2, NC, generator=gen) / D2 ** 0.5}
    bns = []
    for name, n in (("Graph_BN", F), ("bn_den1", D1), ("bn_den2", D2)):
        bn = nn.BatchNorm1d(n)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(n, generator=gen) + 0.5); bn.bias.copy_(torch.randn(n, generator=gen) * 0.2)
            bn.running_mean.copy_(torch.randn(n, generator=gen) * 0.3); bn.running_var.copy_(torch.rand(n, generator=gen) + 0.5)
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[f"{name}.{k}"] = getattr(bn, k).detach().clone()
        bns.append(bn.to(dev).train(training))
    Ws = [sd[f"den{i}.weight"].to(dev).requires_grad_(True) for i in (1, 2, 3)]
    xg = x0.to(dev).requires_grad_(True)
    EF.manual_seed(77, dev)
    rng_before = EF.RngState.get(dev).state.clone()
    x = EF.bn_act(xg, bns[0], training)                                           # models.py:112
    a1 = EF.bn_act(EF.dense_mm(x, Ws[0]), bns[1], training, relu=True, p_drop=p)  # models.py:114-116
    a2 = EF.dense_mm(a1, Ws[1])                                                   # models.py:117 (graph_representation)
    a3 = EF.bn_act(a2, bns[2], training, relu=True)                               # models.py:119
    out = EF.dense_mm(a3, Ws[2])                                                  # models.py:120
    keep = None
    if training and p > 0:
        keep = EF.dropout_keep_mask_flat(rng_before, 1000, p, B * D1).view(B, D1).float().cpu()
        assert abs(float(keep.mean()) - (1 - p)) < 0.03
    one = torch.ones(B, D1)
    relu_masks = ((a1.detach().cpu() > 0).float() + (one - (keep if keep is not None else one)), (a3.detach().cpu() > 0).float())
    ref_sd = O.clone_sd(sd, requires_grad=True)
    x_ref = x0.clone().requires_grad_(True)
    y_ref, g_ref = O.head_forward(ref_sd, None, None, training, p=p, keep=keep, x0=x_ref, relu_masks=relu_masks)
    assert rel_err(out.cpu(), y_ref) <= 2 * TOL
    assert rel_err(a2.cpu(), g_ref) <= 2 * TOL
    gen2 = torch.Generator().manual_seed(3)
    R, R2 = torch.randn(y_ref.shape, generator=gen2), torch.randn(g_ref.shape, generator=gen2)
    ((y_ref * R).sum() + (g_ref * R2).sum()).backward()
    ((out * R.to(dev)).sum() + (a2 * R2.to(dev)).sum()).backward()
    scale = max(float(t.grad.abs().max()) for t in ref_sd.values() if t.grad is not None)
    checks = [(xg.grad, x_ref.grad, "x0")] + [(Ws[i].grad, ref_sd[f"den{i + 1}.weight"].grad, f"den{i + 1}") for i in range(3)]
    for bn, name in zip(bns, ("Graph_BN", "bn_den1", "bn_den2")):
        checks += [(bn.weight.grad, ref_sd[name + ".weight"].grad, name + ".w"), (bn.bias.grad, ref_sd[name + ".bias"].grad, name + ".b")]
    for got, ref, name in checks:
        denom = max(float(ref.abs().max()), 1e-3 * scale)
        if name == "Graph_BN.b" and training:
            denom = scale         # a constant shift ahead of den1 -> bn_den1 is removed by that BatchNorm: analytically 0, noise
        assert float((got.cpu() - ref).abs().max()) / denom <= 5 * TOL, name
    if training:
        for bn, name in zip(bns, ("Graph_BN", "bn_den1", "bn_den2")):
            assert int(bn.num_batches_tracked) == 1
        exp_rm = 0.9 * sd["Graph_BN.running_mean"] + 0.1 * x0.mean(0)
        exp_rv = 0.9 * sd["Graph_BN.running_var"] + 0.1 * x0.var(0, unbiased=True)
        assert rel_err(bns[0].running_mean.cpu(), exp_rm) <= TOL
        assert rel_err(bns[0].running_var.cpu(), exp_rv) <= TOL


# ------------------------------------------------------------------ BASELINE.json configurations as parity cases
@pytest.mark.parametrize("name,dataset,B,n_layers,training", [
    ("config1 freesolv 2-layer fp32 forward", "freesolv", 16, 2, False),
    ("config3 lipo 3-layer + regression head", "lipo", 24, 3, True),
    ("config4 hiv widths 2-layer", "hiv", 12, 2, True),
    ("full 4-layer tox21-width model", "tox21", 8, 4, True),
])
def test_baseline_configs_forward_vs_oracle(name, dataset, B, n_layers, training):
    """The per-dataset layer widths of train.py:61-114 stacked 2 / 3 / 4 deep (+ the reference head), forward
    parity of atom representations, graph representation and logits against the oracle."""
    from eagcn_b200 import models as EM
    from eagcn_b200.data import make_batch, DATASETS
    dev = _cuda()
    cfg = DATASETS[dataset]
    s1, s2 = cfg["sgc1"], cfg["sgc2"]
    widths = [(s1,) * 5, (s2,) * 5, (2 * s2,) * 5, (2 * s2,) * 5][:n_layers]
    torch.manual_seed(1)
    model = EM.EAGCNStack(cfg["kb"], 24, widths, cfg["den"][0], cfg["den"][1], cfg["nclass"], dropout=0.0).to(dev)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if n.endswith("graph_conv.weight") or n.startswith("den"):
                prm.copy_(torch.randn(prm.shape, generator=g) * (1.5 / prm.shape[0] ** 0.5))
            elif n.endswith("att.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g))
    model.train(training)
    batch = make_batch(B, dataset, seed=4)
    dense = [torch.from_numpy(a) for a in batch.dense()]
    sd = O.clone_sd(model.state_dict())
    y, atom, grep = model(*_to(dev, dense), torch.from_numpy(batch.sizes).to(dev))
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    h, _ = O.stack_forward(sd, dense[0], dense[1], codes, n_layers, training)
    y_ref, g_ref = O.head_forward(sd, h, torch.from_numpy(batch.sizes), training)
    # deeper stacks of train-mode BatchNorms amplify fp32 differences a little: 1e-5 per layer
    assert rel_err(atom.materialize(), h) <= n_layers * TOL, name
    assert rel_err(grep.detach().cpu(), g_ref) <= 2 * n_layers * TOL, name
    assert rel_err(y.detach().cpu(), y_ref) <= 2 * n_layers * TOL, name


# ------------------------------------------------------------------ aggregation engines agree
def _dense_graph_batch(B, n, kb, fin, seed):
    """Fully connected molecules (degree n-1 > 32, > 512 edges per row tile): the tile kernels' fallbacks."""
    from eagcn_b200.data import MolBatch, view_channels, NO_EDGE
    rng = np.random.default_rng(seed)
    chans = view_channels(kb, 5)
    adj = np.zeros((B, n, n), np.float32)
    codes = np.full((B, 5, n, n), NO_EDGE, np.uint8)
    iu = np.triu_indices(n, 1)
    for b in range(B):
        adj[b][iu] = 1.0
        adj[b] += adj[b].T
        for v in range(5):
            c = rng.integers(0, chans[v], size=len(iu[0])).astype(np.uint8)
            codes[b, v][iu] = c
            codes[b, v].T[iu] = c
    afm = rng.random((B, n, fin), dtype=np.float32)
    return MolBatch(adj=adj, afm=afm, codes=codes, sizes=np.full(B, n, np.int64), channels=chans)


@pytest.mark.parametrize("kind,B,fin,fo,training,p", [
    ("tox21", 64, 24, (80,) * 5, True, 0.3),
    ("tox21", 48, 400, (140,) * 5, True, 0.3),
    ("tox21", 48, 64, (140,) * 5, False, 0.0),
    ("hiv", 16, 24, (100, 52, 36, 20, 12), True, 0.0),
    ("lipo", 24, 40, (280,) * 5, True, 0.3),
    ("lipo", 12, 24, (512, 400, 260, 8, 4), True, 0.0),
    ("dense", 6, 24, (80,) * 5, True, 0.3),
    ("dense", 5, 32, (140, 140, 36, 20, 300), True, 0.0),
])
def test_agg_engines_agree(kind, B, fin, fo, training, p):
    """The shared-memory tile kernels (default) against the generic warp-per-row kernels: same dropout stream,
    same inputs -> same outputs and gradients up to summation order (the generic path is checked against the
    oracle above; both are checked against it implicitly through every other test in this file)."""
    from eagcn_b200 import functional as EF
    from eagcn_b200.data import DATASETS
    dev = _cuda()
    if kind == "dense":
        kb = DATASETS["tox21"]["kb"]
        batch, layer = _oracle_case(dev, 4, "tox21", fin, fo, seed=7, training=training, p=p)
        batch = _dense_graph_batch(B, 41, kb, fin, seed=3)
    else:
        batch, layer = _oracle_case(dev, B, kind, fin, fo, seed=B + 1, training=training, p=p)
    dense = [torch.from_numpy(a) for a in batch.dense()]
    gen = torch.Generator().manual_seed(5)
    R = None
    res = {}
    try:
        for eng in ("generic", "tile"):
            EF.set_agg_engine(eng)
            EF.manual_seed(1234)
            layer.zero_grad(set_to_none=True)
            ins = _to(dev, dense)
            ins[1].requires_grad_(True)
            x, A = layer(*ins)
            if R is None:
                R = torch.randn(x.shape, generator=gen).to(dev)
            (x * R).sum().backward()
            res[eng] = (x.detach().clone(), ins[1].grad.clone(),
                        {k: prm.grad.clone() for k, prm in layer.named_parameters() if prm.grad is not None})
    finally:
        EF.set_agg_engine("tile")
    xg, hg, pg = res["generic"]
    xt, ht, pt = res["tile"]
    # fully connected graphs make every row of a molecule nearly the same average -> BatchNorm variances far below
    # the squared means; the two engines sum the statistics in different orders, which that conditioning amplifies
    loose = 50.0 if kind == "dense" else 1.0
    assert rel_err(xt, xg) <= 2e-6 * loose
    assert rel_err(ht, hg) <= 1e-5 * loose
    assert pg.keys() == pt.keys()
    scale = max(float(g.abs().max()) for g in pg.values())
    for k in pg:
        denom = max(float(pg[k].abs().max()), 1e-3 * scale)
        assert float((pt[k] - pg[k]).abs().max()) / denom <= 2e-5 * loose, k


# ------------------------------------------------------------------ fused BatchNorm(+ReLU)(+dropout) of the head
@pytest.mark.parametrize("B,C,training,relu,p", [
    (256, 1400, True, False, 0.0),      # Graph_BN (models.py:112)
    (256, 256, True, True, 0.3),        # relu(bn_den1) + dropout (models.py:114-116)
    (256, 64, True, True, 0.0),         # relu(bn_den2) (models.py:119)
    (37, 45, True, True, 0.5),          # ragged sizes
    (64, 96, False, True, 0.3),         # eval: running statistics, dropout off
    (300, 40, True, True, 0.3),         # B > 256: multi-pass path
    (5, 8, True, False, 0.0),
    (512, 700, True, True, 0.3),        # Lipophilicity batch: 4 rows per thread, 8 channels per CTA
    (1000, 64, True, True, 0.3),        # B <= 1024: one float4 group per CTA
    (1100, 64, True, True, 0.3),        # beyond the register-resident range: 32-channel multi-pass kernels
])
@pytest.mark.parametrize("mode", [0, 1])
def test_bn_act_vs_torch(B, C, training, relu, p, mode):
    """eagcn_bn_act_forward/backward against stock torch fp32 ops (batch_norm -> relu -> dropout with the same keep
    mask), values, running statistics and all gradients.  mode 0: float4 kernels (where the layout allows), 1: the
    32-channel kernels."""
    from eagcn_b200 import functional as EF, _lib
    dev = _cuda()
    _lib.lib().eagcn_set_bn_act_mode(mode)
    try:
        _bn_act_case(EF, dev, B, C, training, relu, p)
    finally:
        _lib.lib().eagcn_set_bn_act_mode(0)


def _bn_act_case(EF, dev, B, C, training, relu, p):
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = (torch.randn(B, C, generator=g) * 2.0 + 0.7).to(dev)
    bn = torch.nn.BatchNorm1d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ref_bn = torch.nn.BatchNorm1d(C).to(dev)
    ref_bn.load_state_dict(bn.state_dict())
    bn.train(training); ref_bn.train(training)
    EF.manual_seed(99)
    rng_state = EF.RngState.get(dev).state.clone()
    xa = x.clone().requires_grad_(True)
    y = EF.bn_act(xa, bn, training, relu=relu, p_drop=p, rng_stream=1000)
    xb = x.clone().requires_grad_(True)
    z = ref_bn(xb)
    if relu:
        z = torch.relu(z)
    if training and p > 0:
        keep = EF.dropout_keep_mask_flat(rng_state, 1000, p, B * C).view(B, C).float()
        z = z * keep / (1.0 - p)
        frac = float(keep.mean())
        assert abs(frac - (1.0 - p)) < 0.08
    assert rel_err(y, z) <= TOL
    assert rel_err(bn.running_mean, ref_bn.running_mean) <= TOL
    assert rel_err(bn.running_var, ref_bn.running_var) <= TOL
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked)
    R = torch.randn(B, C, generator=g).to(dev)
    (y * R).sum().backward()
    (z * R).sum().backward()
    assert rel_err(xa.grad, xb.grad) <= 2 * TOL
    assert rel_err(bn.weight.grad, ref_bn.weight.grad) <= 2 * TOL
    assert rel_err(bn.bias.grad, ref_bn.bias.grad) <= 2 * TOL


@pytest.mark.parametrize("training", [False, True])
def test_weighted_sum_model_vs_oracle(training):
    """EAGCN(structure='Weighted_sum') (models.py:33-47, layers.py:314-316): four layers whose un-masked padded rows
    flow into the read-out sum, against the oracle, outputs and gradients."""
    from eagcn_b200 import models as EM
    from eagcn_b200.data import make_batch
    dev = _cuda()
    kb = 5
    batch = make_batch(6, dataset="freesolv", seed=31, kb=kb)
    torch.manual_seed(5)
    model = EM.EAGCN(kb, 24, 2, 1, 1, 1, 1, 3, 1, 1, 1, 2, 8, 4, 3, dropout=0.0, structure="Weighted_sum").to(dev)
    gen = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if n.endswith("bn.weight") or n.endswith("BN.weight") or "bn_den" in n and n.endswith("weight"):
                prm.copy_(torch.rand(prm.shape, generator=gen) + 0.5)
            elif n.endswith("ave.weight"):
                prm.copy_(torch.rand(prm.shape, generator=gen) + 0.2)
    model.train(training)
    dense = [torch.from_numpy(a) for a in batch.dense()]
    sizes = torch.from_numpy(batch.sizes)
    sd = O.clone_sd(model.state_dict(), requires_grad=True)
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    h_ref, _ = O.stack_forward(sd, dense[0], dense[1], codes, 4, training, structure="Weighted_sum",
                               last_flags=[False, False, False, True])
    out_ref, g_ref = O.head_forward(sd, h_ref, sizes, training)
    ins = _to(dev, dense)
    out, atom_rep, g_rep = model(*ins, size=sizes.to(dev))
    assert rel_err(atom_rep.materialize(), h_ref.detach()) <= TOL
    assert rel_err(out.cpu(), out_ref) <= 2 * TOL
    R = torch.randn(out_ref.shape, generator=gen)
    (out_ref * R).sum().backward()
    (out * R.to(dev)).sum().backward()
    named = dict(model.named_parameters())
    scale = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    checked = 0
    for k, prm in named.items():
        ref_g = sd[k].grad
        if ref_g is None or prm.grad is None:
            continue
        denom = max(float(ref_g.abs().max()), 1e-2 * scale)
        assert float((prm.grad.cpu() - ref_g).abs().max()) / denom <= 2e-4, k
        checked += 1
    assert checked >= 40


# ------------------------------------------------------------------ side-stream branches change nothing
def _model_step(dev, overlap, graph):
    """One training step (fwd + bwd) of a 2-layer model; returns outputs and all gradients."""
    from eagcn_b200 import functional as EF, models as EM
    from eagcn_b200.data import make_batch
    from eagcn_b200.plan import GraphPlan
    EF.Overlap.enabled = overlap
    torch.manual_seed(0)
    model = EM.EAGCNStack(17, 24, [(16,) * 5, (24,) * 5], 32, 16, 3, dropout=0.3).to(dev)
    model.train()
    batch = make_batch(24, "freesolv", seed=3)
    dense = [torch.from_numpy(a).to(dev) for a in batch.dense()]
    size = torch.from_numpy(batch.sizes).to(dev)
    T, E = int((batch.adj.sum(2) > 0).sum()), int(batch.adj.sum())
    params = [p for p in model.parameters()]

    def step():
        for p in params:
            p.grad = None
        model.prefetch_params()
        plan = GraphPlan.build(dense[0], dense[2:], t_cap=T, e_cap=E)
        out, _, grep = model(plan, dense[1], size=size)
        (out.sum() + grep.sum()).backward()
        return out

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            EF.manual_seed(7)
            out = step()
        if graph:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            EF.manual_seed(7)
            with torch.cuda.graph(g):
                out = step()
            EF.manual_seed(7)
            g.replay()
        torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return out.detach().clone(), grads


@pytest.mark.parametrize("graph", [False, True])
def test_overlap_branches_bit_identical(graph):
    """Overlap (parameter prep || packing, dW || dH + next layer, head dW || dX; deferred joins) only re-orders
    independent kernels: outputs and every gradient are bit-identical to the single-stream run, eagerly and as a
    captured CUDA graph (the side-stream work becomes parallel graph branches)."""
    from eagcn_b200 import functional as EF
    dev = _cuda()
    old = EF.Overlap.enabled
    try:
        o0, g0 = _model_step(dev, False, False)
        o1, g1 = _model_step(dev, True, graph)
    finally:
        EF.Overlap.enabled = old
    assert torch.equal(o0, o1)
    assert set(g0) == set(g1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k


def test_rng_prefork_matches_forks():
    from eagcn_b200 import functional as EF
    dev = _cuda()
    r = EF.RngState.get(dev)
    EF.manual_seed(11)
    a = [r.fork().clone() for _ in range(3)]
    end_a = r.state.clone()
    EF.manual_seed(11)
    r.prefork(3)
    b = [r.fork().clone() for _ in range(3)]
    assert not r._queue
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert torch.equal(end_a, r.state)


def test_fused_bn_forward_bit_identical():
    """eagcn_layer_forward_b as one launch (statistics reduction + finalize + apply) against the separate kernels:
    same reduction order, so outputs, saved statistics and running statistics are bit-identical."""
    from eagcn_b200 import functional as EF, models as EM, _lib
    from eagcn_b200.data import make_batch
    dev = _cuda()
    batch = make_batch(32, "tox21", seed=5)
    dense = [torch.from_numpy(a).to(dev) for a in batch.dense()]
    size = torch.from_numpy(batch.sizes).to(dev)
    res = []
    for mode in (1, 0):
        _lib.lib().eagcn_set_fuse_mode(mode)
        try:
            torch.manual_seed(0)
            model = EM.EAGCNStack(30, 24, [(16,) * 5, (28,) * 5], 32, 16, 3, dropout=0.3).to(dev)
            model.train()
            EF.manual_seed(5)
            out, atom, _ = model(*dense, size)
            out.sum().backward()
            torch.cuda.synchronize()
            res.append((out.detach().clone(), atom.materialize().clone(),
                        {k: v.detach().clone() for k, v in model.state_dict().items()},
                        {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}))
        finally:
            _lib.lib().eagcn_set_fuse_mode(0)
    (o0, a0, sd0, g0), (o1, a1, sd1, g1) = res
    assert torch.equal(o0, o1) and torch.equal(a0, a1)
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k


def test_padded_widths_hiv_config_vs_oracle():
    """GraphConv_Layer.pad_widths: HIV widths (5 x 250 in layer 2: row stride 1 250 floats) on the padded layout --
    tensor-core GEMM + float4 kernels instead of the FFMA / scalar fallbacks; same parity bar."""
    from eagcn_b200 import layers as EL
    assert EL.GraphConv_Layer.pad_widths                      # the default
    test_baseline_configs_forward_vs_oracle("config4 hiv widths 2-layer (padded)", "hiv", 12, 2, True)
    EL.GraphConv_Layer.pad_widths = False                     # and the un-padded fallback engines (FFMA GEMM, scalar kernels)
    try:
        test_baseline_configs_forward_vs_oracle("config4 hiv widths 2-layer (fallback engines)", "hiv", 12, 2, True)
    finally:
        EL.GraphConv_Layer.pad_widths = True


@pytest.mark.parametrize("case", golden_cases("stack_"))
@pytest.mark.parametrize("pad", [False, True])
def test_stack_vs_golden(case, pad):
    """BASELINE.json configs 3 / 4 against goldens from the reference itself: GraphConv_Layer stacked 3 / 2 deep +
    the head of models.py:108-120 (EAGCNStack), forward and every gradient; ``pad``: widths that are not multiples of 4
    through the padded layout (GraphConv_Layer.pad_widths)."""
    from eagcn_b200 import layers as EL, models as EM
    dev = _cuda()
    g = Golden(case)
    n_layers = int(g.meta["n_layers"])
    widths = [tuple(g.sd[f"layer{l + 1}.block{v + 1}.graph_conv.weight"].shape[1] for v in range(5)) for l in range(n_layers)]
    kb = g.sd["layer1.block1.att.weight"].shape[1]
    d1, d2, nc = g.sd["den1.weight"].shape[1], g.sd["den2.weight"].shape[1], g.sd["den3.weight"].shape[1]
    EL.GraphConv_Layer.pad_widths = pad
    try:
        model = EM.EAGCNStack(kb, 24, widths, d1, d2, nc, dropout=0.0).to(dev)
        model.load_state_dict(g.sd, strict=True)
        model.train()
        ins = _to(dev, g.dense())
        y, atom_rep, graph_rep = model(*ins, torch.from_numpy(g.batch.sizes).to(dev))
        assert rel_err(atom_rep.materialize(), g.out["atom_rep"]) <= n_layers * TOL
        assert rel_err(y.cpu(), g.out["y"]) <= 5 * TOL
        assert rel_err(graph_rep.cpu(), g.out["graph_rep"]) <= 5 * TOL
        (y * g.cot["y"].to(dev)).sum().backward()
        scale = max(float(v.abs().max()) for v in g.grad.values())
        named = dict(model.named_parameters())
        for k, ref in g.grad.items():
            got = named[k].grad
            assert got is not None, k
            bound = max(2e-4 * float(ref.abs().max()), 2e-5 * scale)
            if k.endswith("graph_conv.bias"):
                bound = 2e-4 * scale
            assert float((got.cpu() - ref).abs().max()) <= bound, k
    finally:
        EL.GraphConv_Layer.pad_widths = True


# ------------------------------------------------------------------ the BENCHMARKED configuration vs the oracle
def test_bench_config_vs_oracle():
    """bench.py's own step -- Tox21 B = 256, 24 -> 400 -> 700 + head 256/64/12, TRAIN mode, dropout 0.3, tcgen05 engine,
    side-stream branches on, captured into a CUDA graph and REPLAYED -- against the oracle on the same batch: the
    dropout keep masks of the three sites are exported from the library's Philox stream, the ReLU decisions are taken
    from the implementation under test (and bounded to the kink), everything else is the oracle's own arithmetic.
    Outputs <= 1e-5 per layer, every gradient <= 5e-5 of the gradient scale."""
    import bench
    from eagcn_b200 import functional as EF
    from eagcn_b200.plan import GraphPlan
    dev = _cuda()
    assert EF.Overlap.enabled
    model = bench.build_model(dev)
    hb, T, E = bench.host_batch(seed=3)
    dense = [torch.from_numpy(a) for a in hb.dense()]
    ins = _to(dev, dense)
    size = torch.from_numpy(hb.sizes).to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    p_drop, widths, nl = bench.P_DROP, bench.WIDTHS, len(bench.WIDTHS)
    rec = {}
    orig_bn_act = EF.bn_act

    def rec_bn_act(x, bn, training, relu=False, p_drop=0.0, rng_stream=1000):
        y = orig_bn_act(x, bn, training, relu=relu, p_drop=p_drop, rng_stream=rng_stream)
        rec.setdefault("bn_act", []).append(y)
        return y

    def step():
        for p in params:
            p.grad = None
        rec.clear()
        model.prefetch_params()
        plan = GraphPlan.build(ins[0], ins[2:], t_cap=T, e_cap=E)
        rec["plan"] = plan
        hooks = [l.register_forward_hook(lambda m, i, o, k=k: rec.__setitem__(f"x{k}", o[0].rows)) for k, l in
                 enumerate(model.conv_layers)]
        try:
            out, atom, grep = model(plan, ins[1], size=size)
        finally:
            for h in hooks:
                h.remove()
        out.sum().backward()
        rec["out"], rec["grep"] = out, grep
        return out

    work = torch.cuda.Stream()
    work.wait_stream(torch.cuda.current_stream())
    EF.bn_act = rec_bn_act
    try:
        with torch.cuda.stream(work):
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            rng = EF.RngState.get(dev)
            EF.manual_seed(4242, dev)
            rng_before = rng.state.clone()
            g.replay()
            g.replay()                       # second replay: same static buffers, the generator has moved on
            rng_used = rng_before.clone(); rng_used[1] += (nl + 1) << 20
            torch.cuda.synchronize()
    finally:
        EF.bn_act = orig_bn_act
    plan = rec["plan"].check()
    inc = 1 << 20
    snaps = [rng_used + torch.tensor([0, i * inc], device=dev) for i in range(nl + 1)]     # prefork: site i = offset + i*inc
    keeps, relus = [], []
    for l in range(nl):
        C = sum(widths[l])
        cfg = EF.LayerConfig(fin=0, fo=widths[l], training=True, p_drop=p_drop, rng_stream=l)
        keep = plan.scatter(EF.dropout_keep_mask(plan, cfg, C, snaps[l]).float()).cpu()
        xl = plan.scatter(rec[f"x{l}"].detach()).cpu()
        assert abs(float(keep.sum()) / (T * C) - (1 - p_drop)) < 0.01
        assert float((xl * (1 - keep)).abs().max()) == 0.0                  # dropped elements are exact zeros
        kv, rv, off = [], [], 0
        for w in widths[l]:
            kv.append(keep[:, :, off:off + w]); rv.append((xl[:, :, off:off + w] > 0).float() + (1 - keep[:, :, off:off + w]))
            off += w
        keeps.append(kv); relus.append(rv)
    D1 = bench.DEN[0]
    keep_h = EF.dropout_keep_mask_flat(snaps[nl], 1000, p_drop, hb.B * D1).view(hb.B, D1).float().cpu()
    a_gbn, a1, a2 = [t.detach().cpu() for t in rec["bn_act"]]
    relu_h = ((a1 > 0).float() + (1 - keep_h), (a2 > 0).float())
    sd = O.clone_sd(model.state_dict(), requires_grad=True)
    for k, t in sd.items():
        if "running" in k or "num_batches" in k:
            t.requires_grad_(False) if t.is_floating_point() else None
    codes = [O.codes_from_onehot(dense[0], r) for r in dense[2:]]
    h, outs = O.stack_forward(sd, dense[0], dense[1], codes, nl, True, p=p_drop, keeps=keeps, relu_masks=relus)
    y_ref, g_ref = O.head_forward(sd, h, torch.from_numpy(hb.sizes), True, p=p_drop, keep=keep_h, relu_masks=relu_h)
    # ReLU decisions of the implementation differ from the oracle's own only at the kink
    m = dense[0].max(2).values.unsqueeze(2)
    for l in range(nl):
        for v, Zv in enumerate(outs[l]["Z"]):
            dis = ((Zv.detach() > 0).float() != relus[l][v].clamp(max=1)) & (m > 0) & (keeps[l][v] > 0)
            if bool(dis.any()):
                assert float(Zv.detach().abs()[dis].max()) <= 1e-4 * float(Zv.detach().abs().max())
                assert int(dis.sum()) <= 1e-5 * dis.numel() + 8
    for l in range(nl):
        assert rel_err(plan.scatter(rec[f"x{l}"].detach()).cpu(), outs[l]["x"]) <= (l + 1) * TOL, f"layer {l + 1}"
    assert rel_err(rec["grep"].cpu(), g_ref) <= 3 * TOL
    assert rel_err(rec["out"].cpu(), y_ref) <= 3 * TOL
    y_ref.sum().backward()
    named = dict(model.named_parameters())
    scale = max(float(t.grad.abs().max()) for t in sd.values() if t.grad is not None)
    worst = 0.0
    for k, t in sd.items():
        if t.grad is None:
            continue
        got = named[k].grad
        assert got is not None, k
        err = float((got.cpu() - t.grad).abs().max()) / scale
        worst = max(worst, err)
        assert err <= 5e-5, (k, err)
    print(f"[bench-config parity] worst gradient error {worst:.2e} of scale")


# ------------------------------------------------------------------ fused layer kernel (layer_fused.cu)
def test_row_tiles_are_whole_molecules():
    """plan.tile_row: greedy runs of whole molecules, <= 128 active rows each, covering [0, T) in order."""
    from eagcn_b200.plan import GraphPlan
    from eagcn_b200.data import make_batch
    dev = _cuda()
    for seed, B, ds, fixed in ((0, 256, "tox21", None), (1, 64, "hiv", None), (2, 9, "tox21", 128), (3, 5, "tox21", 150)):
        batch = make_batch(B, ds, seed=seed, fixed_n=fixed)
        plan = GraphPlan.from_codes(torch.from_numpy(batch.codes).to(dev), batch.channels).check()
        cnt = plan.counts.cpu().tolist()
        T, nt, big = cnt[0], cnt[3], cnt[4]
        tr = plan.tile_row[:nt + 1].cpu().tolist()
        mp = plan.mol_ptr.cpu().tolist()
        assert tr[0] == 0 and tr[nt] == T and all(b > a for a, b in zip(tr, tr[1:]))
        assert all(b - a <= 128 for a, b in zip(tr, tr[1:]))
        assert big == (1 if max(b - a for a, b in zip(mp, mp[1:])) > 128 else 0)
        if not big:
            starts = set(mp)
            assert all(t in starts for t in tr)                       # every tile starts (and ends) at a molecule boundary
            # greedy: a tile cannot take the next molecule as well
            for k in range(nt - 1):
                nxt = min(m for m in mp if m > tr[k + 1])
                assert nxt - tr[k] > 128


@pytest.mark.parametrize("B,dataset,fin,fo,training,fixed", [
    (64, "tox21", 24, (80,) * 5, True, None),
    (48, "tox21", 400, (140,) * 5, True, None),
    (32, "lipo", 300, (100,) * 5, False, None),
    (16, "hiv", 24, (100, 52, 36, 20, 12), True, None),
    (6, "tox21", 24, (16, 12, 8, 8, 4), True, 128),          # every molecule fills a whole tile
    (300, "freesolv", 40, (260, 8, 8, 8, 4), True, None),    # a view wider than one 256-column chunk
])
def test_fused_forward_equals_two_kernel_form(B, dataset, fin, fo, training, fixed):
    """eagcn_layer_forward_a as ONE launch (tcgen05 projection with the attention / aggregation epilogue) against the
    projection GEMM + aggregation kernel pair: same arithmetic per element (Z tile identical, weights and row order
    identical), so the outputs agree to rounding of the statistics reduction order (<= 1e-6) and gradients likewise."""
    from eagcn_b200 import functional as EF, layers as EL, _lib
    from eagcn_b200.data import make_batch, DATASETS
    dev = _cuda()
    kb = DATASETS[dataset]["kb"]
    batch = make_batch(B, dataset=dataset, seed=B, kb=kb, n_afeat=fin, fixed_n=fixed)
    assert batch.N <= 128
    torch.manual_seed(B)
    layer = EL.GraphConv_Layer(fin, kb, *fo, dropout=0.3 if training else 0.0, structure="Concate").to(dev)
    layer.train(training)
    ins = _to(dev, [torch.from_numpy(a) for a in batch.dense()])
    res = []
    L = _lib.lib()
    with torch.no_grad():
        layer(*ins)                                                  # builds (and caches) the graph plan of this batch
    for fused in (1, 0):
        L.eagcn_set_fwd_fused(fused)
        try:
            c0 = _lib.launch_count()
            EF.manual_seed(7, dev)
            for prm in layer.parameters():
                prm.grad = None
            sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
            afm = ins[1].clone().requires_grad_(True)
            x, _ = layer(ins[0], afm, *ins[2:])
            x.square().sum().backward()
            torch.cuda.synchronize()
            res.append((x.detach().clone(), afm.grad.clone(), {n: q.grad.clone() for n, q in layer.named_parameters() if q.grad is not None},
                        {k: v.clone() for k, v in layer.state_dict().items()}, _lib.launch_count() - c0))
            layer.load_state_dict(sd0)
        finally:
            L.eagcn_set_fwd_fused(1)
    (x1, g1, pg1, sd1, n1), (x0, g0, pg0, sd0_, n0) = res
    assert n1 == n0 - 1                                              # one launch less per layer call
    assert rel_err(x1, x0) <= 5e-6
    assert rel_err(g1, g0) <= 2e-5
    for k in pg0:
        assert rel_err(pg1[k], pg0[k]) <= 5e-5 or float((pg1[k] - pg0[k]).abs().max()) <= 1e-6 * float(x0.abs().max()), k
    for k in sd0_:
        if sd0_[k].is_floating_point():
            assert rel_err(sd1[k], sd0_[k]) <= 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("B,dataset,fin,fo,training", [
    (64, "tox21", 30, (30, 10, 10, 10, 10), True),                  # 70 channels: one 128-channel block
    (256, "tox21", 60, (60, 20, 20, 20, 20), True),                 # 140 channels: two blocks, ~160 row tiles
    (5, "hiv", 20, (16, 8, 8, 8, 8), False),                        # eval: d bias = gamma * invstd * sum g
])
def test_bn_backward_last_cta_reduction(B, dataset, fin, fo, training):
    """The BatchNorm-backward sums reduced by the last CTA of each channel block (work.tickets) against the separate
    stat_reduce launch: same tile partials, both in double precision and fixed order (8 vs 32 tile lanes), so the
    gradients agree to the rounding of one double sum; one launch less; and two runs are bit-identical (the result does
    not depend on which CTA arrives last)."""
    from eagcn_b200 import functional as EF, layers as EL, _lib
    from eagcn_b200.data import make_batch, DATASETS
    dev = _cuda()
    kb = DATASETS[dataset]["kb"]
    batch = make_batch(B, dataset=dataset, seed=B + 1, kb=kb, n_afeat=fin)
    torch.manual_seed(B)
    layer = EL.GraphConv_Layer(fin, kb, *fo, dropout=0.2 if training else 0.0, structure="Concate").to(dev)
    layer.train(training)
    ins = _to(dev, [torch.from_numpy(a) for a in batch.dense()])
    with torch.no_grad():
        layer(*ins)
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    res = []
    old = EF._bwd_tickets_enabled
    try:
        for tickets in (True, True, False):
            EF._bwd_tickets_enabled = tickets
            layer.load_state_dict(sd0)
            EF.manual_seed(3, dev)
            for prm in layer.parameters():
                prm.grad = None
            afm = ins[1].clone().requires_grad_(True)
            c0 = _lib.launch_count()
            x, _ = layer(ins[0], afm, *ins[2:])
            x.square().sum().backward()
            torch.cuda.synchronize()
            res.append((afm.grad.clone(), {n: q.grad.clone() for n, q in layer.named_parameters() if q.grad is not None},
                        _lib.launch_count() - c0))
    finally:
        EF._bwd_tickets_enabled = old
    (ga, pa, na), (gb, pb, nb), (gc, pc, nc) = res
    assert na == nb == nc - 1
    assert torch.equal(ga, gb) and all(torch.equal(pa[k], pb[k]) for k in pa)
    assert rel_err(ga, gc) <= 1e-6
    scale = max(float(v.abs().max()) for v in pc.values())
    for k in pc:
        assert float((pa[k] - pc[k]).abs().max()) <= 1e-6 * max(float(pc[k].abs().max()), 1e-3 * scale), k
