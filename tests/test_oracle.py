"""CPU: the oracle restatement vs golden vectors produced by the unmodified reference, and vs the
reference itself when it is present (build container).  Tolerance: fp32, max|d|/max|ref| <= 1e-5
(north_star); the indexing part (codes) is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import eagcn_oracle as O
from tests.util import Golden, golden_cases, rel_err

TOL = 1e-5


def _layer_oracle(g, dtype=torch.float32):
    sd = O.clone_sd({("layer1." + k): v for k, v in g.sd.items()}, dtype=dtype, requires_grad=True)
    adj, afm, *rels = g.dense()
    codes = [O.codes_from_onehot(adj, r) for r in rels]
    afm = afm.to(dtype).requires_grad_(True)
    out = O.layer_forward(sd, "layer1.", adj.to(dtype), afm, codes, bool(g.meta["training"]),
                          structure=str(g.meta["structure"]), last=bool(g.meta["last"]))
    return sd, afm, codes, out


@pytest.mark.parametrize("case", golden_cases("layer_"))
def test_codes_bit_exact(case):
    g = Golden(case)
    adj, afm, *rels = g.dense()
    for v, r in enumerate(rels):
        code = O.codes_from_onehot(adj, r)
        assert torch.equal(code, g.codes_i64()[v])                      # indexing: bit-exact


@pytest.mark.parametrize("case", golden_cases("layer_"))
def test_layer_forward_backward_vs_golden(case):
    g = Golden(case)
    sd, afm, codes, out = _layer_oracle(g)
    assert rel_err(out["x"], g.out["x"]) <= TOL
    assert rel_err(out["A_weight"], g.out["A"]) <= TOL
    loss = (out["x"] * g.cot["x"]).sum() + (out["A_weight"] * g.cot["A"]).sum()
    loss.backward()
    assert rel_err(afm.grad, g.grad["afm"]) <= TOL
    scale = max(float(v.abs().max()) for v in g.grad.values())
    for k, ref in g.grad.items():
        if k == "afm":
            continue
        got = sd["layer1." + k].grad
        assert got is not None, k
        # graph_conv.bias grads are analytically 0 through BatchNorm (pure rounding noise) ->
        # compare against the overall gradient scale
        denom = max(float(ref.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and bool(g.meta["training"]):
            denom = scale           # sum of BN input-grads: exactly 0 in exact arithmetic
        assert float((got - ref).abs().max()) / denom <= 5 * TOL, k
    # padded rows of x are exactly zero (layers.py:313)
    m = O.row_mask(g.dense()[0])
    if str(g.meta["structure"]) == "Concate":
        assert float((out["x"].detach() * (1 - m).unsqueeze(2)).abs().max()) == 0.0


@pytest.mark.parametrize("case", ["layer_train", "layer_wide"])
def test_running_stats_update(case):
    g = Golden(case)
    sd, afm, codes, out = _layer_oracle(g)
    for v in range(5):
        rm, rv = out["stats"][v]
        assert rel_err(rm, g.post[f"block{v + 1}.batch_norm.bn.running_mean"]) <= TOL
        assert rel_err(rv, g.post[f"block{v + 1}.batch_norm.bn.running_var"]) <= TOL


@pytest.mark.parametrize("case", golden_cases("model_"))
def test_model_vs_golden(case):
    g = Golden(case)
    sd = O.clone_sd(g.sd, requires_grad=True)
    adj, afm, *rels = g.dense()
    codes = [O.codes_from_onehot(adj, r) for r in rels]
    training = bool(g.meta["training"])
    structure = str(g.meta["structure"]) if "structure" in g.meta else "Concate"
    h, outs = O.stack_forward(sd, adj, afm, codes, 4, training, structure=structure, last_flags=[0, 0, 0, 1])
    y, grep = O.head_forward(sd, h, torch.from_numpy(g.batch.sizes), training, molfp_mode=str(g.meta["molfp"]),
                             A_last=outs[-1]["A_weight"])
    assert rel_err(h, g.out["atom_rep"]) <= TOL
    assert rel_err(y, g.out["y"]) <= 2 * TOL
    assert rel_err(grep, g.out["graph_rep"]) <= 2 * TOL
    (y * g.cot["y"]).sum().backward()
    scale = max(float(v.abs().max()) for v in g.grad.values())
    for k, ref in g.grad.items():
        got = sd[k].grad
        assert got is not None, k
        denom = max(float(ref.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and training:
            denom = scale
        # 4 stacked train-mode BatchNorms amplify fp32 reassociation noise; still ~1e-5 class
        assert float((got - ref).abs().max()) / denom <= 2e-4, k


@pytest.mark.parametrize("case", golden_cases("stack_"))
def test_stack_vs_golden(case):
    """BASELINE.json's 3-layer (Lipophilicity, regression head) and 2-layer (HIV widths, not multiples of 4)
    configurations: the reference's GraphConv_Layer stacked 3 / 2 deep + the head of models.py:108-120."""
    g = Golden(case)
    sd = O.clone_sd(g.sd, requires_grad=True)
    adj, afm, *rels = g.dense()
    codes = [O.codes_from_onehot(adj, r) for r in rels]
    n_layers = int(g.meta["n_layers"])
    h, _ = O.stack_forward(sd, adj, afm, codes, n_layers, True)
    y, grep = O.head_forward(sd, h, torch.from_numpy(g.batch.sizes), True)
    assert rel_err(h, g.out["atom_rep"]) <= TOL
    assert rel_err(y, g.out["y"]) <= 2 * TOL
    assert rel_err(grep, g.out["graph_rep"]) <= 2 * TOL
    (y * g.cot["y"]).sum().backward()
    scale = max(float(v.abs().max()) for v in g.grad.values())
    for k, ref in g.grad.items():
        got = sd[k].grad
        assert got is not None, k
        # 2e-4 of the tensor's own largest gradient, but never tighter than 5e-6 of the model's gradient scale (fp32
        # reassociation noise through 2-3 stacked train-mode BatchNorms; d bias through a BatchNorm is noise around 0)
        bound = max(2e-4 * float(ref.abs().max()), 5e-6 * scale)
        if k.endswith("graph_conv.bias"):
            bound = 2e-4 * scale
        assert float((got - ref).abs().max()) <= bound, k


def test_non_onehot_rejected():
    g = Golden("layer_train")
    adj, afm, *rels = g.dense()
    bad = rels[0].clone()
    b, i, j = [int(t[0]) for t in torch.nonzero(adj, as_tuple=True)]
    bad[b, :, i, j] = 0.5
    with pytest.raises(ValueError):
        O.codes_from_onehot(adj, bad)


@pytest.mark.reference
def test_oracle_vs_live_reference():
    """Fresh seeded inputs through the unmodified reference classes, here and now."""
    from oracle import ref_loader
    from eagcn_b200.data import make_batch
    L, _, _ = ref_loader.load()
    torch.manual_seed(3)
    batch = make_batch(7, dataset="lipo", seed=5, kb=18)
    layer = L.GraphConv_Layer(24, 18, 10, 9, 8, 7, 6, dropout=0.0, structure="Concate")
    for p in layer.parameters():
        if not torch.isfinite(p).all():
            p.data.fill_(1.0)
    layer.train()
    ins = [torch.from_numpy(a) for a in batch.dense()]
    sd = O.clone_sd({("layer1." + k): v for k, v in layer.state_dict().items()})
    x_ref, A_ref = layer(*ins)
    codes = [O.codes_from_onehot(ins[0], r) for r in ins[2:]]
    out = O.layer_forward(sd, "layer1.", ins[0], ins[1], codes, True)
    assert rel_err(out["x"], x_ref.detach()) <= TOL
    assert rel_err(out["A_weight"], A_ref.detach()) <= TOL
