"""Pins of the test infrastructure against the LIVE reference (marker ``reference``: /root/reference in the build
container, the shipped copy oracle/_ref on the GPU box -- tools/make_oracle_ref.sh).

  * O.weight_tensor_loop / O.bce_loss_loop / eagcn_b200.losses  vs  utils.weight_tensor (utils.py:653-679) and the
    loss expression of train.py:326-331;
  * O.model_forward_conv (bench.py's "port" CPU arm)            vs  the reference classes, outputs AND gradients;
  * oracle.ref_stack.RefStack (bench.py's "reference" CPU arm)  vs  models.EAGCN (models.py:14-121), so the stacked
    form used for the 2- / 3-layer configurations is the reference's own wiring.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import eagcn_oracle as O
from oracle import ref_loader
from oracle.ref_stack import make_ref_stack, seeded_init
from tests.util import rel_err

pytestmark = pytest.mark.reference


def _labels(B=23, T=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 2, (B, T), generator=g).float()
    labels[torch.rand(B, T, generator=g) < 0.3] = -1.0          # missing label as the tox21 CSV loader encodes it
    labels[0, 0] = float("nan")                                  # and as a NaN (int(nan) -> ValueError branch)
    weights = {j: [5000.0 / (10 + 3 * j), 5000.0 / (200 + j)] for j in range(T)}
    return labels, weights


def test_weight_tensor_vs_live_reference():
    from eagcn_b200 import losses
    with ref_loader.cpu_only():
        _, _, U = ref_loader.load()
        labels, weights = _labels()
        w_ref = U.weight_tensor(weights, labels)                 # utils.py:653-679, executed
    assert torch.equal(O.weight_tensor_loop(weights, labels), w_ref)
    table = losses.bce_weight_table(weights, labels.shape[1])
    assert torch.equal(losses.label_weights(table, labels).reshape(-1), w_ref)


def test_bce_loss_vs_live_reference():
    from eagcn_b200 import losses
    with ref_loader.cpu_only():
        _, _, U = ref_loader.load()
        labels, weights = _labels(seed=1)
        labels = torch.nan_to_num(labels, nan=-1.0)              # train.py feeds labels.float() to BCE: NaN would poison it
        out = torch.randn(labels.shape, generator=torch.Generator().manual_seed(2))
        o_ref = out.clone().requires_grad_(True)
        w = U.weight_tensor(weights, labels)
        non_nan = torch.FloatTensor([(labels == 1).sum() + (labels == 0).sum()])
        # train.py:326-331 (size_average=False is reduction='sum' on current torch)
        loss_ref = F.binary_cross_entropy_with_logits(o_ref.view(-1), labels.float().view(-1), weight=w,
                                                      reduction="sum") / non_nan
        loss_ref.backward()
    o1 = out.clone().requires_grad_(True)
    l1 = O.bce_loss_loop(o1, labels, weights)
    l1.backward()
    o2 = out.clone().requires_grad_(True)
    l2 = losses.weighted_bce_with_logits(o2, labels, losses.bce_weight_table(weights, labels.shape[1]))
    l2.backward()
    # the reference feeds label -1 into BCE with weight 0: value and gradient of those elements are exactly 0
    for l, o in ((l1, o1), (l2, o2)):
        assert abs(float(l) - float(loss_ref)) <= 1e-6 * abs(float(loss_ref))
        assert rel_err(o.grad, o_ref.grad) <= 1e-6


def _batch(B, dataset, seed, kb):
    from eagcn_b200.data import make_batch
    b = make_batch(B, dataset=dataset, seed=seed, kb=kb)
    return b, [torch.from_numpy(a) for a in b.dense()], torch.from_numpy(b.sizes)


@pytest.mark.parametrize("training,p", [(True, 0.0), (False, 0.0), (True, 0.3)])
def test_model_forward_conv_vs_live_reference(training, p):
    """bench.py's 'port' arm: same op sequence as the reference classes -> same values (and, with dropout, the same
    torch-generator draws), outputs and every gradient."""
    with ref_loader.cpu_only():
        L, _, _ = ref_loader.load()
        torch.manual_seed(0)
        widths = [(10, 9, 8, 7, 6), (12, 8, 8, 4, 4)]
        ref = seeded_init(make_ref_stack(L, 30, 24, widths, 16, 8, 3, dropout=p), seed=3)
        ref.train(training)
        batch, dense, sizes = _batch(9, "tox21", 4, 30)
        torch.manual_seed(11)
        y_ref, atom_ref, _ = ref(*dense, sizes)
        y_ref.sum().backward()
    sd = O.clone_sd(ref.state_dict())
    for k, t in sd.items():
        if t.is_floating_point() and "running" not in k:
            t.requires_grad_(True)
    torch.manual_seed(11)
    y = O.model_forward_conv(sd, dense[0], dense[1], dense[2:], sizes, 2, training, p)
    y.sum().backward()
    assert torch.equal(y.detach(), y_ref.detach())
    for k, prm in ref.named_parameters():
        if prm.grad is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
        else:
            assert torch.equal(sd[k].grad, prm.grad), k


def test_ref_stack_is_models_eagcn():
    """Four layers at models.EAGCN's widths: RefStack == the reference model (same keys, same outputs, same grads)."""
    with ref_loader.cpu_only():
        L, M, U = ref_loader.load()
        torch.manual_seed(0)
        s1, s2 = 6, 10
        model = M.EAGCN(30, 24, *([s1] * 5), *([s2] * 5), 16, 8, 2, dropout=0.0, structure="Concate", molfp_mode="sum")
        seeded_init(model, seed=1)
        widths = [(s1,) * 5, (s2,) * 5, (2 * s2,) * 5, (2 * s2,) * 5]
        stack = make_ref_stack(L, 30, 24, widths, 16, 8, 2, dropout=0.0, last_flags=[0, 0, 0, 1])
        assert set(stack.state_dict().keys()) == set(model.state_dict().keys())
        stack.load_state_dict(model.state_dict(), strict=True)
        batch, dense, sizes = _batch(6, "tox21", 2, 30)
        model.train(); stack.train()
        y0, a0, g0 = model(*dense, sizes)
        y1, a1, g1 = stack(*dense, sizes)
        assert torch.equal(y0, y1) and torch.equal(a0, a1) and torch.equal(g0, g1)
        y0.sum().backward(); y1.sum().backward()
        for (k, p0), (_, p1) in zip(model.named_parameters(), stack.named_parameters()):
            assert (p0.grad is None) == (p1.grad is None), k
            if p0.grad is not None:
                assert torch.equal(p0.grad, p1.grad), k


@pytest.mark.parametrize("dataset,kb,molfp,training,B,seed", [
    ("tox21", 30, "sum", True, 9, 0),
    ("tox21", 30, "ave", False, 5, 1),
    ("lipo", 18, "pool", True, 6, 2),
    ("hiv", 30, "sum", False, 4, 3),
    ("freesolv", 17, "ave", True, 8, 4),
])
def test_lookup_oracle_full_model_vs_live_reference(dataset, kb, molfp, training, B, seed):
    """The oracle the GPU parity tests compare against -- O.stack_forward + O.head_forward, the LOOKUP form on uint8 edge
    codes -- against the unmodified models.EAGCN (models.py:14-121): outputs, graph representation, atom representations
    and EVERY parameter gradient, in training and eval mode, for the three read-outs and four data sets' bond vocabularies
    (batches include ragged sizes, padded rows and molecules whose atoms have a single bond)."""
    with ref_loader.cpu_only():
        _, M, _ = ref_loader.load()
        torch.manual_seed(seed)
        s1, s2 = 8, 12
        ref = M.EAGCN(30, 24, *([s1] * 5), *([s2] * 5), 16, 8, 3, dropout=0.0, structure="Concate", molfp_mode=molfp)
        # models.EAGCN is written for Tox21's bond vocabulary sizes; the data sets differ only in the channel counts
        batch, dense, sizes = _batch(B, dataset, seed + 10, kb)
        chans = [int(r.shape[1]) for r in dense[2:]]
        for l, layer in enumerate((ref.layer1, ref.layer2, ref.layer3, ref.layer4)):
            for v, blk in enumerate((layer.block1, layer.block2, layer.block3, layer.block4, layer.block5)):
                if blk.att.weight.shape[1] != chans[v]:
                    blk.att = torch.nn.Conv2d(chans[v], 1, kernel_size=1, stride=1, padding=0, bias=False)
        seeded_init(ref, seed=seed + 1)
        ref.train(training)
        if not training:                                          # eval mode normalises with the running statistics
            with torch.no_grad():
                for k, t in ref.state_dict().items():
                    if k.endswith("running_mean"):
                        t.copy_(0.1 * torch.randn(t.shape, generator=torch.Generator().manual_seed(len(k))))
                    elif k.endswith("running_var"):
                        t.copy_(0.5 + torch.rand(t.shape, generator=torch.Generator().manual_seed(len(k) + 1)))
        y_ref, atom_ref, g_ref = ref(*dense, sizes)
        cot = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(7))
        (y_ref * cot).sum().backward()
    sd = O.clone_sd(ref.state_dict())
    for k, t in sd.items():
        if t.is_floating_point() and "running" not in k:
            t.requires_grad_(True)
    adj, afm = dense[0], dense[1]
    codes = [O.codes_from_onehot(adj, r) for r in dense[2:]]
    h, outs = O.stack_forward(sd, adj, afm, codes, 4, training, last_flags=[0, 0, 0, 1])
    y, g = O.head_forward(sd, h, sizes, training, molfp_mode=molfp, A_last=outs[-1]["A_weight"])
    (y * cot).sum().backward()
    assert rel_err(h.detach(), atom_ref.detach()) <= 2e-5
    assert rel_err(y.detach(), y_ref.detach()) <= 5e-5 and rel_err(g.detach(), g_ref.detach()) <= 5e-5
    scale = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    checked = 0
    for k, prm in ref.named_parameters():
        if prm.grad is None:
            continue
        got = sd[k].grad
        assert got is not None, k
        denom = max(float(prm.grad.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and training:
            denom = scale                                         # d bias through a train-mode BatchNorm: analytically zero
        assert float((got.reshape(prm.grad.shape) - prm.grad).abs().max()) / denom <= 2e-4, k
        checked += 1
    assert checked >= 100                                         # 4 layers x 5 views x (att, self_r, W, bias, gamma, beta) + head
