"""TEST INFRASTRUCTURE ONLY -- the reference's OWN classes stacked n_layers deep.

reference models.py:50-61 hard-codes four GraphConv_Layers; BASELINE.json's "2-layer" / "3-layer" configurations
are therefore stacks of the same reference class.  ``RefStack`` wires unmodified reference classes
(``layers.GraphConv_Layer``, ``layers.Dense``, ``nn.BatchNorm1d``) exactly as ``EAGCN.__init__`` / ``EAGCN.forward``
do (models.py:50-61, :75-87, :96-121) for any depth; with four layers and the model's widths it IS models.EAGCN
(tests/test_reference_pins.py checks that, state_dict keys included).

Used by: tests/ (pins of the oracle), bench.py --impl reference and its cpu_baseline leg (the reference's CPU
path on the box's host cores).  Never imported by the product path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def make_ref_stack(L, n_bfeat, n_afeat, widths, n_den1, n_den2, nclass, dropout, molfp_mode="sum", last_flags=None):
    """L: the reference's ``layers`` module (oracle.ref_loader.load()[0])."""

    class RefStack(nn.Module):
        def __init__(self):
            super().__init__()
            fin = n_afeat
            self.n_layers = len(widths)
            for l, w in enumerate(widths):
                last = bool(last_flags[l]) if last_flags is not None else False
                setattr(self, f"layer{l + 1}", L.GraphConv_Layer(                       # models.py:50-61
                    node_feature_in=fin, bond_feature_num=n_bfeat, node_out_1=w[0], node_out_2=w[1], node_out_3=w[2],
                    node_out_4=w[3], node_out_5=w[4], dropout=dropout, structure="Concate", last=last))
                fin = sum(w)
            self.den1 = L.Dense(fin, n_den1)                                           # models.py:75-77
            self.den2 = L.Dense(n_den1, n_den2)
            self.den3 = L.Dense(n_den2, nclass)
            self.Graph_BN = nn.BatchNorm1d(fin)                                        # models.py:79-87
            self.bn_den1 = nn.BatchNorm1d(n_den1)
            self.bn_den2 = nn.BatchNorm1d(n_den2)
            self.dropout, self.molfp_mode = dropout, molfp_mode

        def forward(self, adjs, afms, TypeAtt, OrderAtt, AromAtt, ConjAtt, RingAtt, size):
            x2 = afms
            for l in range(self.n_layers):                                             # models.py:97-100
                x2, A = getattr(self, f"layer{l + 1}")(adjs, x2, TypeAtt, OrderAtt, AromAtt, ConjAtt, RingAtt)
            atom_representations = x2.data.cpu()                                       # models.py:102
            x = torch.sum(x2, 1)                                                       # models.py:108
            if self.molfp_mode == "ave":                                               # models.py:109-111
                x = x / size.view(-1, 1).to(x.dtype)
            x = self.Graph_BN(x)                                                       # models.py:112
            x = self.den1(x)
            x = F.relu(self.bn_den1(x))
            x = F.dropout(x, p=self.dropout, training=self.training)
            x = self.den2(x)
            graph_representation = x
            x = F.relu(self.bn_den2(x))
            x = self.den3(x)
            return x, atom_representations, graph_representation

    return RefStack()


def seeded_init(model, seed=0):
    """train.py:302 ``model.apply(weights_init)`` (utils.py:702-708) with a private generator: GraphConv_base weights
    ~ N(0, 0.02), every *BatchNorm* weight ~ N(1, 0.02) / bias 0 -- and finite values for the parameters the
    reference leaves uninitialised (AFM_BatchNorm.weight / .bias, layers.py:402-404)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            name = mod.__class__.__name__
            if "GraphConv_base" in name:
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.02)
            elif "BatchNorm" in name:
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.02 + 1.0)
                mod.bias.zero_()
        for p in model.parameters():
            if not torch.isfinite(p).all():
                p.fill_(0.0)
    return model
