"""Mirror of the reference's ``models.py`` (EAGCN, models.py:14-121) on the CUDA layer path.

``EAGCN`` keeps the constructor arguments, sub-module names (layer1..layer4, den1..den3, Graph_BN,
bn_den1, bn_den2), state_dict keys and the 3-tuple return of the reference.  Differences, all about
not doing dead work (results are unchanged):
  * one GraphPlan is built per batch and the atom features stay in packed-row form between layers
    (the reference re-reads the dense padded tensors and all one-hot planes in each layer, models.py:97-100);
  * the dense attention stack is not materialised for the 'sum' / 'ave' read-outs (only 'pool' uses it);
  * ``atom_representations`` (models.py:102 does a blocking ``x2.data.cpu()`` every forward) is returned
    as a lazy handle that performs the scatter + D2H copy only when something actually reads it.

``EAGCNStack`` is the same model with a configurable number of GraphConv_Layers (BASELINE.json's
"2-layer" / "3-layer" configurations; the reference hard-codes 4, models.py:50-61).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import functional as EF
from ._lib import EagcnError
from .layers import Diff_Pooling, GraphConv_Layer, PackedRows, _param_device
from .plan import GraphPlan


class Dense(nn.Module):
    """layers.py:360-392: bias-free fully connected layer (weight [in,out])."""

    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        dev = _param_device()
        self.weight = Parameter(torch.empty(in_features, out_features, device=dev))
        if bias:
            self.bias = Parameter(torch.empty(out_features, device=dev))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    # 'tile' : eagcn_mm_tile (mm_tile.cu: 32x32 tiles, whole K range staged at once, split-K combined inside the launch,
    #          weight gradient on the side stream) -- strict fp32, bit-reproducible;
    # 'torch': library GEMM (un-split 32x32x16 SIMT kernel on these skinny shapes: 63 us per step over 9 products)
    mm_engine = "tile"

    def forward(self, input):
        if self.mm_engine == "tile" and input.is_cuda and input.dim() == 2 and input.dtype == torch.float32:
            out = EF.dense_mm(input, self.weight, self.mm_engine)                   # layers.py:382-388
        else:
            out = torch.mm(input, self.weight)
        return out + self.bias if self.bias is not None else out


class LazyAtomRep:
    """Stand-in for ``x2.data.cpu()`` (models.py:102): materialises on first use."""

    def __init__(self, packed):
        # PackedRows (Concate: dense() re-inserts the exact zeros of padded rows) or an already dense tensor
        self._packed = PackedRows(packed.rows.detach(), packed.plan) if isinstance(packed, PackedRows) else packed.detach()
        self._value = None

    def materialize(self):
        if self._value is None:
            src = self._packed
            self._value = (src.dense() if isinstance(src, PackedRows) else src).cpu()
            self._packed = None
        return self._value

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    def __getitem__(self, i):
        return self.materialize()[i]


class EAGCNStack(nn.Module):
    """``n_layers`` GraphConv_Layers + sum/ave read-out + the reference's dense head.

    widths: per layer the 5 per-view output widths, e.g. [(80,)*5, (140,)*5] for the Tox21 2-layer config.
    """

    def __init__(self, n_bfeat, n_afeat, widths, n_den1, n_den2, nclass, dropout, molfp_mode="sum",
                 last_flags=None, structure="Concate", pool_num=5):
        super().__init__()
        if structure not in ("Concate", "Weighted_sum"):
            raise EagcnError("EAGCNStack implements structure 'Concate' and 'Weighted_sum' (models.py:30-61); GCN / GAT "
                             "are the reference's comparison baselines and stay on stock PyTorch")
        self.structure = structure
        if molfp_mode not in ("sum", "ave", "pool"):
            raise EagcnError("read-out modes: 'sum' / 'ave' / 'pool' (models.py:104-111)")
        self.molfp_mode, self.dropout = molfp_mode, dropout
        self.head_bn = "cuda"          # 'cuda': fused BatchNorm/ReLU/dropout kernels; 'torch': stock modules
        fin = n_afeat
        self.n_layers = len(widths)
        for l, w in enumerate(widths):
            last = bool(last_flags[l]) if last_flags is not None else False
            layer = GraphConv_Layer(fin, n_bfeat, *w, dropout=dropout, structure=structure, last=last)
            layer.materialize_A = False
            layer.rng_stream = l
            setattr(self, f"layer{l + 1}", layer)
            fin = layer.total_output                                                # sum(w) | w[0] (layers.py:277-281)
        self.out_width = fin
        if molfp_mode == "pool":                                                    # models.py:90-92
            if structure != "Concate" or last_flags is None or not last_flags[-1]:
                raise EagcnError("molfp_mode='pool' needs structure='Concate' and last=True on the final layer")
            self.pool1 = Diff_Pooling(fin, fin, pool_num)
            self.pool3 = Diff_Pooling(fin, fin, 1)            # constructed but unused by the reference's forward
            layer.materialize_A = True                        # Diff_Pooling consumes the final layer's attention
        self.den1 = Dense(fin, n_den1)
        self.den2 = Dense(n_den1, n_den2)
        self.den3 = Dense(n_den2, nclass)
        dev = _param_device()
        self.Graph_BN = nn.BatchNorm1d(fin).to(dev)
        self.bn_den1 = nn.BatchNorm1d(n_den1).to(dev)
        self.bn_den2 = nn.BatchNorm1d(n_den2).to(dev)

    @property
    def conv_layers(self):
        return [getattr(self, f"layer{l + 1}") for l in range(self.n_layers)]

    def prefetch_params(self):
        """Optional hint, called BEFORE the batch's GraphPlan is built: launches everything of the coming forward pass
        that depends only on the parameters (per-layer W_all concat / split / sigmoid tables, the dropout-generator
        fork of every dropout site) on a side stream, so that it runs beside the packing kernels.  ``forward`` picks
        the result up (and falls back to preparing in line when the parameters changed in between).  ``forward``
        calls it itself when it is handed dense tensors and has to build the plan."""
        if not EF.Overlap.enabled or self.structure != "Concate":
            return
        dev = self.den1.weight.device
        if dev.type != "cuda":
            return
        side = EF.Overlap.fork_fwd(dev)
        # side-stream order = order of first use: layer 1's tables (its forward kernel is the first consumer), the dropout
        # snapshots (first used after that kernel), then the later layers' tables (they run beside layer 1's forward pass);
        # every consumer waits for its own item only (LayerPrep.ready / RngState._queue_ready)
        layers = list(self.conv_layers)
        preps = [layers[0].prepare(side)]
        if self.training and float(self.dropout) > 0.0:
            EF.RngState.get(dev).prefork(self.n_layers + 1, side)       # one site per layer + the head's dropout
        preps += [layer.prepare(side) for layer in layers[1:]]
        self._prefetched = preps

    def forward(self, adjs, afms, TypeAtt=None, OrderAtt=None, AromAtt=None, ConjAtt=None, RingAtt=None, size=None):
        if isinstance(adjs, GraphPlan):
            plan = adjs
        else:
            if getattr(self, "_prefetched", None) is None and torch.is_tensor(adjs) and adjs.is_cuda:
                self.prefetch_params()
            plan = GraphConv_Layer._plan_for(adjs, (TypeAtt, OrderAtt, AromAtt, ConjAtt, RingAtt))
        preps = getattr(self, "_prefetched", None) or [None] * self.n_layers
        self._prefetched = None
        if self.structure == "Weighted_sum":
            # the un-masked padded rows of this structure (layers.py:314-316) travel between layers and into the
            # read-out sum exactly as in the reference: dense tensors end to end
            if isinstance(afms, PackedRows):
                raise EagcnError("structure='Weighted_sum' takes dense [B,N,F] atom features")
            h = afms
            for layer in self.conv_layers:                                          # models.py:97-100
                h, _ = layer(plan, h)
            atom_representations = LazyAtomRep(h)                                   # models.py:102
            x = h.sum(1)                                                            # models.py:108 (padded rows included)
        else:
            h = afms if isinstance(afms, PackedRows) else PackedRows(EF.gather_rows(plan, afms), plan)
            for layer, prep in zip(self.conv_layers, preps):                        # models.py:97-100
                h, A = layer(plan, h, prep=prep)
            EF.Overlap.join_fwd(plan.device)                                        # no-op once a layer consumed the prefetch
            atom_representations = LazyAtomRep(h)                                   # models.py:102
            if self.molfp_mode == "pool":                                           # models.py:104-106
                _, xp = self.pool1(A, h.dense())
                x = xp.sum(1)
            else:
                x = EF.readout_sum(plan, h.rows)                                    # models.py:108
        if self.molfp_mode == "ave":                                                # models.py:109-111
            x = x / size.view(-1, 1).to(x.dtype)
        bns = (self.Graph_BN, self.bn_den1, self.bn_den2)
        if self.head_bn == "cuda" and all(bn.affine and bn.track_running_stats and bn.momentum is not None for bn in bns):
            # library GEMMs between fused BatchNorm(+ReLU)(+dropout) kernels: 3 + 3 launches instead of ~25
            x = EF.bn_act(x, self.Graph_BN, self.training)                          # models.py:112
            x = self.den1(x)
            x = EF.bn_act(x, self.bn_den1, self.training, relu=True, p_drop=float(self.dropout))   # models.py:114-116
            x = self.den2(x)
            graph_representation = x
            x = EF.bn_act(x, self.bn_den2, self.training, relu=True)                # models.py:119
            x = self.den3(x)
            return x, atom_representations, graph_representation
        if self.head_bn == "sync":
            # global-batch BatchNorm across data-parallel ranks (eagcn_b200.parallel.set_bn_sync): the head's three
            # BatchNorm1d see the GLOBAL batch like the reference's single process does (models.py:112,115,119)
            gbn, bn1, bn2 = (self._sync_bn(bn) for bn in bns)
        else:
            gbn, bn1, bn2 = bns
        x = gbn(x)                                                                  # models.py:112
        x = self.den1(x)
        x = F.relu(bn1(x))
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.den2(x)
        graph_representation = x
        x = F.relu(bn2(x))
        x = self.den3(x)
        return x, atom_representations, graph_representation

    def _sync_bn(self, bn):
        """torch.nn.SyncBatchNorm over the SAME parameter / buffer tensors as ``bn`` (state_dict keys unchanged); the
        wrappers live outside the module tree."""
        cache = self.__dict__.setdefault("_sync_bn_cache", {})
        sbn = cache.get(id(bn))
        if sbn is None:
            sbn = nn.SyncBatchNorm(bn.num_features, bn.eps, bn.momentum, bn.affine, bn.track_running_stats)
            sbn.weight, sbn.bias = bn.weight, bn.bias
            sbn.running_mean, sbn.running_var, sbn.num_batches_tracked = bn.running_mean, bn.running_var, bn.num_batches_tracked
            cache[id(bn)] = sbn
        sbn.train(self.training)
        return sbn


class EAGCN(EAGCNStack):
    """Reference constructor signature (models.py:22-25); structure 'Concate', molfp 'sum' | 'ave'."""

    def __init__(self, n_bfeat, n_afeat, n_sgc1_1, n_sgc1_2, n_sgc1_3, n_sgc1_4, n_sgc1_5,
                 n_sgc2_1, n_sgc2_2, n_sgc2_3, n_sgc2_4, n_sgc2_5, n_den1, n_den2, nclass, dropout,
                 structure="Concate", molfp_mode="sum", pool_num=5):
        l1 = (n_sgc1_1, n_sgc1_2, n_sgc1_3, n_sgc1_4, n_sgc1_5)
        l2 = (n_sgc2_1, n_sgc2_2, n_sgc2_3, n_sgc2_4, n_sgc2_5)
        if structure == "Weighted_sum":                                             # models.py:33-47: every view full width
            l1, l2 = (sum(l1),) * 5, (sum(l2),) * 5
        l3 = tuple(2 * w for w in l2)                                               # models.py:56-61
        super().__init__(n_bfeat, n_afeat, [l1, l2, l3, l3], n_den1, n_den2, nclass, dropout, molfp_mode,
                         last_flags=[False, False, False, True], structure=structure, pool_num=pool_num)
        if structure == "Weighted_sum":
            self.ngc1, self.ngc2 = l1[0], l2[0]
            return
        self.ngc1, self.ngc2 = sum(l1), sum(l2)
