"""Drop-in mirror of the reference's ``layers.py`` module surface for the EAGCN hot path.

Same class names, constructor arguments, attribute names, parameter shapes and ``state_dict`` keys as
reference eagcn_pytorch/layers.py (GraphConv_Layer :262-325, GraphConv_block :52-95, GraphConv_base
:16-50, AFM_BatchNorm :394-412, Ave_multi_view :414-437), so that ``models.EAGCN`` / ``train.py`` /
``check_model.py`` (which reads ``model.layerN.blockM.self_r`` and ``.att.weight``) and
``utils.weights_init`` (which matches the class-name substrings 'GraphConv_base' and 'BatchNorm') keep
working -- but ``GraphConv_Layer.forward`` runs on hand-written sm_100a CUDA kernels through the C ABI
(include/eagcn_b200.h) instead of ~60 ATen ops.  There is no CPU path: CPU tensors raise.

Extensions over the reference signature (all optional, defaults reproduce the reference):
  * ``adjs`` may be a prebuilt ``GraphPlan`` (then the relation arguments are ignored) and ``afms`` may be
    a ``PackedRows`` -- the layer then returns ``PackedRows`` and skips the dense padded round trip;
  * ``GraphConv_Layer.materialize_A`` (default True): build the dense attention stack the reference
    returns (layers.py:318); models that discard it (sum/ave read-out) switch it off.
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import functional as EF
from ._lib import EagcnError
from .plan import GraphPlan

_DEV = None


def _param_device():
    """The reference places parameters on the GPU when one is visible (layers.py:10-14)."""
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class PackedRows:
    """Active atom rows of a batch: ``rows`` [t_cap, F] (row order of ``plan``)."""
    __slots__ = ("rows", "plan")

    def __init__(self, rows, plan):
        self.rows, self.plan = rows, plan

    def dense(self):
        """[B,N,F] with exact zeros on padded / bond-less rows (layers.py:313 '* mask3')."""
        return EF.scatter_rows(self.plan, self.rows)

    @property
    def shape(self):
        return (self.plan.B, self.plan.N, self.rows.shape[1])


class GraphConv_base(nn.Module):
    """Parameter holder of the projection (layers.py:16-50): weight [in,out], bias [out].

    Its own ``forward`` (bmm + mm, layers.py:38-45) is kept for stand-alone users such as Diff_Pooling;
    inside GraphConv_Layer the projection is part of the fused CUDA path.
    """

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

    def reset_parameters(self):                           # layers.py:32-36
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, adjs, afms):                        # layers.py:38-45
        support = torch.bmm(adjs, afms)
        out = torch.mm(support.reshape(-1, self.in_features), self.weight).view(-1, adjs.shape[1], self.out_features)
        return out + self.bias if self.bias is not None else out

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


class Diff_Pooling(nn.Module):
    """layers.py:492-506 (the definition in effect): soft assignment of the atoms to ``out_size`` clusters from the
    last layer's normalised attention A [B,N,N] and atom features X [B,N,F].  Not on the hot path: two bmm + two mm on
    one dense [B,N,N] map, composed from library ops; A comes from the CUDA path (eagcn_attention_dense + the
    'last' normalisation of GraphConv_Layer) and carries the gradients of self_r / ave_A / att."""

    def __init__(self, in_feature, out_feature, out_size):
        super().__init__()
        self.feature_layer = GraphConv_base(in_feature, out_feature)
        self.adjacent_layer = GraphConv_base(in_feature, out_size)

    def forward(self, A, X):
        X_feature = F.relu(self.feature_layer(A, X))                                  # layers.py:499
        S = F.softmax(self.adjacent_layer(A, X), dim=2)                               # layers.py:500
        S_T = torch.transpose(S, 1, 2)
        X_feature = F.relu(torch.bmm(S_T, X_feature))                                 # layers.py:503
        A_update = torch.bmm(torch.bmm(S_T, A), S)                                    # layers.py:504
        A_update = F.dropout(A_update, p=0.3)              # layers.py:505: always on (training defaults to True)
        return A_update, X_feature


class AFM_BatchNorm(nn.Module):
    """layers.py:394-412: BatchNorm1d over the feature axis of [B,N,F], statistics over ALL B*N
    positions.  ``weight`` [1,1] / ``bias`` [1] exist only for state_dict compatibility (the reference
    creates them uninitialised and never uses them; here they start at 1 / 0)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, bias=True):
        super().__init__()
        dev = _param_device()
        self.bn = nn.BatchNorm1d(num_features, eps, momentum, affine).to(dev)
        self.weight = Parameter(torch.ones(1, 1, device=dev))
        if bias:
            self.bias = Parameter(torch.zeros(1, device=dev))
        else:
            self.register_parameter("bias", None)

    def forward(self, x):                                  # stand-alone use only
        return self.bn(x.permute(0, 2, 1).contiguous()).permute(0, 2, 1)


class Ave_multi_view(nn.Module):
    """layers.py:414-437: learned weighted sum over the leading (view) axis."""

    def __init__(self, ave_source_num, feature_size=0, bias=False):
        super().__init__()
        self.ave_source_num, self.feature_size = ave_source_num, feature_size
        self.weight = Parameter(torch.empty(ave_source_num, device=_param_device()))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(0))
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, input):
        return torch.sum(input * self.weight.view(self.ave_source_num, 1, 1, 1), 0)


class GraphConv_block(nn.Module):
    """One bond-relation view (layers.py:52-95): parameter holder with the reference's attribute
    names (``att``, ``graph_conv``, ``batch_norm``, ``self_r``, ``dropout``)."""

    def __init__(self, node_feature_in, bond_feature_num, node_feature_out, dropout):
        super().__init__()
        self.node_feature_in, self.bond_feature_num = node_feature_in, bond_feature_num
        self.node_feature_out, self.dropout = node_feature_out, dropout
        dev = _param_device()
        self.att = nn.Conv2d(bond_feature_num, 1, kernel_size=1, stride=1, padding=0, bias=False).to(dev)
        self.graph_conv = GraphConv_base(node_feature_in, node_feature_out, bias=True)
        self.batch_norm = AFM_BatchNorm(node_feature_out)
        self.self_r = Parameter(torch.empty(1, device=dev))
        self.reset_parameters()

    def reset_parameters(self):                           # layers.py:77-79
        self.self_r.data.uniform_(-0.01, 0.01)

    def _view_params(self):
        bn = self.batch_norm.bn
        return (self.att.weight, self.self_r, self.graph_conv.weight, self.graph_conv.bias, bn.weight, bn.bias)

    def _bn_buffers(self):
        bn = self.batch_norm.bn
        return (bn.running_mean, bn.running_var, bn.num_batches_tracked)

    def forward(self, adjs, afms, bond_relation_tensor, mask_tiny=None, mask2=None, identity=None):
        """Single-view call with the reference signature (layers.py:81).  Routed through the same
        CUDA path as a one-view layer; the mask arguments are recomputed internally and ignored."""
        if self.training:
            raise EagcnError("stand-alone GraphConv_block.forward supports eval mode only; "
                             "use GraphConv_Layer for training")
        plan = GraphPlan.build(adjs, [bond_relation_tensor])
        cfg = EF.LayerConfig(fin=self.node_feature_in, fo=(self.node_feature_out,), training=False,
                             p_drop=float(self.dropout))
        H = EF.gather_rows(plan, afms)
        X = EF.graph_conv_layer(plan, cfg, H, self._view_params(), self._bn_buffers())
        # the reference block returns rows *before* the layer's mask3 (layers.py:313): padded rows there
        # hold relu(BN(bias)).  Reproduce that constant analytically on the dense output.
        x = EF.scatter_rows(plan, X)
        bn = self.batch_norm.bn
        pad = F.relu((self.graph_conv.bias - bn.running_mean) / torch.sqrt(bn.running_var + bn.eps) * bn.weight + bn.bias)
        x = x + (1.0 - plan.row_mask()).unsqueeze(2) * pad
        A1 = EF.attention_dense(plan, [self.att.weight])[0]
        return x, A1


class GraphConv_Layer(nn.Module):
    """All bond relations of one layer (layers.py:262-325), CUDA path."""

    _plan_cache = None          # (key, weakref to adjs, plan): the 4 layers of a model share one plan
    # True (default): a layer whose widths are not multiples of 4 floats (HIV layer 2: 5 x 250, train.py:70-71) runs on
    # widths rounded up to 4 (EF.pad_layer_args: zero-padded parameters, real channels cut out afterwards) and so stays
    # on the tensor-core GEMM and the float4 kernels instead of the FFMA / scalar fallbacks.  The padding algebra is
    # CPU-tested (tests/test_pad_widths.py) and checked on the GPU against the reference-generated golden of that
    # configuration (tests/test_gpu_parity.py); False keeps the un-padded fallback engines.
    pad_widths = True

    def __init__(self, node_feature_in, bond_feature_num, node_out_1, node_out_2, node_out_3, node_out_4,
                 node_out_5, dropout, structure, last=False, adj_size=0):
        super().__init__()
        self.block1 = GraphConv_block(node_feature_in, bond_feature_num, node_out_1, dropout)
        self.block2 = GraphConv_block(node_feature_in, 4, node_out_2, dropout)       # layers.py:270-273
        self.block3 = GraphConv_block(node_feature_in, 2, node_out_3, dropout)
        self.block4 = GraphConv_block(node_feature_in, 2, node_out_4, dropout)
        self.block5 = GraphConv_block(node_feature_in, 2, node_out_5, dropout)
        self.structure, self.last = structure, last
        self.node_feature_in = node_feature_in
        if structure == "Concate":
            self.total_output = node_out_1 + node_out_2 + node_out_3 + node_out_4 + node_out_5
        elif structure == "Weighted_sum":
            self.total_output = node_out_1
            self.ave = Ave_multi_view(5)
        else:
            print("error, structure not support")                                   # layers.py:283
        self.ave_A = Ave_multi_view(5)
        self.self_r = Parameter(torch.empty(1, device=_param_device()))
        self.materialize_A = True
        self.stat_allreduce = None      # set by eagcn_b200.parallel for global-batch BatchNorm
        self.rng_stream = 0
        self.reset_parameters()

    def reset_parameters(self):                           # layers.py:290-291
        self.self_r.data.uniform_(-0.01, 0.01)

    @property
    def blocks(self):
        return (self.block1, self.block2, self.block3, self.block4, self.block5)

    # ------------------------------------------------------------------------------------
    @classmethod
    def _plan_for(cls, adjs, rels):
        key = (adjs.data_ptr(), adjs._version, tuple(adjs.shape)) + tuple((r.data_ptr(), r._version) for r in rels)
        c = cls._plan_cache
        if c is not None and c[0] == key and c[1]() is adjs:
            return c[2]
        plan = GraphPlan.build(adjs, rels)
        plan.check()                     # eager drop-in path: validate the batch once (one small D2H)
        cls._plan_cache = (key, weakref.ref(adjs), plan)
        return plan

    def _weighted_sum(self, plan, cfg, H, params, buffers, p_drop):
        """layers.py:314-316 + 431-437: x = sum_v ave.weight[v] * x_v, WITHOUT the row mask of the 'Concate' branch.
        Rows with bonds come from the kernels; every other row (padding, bond-less atoms) holds the same value per
        view -- dropout(relu(BatchNorm(bias))) -- which the reference leaves in its output, so it is reproduced here
        (value from the statistics the kernels used; its gradient re-enters the BatchNorm backward sums)."""
        V, fo = 5, self.total_output
        X, z_pad = EF.graph_conv_layer(plan, cfg, H, params, buffers)
        w = self.ave.weight
        x = EF.scatter_rows(plan, (X.view(plan.t_cap, V, fo) * w.view(1, V, 1)).sum(1))
        pad_v = torch.relu(z_pad).view(V, fo)
        inactive = 1.0 - plan.row_mask()                                               # [B,N]
        if self.training and p_drop > 0.0:
            # the reference draws an independent keep mask for every padded element too (F.dropout on the dense tensor)
            keep = (torch.rand(plan.B, plan.N, V, fo, device=X.device) >= p_drop).to(X.dtype) / (1.0 - p_drop)
            pad = (keep * (pad_v * w.view(V, 1)).view(1, 1, V, fo)).sum(2)
            return x + inactive.unsqueeze(2) * pad
        return x + inactive.unsqueeze(2) * (pad_v * w.view(V, 1)).sum(0).view(1, 1, fo)

    def _flat_params(self):
        params, buffers = [], []
        for b in self.blocks:
            params += b._view_params()
            buffers += b._bn_buffers()
        return params, buffers

    def prepare(self, stream=None):
        """Per-step parameter preparation of this layer ahead of its forward call (EF.LayerPrep): pass the result as
        ``forward(..., prep=...)``.  With ``stream`` (a side stream) it overlaps whatever runs on the current stream --
        the packing of the batch, typically."""
        params, _ = self._flat_params()
        chans = tuple(b.bond_feature_num for b in self.blocks)
        return EF.LayerPrep(self.node_feature_in, tuple(b.node_feature_out for b in self.blocks), chans,
                            [p.detach() for p in params], params[0].device, stream)

    def forward(self, adjs, afms, TypeAtt=None, OrderAtt=None, AromAtt=None, ConjAtt=None, RingAtt=None, prep=None):
        if self.structure not in ("Concate", "Weighted_sum"):
            raise EagcnError(f"structure {self.structure!r} is not supported (layers.py:279-283)")
        wsum = self.structure == "Weighted_sum"
        if wsum and (isinstance(afms, PackedRows) or any(b.node_feature_out != self.total_output for b in self.blocks)):
            raise EagcnError("structure='Weighted_sum' (layers.py:314-316) takes dense [B,N,F] features and five views "
                             "of equal width: its un-masked padded rows have no packed-row representation")
        if isinstance(adjs, GraphPlan):
            plan = adjs
        else:
            if not torch.is_tensor(adjs) or not adjs.is_cuda:
                raise EagcnError("eagcn_b200.GraphConv_Layer is CUDA-only (sm_100a, no CPU fallback)")
            plan = self._plan_for(adjs, (TypeAtt, OrderAtt, AromAtt, ConjAtt, RingAtt))
        if plan.V != 5:
            raise ValueError("GraphConv_Layer has 5 views (layers.py:269-273); plan has %d" % plan.V)
        packed_io = isinstance(afms, PackedRows)
        H = afms.rows if packed_io else EF.gather_rows(plan, afms)
        p_drop = float(self.block1.dropout)
        bn1 = self.block1.batch_norm.bn                    # AFM_BatchNorm(eps=1e-5, momentum=0.1), layers.py:399-401
        if any(b.batch_norm.bn.eps != bn1.eps or b.batch_norm.bn.momentum != bn1.momentum for b in self.blocks):
            raise EagcnError("the five views of a GraphConv_Layer must share BatchNorm eps / momentum")
        cfg = EF.LayerConfig(fin=self.node_feature_in, fo=tuple(b.node_feature_out for b in self.blocks),
                             training=self.training, p_drop=p_drop, rng_stream=self.rng_stream,
                             eps=float(bn1.eps), momentum=float(bn1.momentum if bn1.momentum is not None else 0.1),
                             stat_allreduce=self.stat_allreduce, want_pad=wsum, prep=prep)
        if wsum and self.stat_allreduce is not None:
            raise EagcnError("structure='Weighted_sum' with global-batch BatchNorm is not implemented")
        params, buffers = self._flat_params()
        if wsum:
            x = self._weighted_sum(plan, cfg, H, params, buffers, p_drop)
        else:
            if self.pad_widths and (cfg.fin % 4 or any(f % 4 for f in cfg.fo)):
                cfg_p, H_p, params_p, buffers_p, finish = EF.pad_layer_args(cfg, H, params, buffers)
                X = finish(EF.graph_conv_layer(plan, cfg_p, H_p, params_p, buffers_p))
            else:
                X = EF.graph_conv_layer(plan, cfg, H, params, buffers)
            x = PackedRows(X, plan) if packed_io else EF.scatter_rows(plan, X)        # layers.py:313

        A_weight = None
        if self.materialize_A:
            A_weight = EF.attention_dense(plan, [b.att.weight for b in self.blocks])   # layers.py:318
            if self.last:                                                             # layers.py:319-324
                adj = plan._src[0] if len(plan._src) == 2 else None
                if adj is None:
                    raise EagcnError("last=True needs the dense adjacency (plan built from codes)")
                m = plan.row_mask().unsqueeze(2)
                ident = m * torch.eye(plan.N, device=adj.device, dtype=adj.dtype)
                Aw = self.ave_A(A_weight)
                Aw = torch.sigmoid(Aw) * adj + torch.sigmoid(self.self_r) * ident + (1.0 - adj) * 1e-9
                A_weight = Aw / Aw.sum(dim=2, keepdim=True) * m
        return x, A_weight
