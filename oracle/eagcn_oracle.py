"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the EAGCN multi-view edge-attention layer stack.

A plain torch-CPU restatement (differentiable, so ``torch.autograd`` on it is the oracle for the
backward pass too) of the reference hot path, written in the *lookup* form: on one-hot relation
tensors the reference's 1x1 Conv2d attention score (layers.py:64,82) is exactly the table lookup
``a_v[type_v(i,j)]``.  Each function cites the reference lines it follows.

Pinned against (tests/test_oracle.py):
  * the unmodified reference classes, imported from /root/reference in the build container
    (oracle/ref_loader.py), and
  * golden vectors generated FROM those reference classes (tests/golden/make_golden.py ->
    tests/golden/*.npz), which travel to the GPU box.
The reference itself ships no tests / golden vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product path (eagcn_b200/) never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPS_BN = 1e-5       # layers.py:399
MOM_BN = 0.1        # layers.py:399
TINY = 1e-9         # layers.py:294


# --------------------------------------------------------------------------------------------
# indexing (bit-exact part)
# --------------------------------------------------------------------------------------------
def codes_from_onehot(adj: torch.Tensor, rel: torch.Tensor) -> torch.Tensor:
    """Edge-type codes of one view: int64 [B,N,N].

    rel [B,C,N,N] (neural_fp.py:111-120) -> code c in [0,C) where rel[b,c,i,j]==1 on bonded
    pairs; C where the channel vector is all-zero; -1 off-graph (adj==0).  Raises on anything
    that is not one-hot 0/1 on a bonded pair (the lookup form is exact only for those).
    """
    B, C, N, _ = rel.shape
    on = adj > 0
    cnt = (rel != 0).sum(1)
    ok = ((rel == 0) | (rel == 1)).all(1)
    if not bool((ok | ~on).all()) or not bool(((cnt <= 1) | ~on).all()):
        raise ValueError("relation tensor is not one-hot on bonded pairs")
    code = rel.argmax(1)
    code = torch.where(cnt == 0, torch.full_like(code, C), code)
    return torch.where(on, code, torch.full_like(code, -1))


def row_mask(adj: torch.Tensor) -> torch.Tensor:
    """m[b,i] = max_j adj[b,i,j]  (layers.py:295)."""
    return adj.max(dim=2).values


# --------------------------------------------------------------------------------------------
# one view (GraphConv_block.forward, layers.py:81-95)
# --------------------------------------------------------------------------------------------
def attention(adj, code, a, self_r, m):
    """A1 = sigmoid(a[code])*adj (layers.py:82-83); A_hat = row-normalised (layers.py:84-90)."""
    B, N, _ = adj.shape
    a_ext = torch.cat([a.reshape(-1), a.new_zeros(1)])
    s = a_ext[code.clamp(min=0)]
    A1 = torch.sigmoid(s) * adj                                          # layers.py:83
    eye = torch.eye(N, dtype=adj.dtype)
    ident = m.unsqueeze(2) * eye                                          # layers.py:299-304
    tiny = (1.0 - adj) * TINY                                             # layers.py:294
    U = A1 + torch.sigmoid(self_r) * ident + tiny                         # layers.py:84
    A = (U / U.sum(dim=2, keepdim=True)) * m.unsqueeze(2)                 # layers.py:87-90
    return A1, A


def batch_norm_all_positions(Y, gamma, beta, running_mean, running_var, training):
    """AFM_BatchNorm.forward (layers.py:408-412): BatchNorm1d over ALL B*N positions, padded
    rows included.  Returns (out, new_running_mean, new_running_var)."""
    B, N, C = Y.shape
    if training:
        flat = Y.reshape(B * N, C)
        mean = flat.mean(0)
        var = flat.var(0, unbiased=False)
        M = B * N
        with torch.no_grad():
            new_rm = (1 - MOM_BN) * running_mean + MOM_BN * mean
            new_rv = (1 - MOM_BN) * running_var + MOM_BN * var * (M / max(M - 1, 1))
    else:
        mean, var = running_mean, running_var
        new_rm, new_rv = running_mean, running_var
    out = (Y - mean) / torch.sqrt(var + EPS_BN) * gamma + beta
    return out, new_rm, new_rv


def block_forward(sd, pre, adj, afm, code, m, training, p=0.0, keep=None, relu_mask=None):
    """One view.  ``sd`` holds the reference state_dict keys under prefix ``pre``.

    relu_mask: optional 0/1 tensor that replaces the ReLU decision (X = Z * mask).  Gradient parity at
    realistic sizes needs it: with ~10^6 pre-activations a handful lie within 1e-6 of zero, where any two
    fp32 implementations may disagree on relu'(z) -- a measure-zero kink, not an error.  The tests take
    the mask from the implementation under test and separately assert that it differs from (Z > 0) only
    where |Z| is tiny.

    keep: optional [B,N,Fo] 0/1 dropout keep-mask (reference draws it from torch's global RNG,
    layers.py:94 -- not reproducible across implementations, so it is an input here).
    Returns (X_v, A1_v, Y_v, (new_running_mean, new_running_var)).
    """
    a = sd[pre + "att.weight"]
    W = sd[pre + "graph_conv.weight"]
    b = sd[pre + "graph_conv.bias"]
    A1, A = attention(adj, code, a, sd[pre + "self_r"], m)
    support = torch.bmm(A, afm)                                           # layers.py:39
    Y = support.reshape(-1, W.shape[0]).mm(W).reshape(adj.shape[0], adj.shape[1], -1) + b  # :40-43
    Z, rm, rv = batch_norm_all_positions(Y, sd[pre + "batch_norm.bn.weight"], sd[pre + "batch_norm.bn.bias"],
                                         sd[pre + "batch_norm.bn.running_mean"],
                                         sd[pre + "batch_norm.bn.running_var"], training)
    X = F.relu(Z) if relu_mask is None else Z * relu_mask                 # layers.py:93
    if training and p > 0.0:                                              # layers.py:94
        if keep is None:
            raise ValueError("training with dropout>0 needs an explicit keep mask")
        X = X * keep / (1.0 - p)
    return X, A1, Y, (rm, rv), Z


# --------------------------------------------------------------------------------------------
# one layer (GraphConv_Layer.forward, layers.py:293-325)
# --------------------------------------------------------------------------------------------
def layer_forward(sd, pre, adj, afm, codes, training, p=0.0, keeps=None, structure="Concate",
                  last=False, n_views=5, relu_masks=None):
    """codes: list of V int64 [B,N,N] tensors.  Returns dict(x, A_weight, Y=[..], stats=[..], Z=[..])."""
    m = row_mask(adj)
    xs, A1s, Ys, stats, Zs = [], [], [], [], []
    for v in range(n_views):
        X, A1, Y, st, Z = block_forward(sd, f"{pre}block{v + 1}.", adj, afm, codes[v], m, training, p,
                                        None if keeps is None else keeps[v],
                                        None if relu_masks is None else relu_masks[v])
        xs.append(X); A1s.append(A1); Ys.append(Y); stats.append(st); Zs.append(Z)
    if structure == "Concate":
        x = torch.cat(xs, dim=2) * m.unsqueeze(2)                         # layers.py:313
    elif structure == "Weighted_sum":
        w = sd[pre + "ave.weight"]
        x = (torch.stack(xs, 0) * w.view(-1, 1, 1, 1)).sum(0)             # layers.py:315-316,431-437
    else:
        raise ValueError(structure)
    A_weight = torch.stack(A1s, 0)                                        # layers.py:318
    if last:                                                              # layers.py:319-324
        N = adj.shape[1]
        Aw = (A_weight * sd[pre + "ave_A.weight"].view(-1, 1, 1, 1)).sum(0)
        ident = m.unsqueeze(2) * torch.eye(N, dtype=adj.dtype)
        Aw = torch.sigmoid(Aw) * adj + torch.sigmoid(sd[pre + "self_r"]) * ident + (1.0 - adj) * TINY
        A_weight = Aw / Aw.sum(2, keepdim=True) * m.unsqueeze(2)
    return dict(x=x, A_weight=A_weight, Y=Ys, stats=stats, Z=Zs)


def stack_forward(sd, adj, afm, codes, n_layers, training, p=0.0, keeps=None, structure="Concate",
                  last_flags=None, relu_masks=None):
    """EAGCN.forward chaining (models.py:96-100) for ``n_layers`` GraphConv_Layers ``layer1..``.
    keeps / relu_masks: per layer, per view (see block_forward)."""
    h = afm
    outs = []
    for l in range(n_layers):
        last = bool(last_flags[l]) if last_flags is not None else False
        o = layer_forward(sd, f"layer{l + 1}.", adj, h, codes, training, p,
                          None if keeps is None else keeps[l], structure, last,
                          relu_masks=None if relu_masks is None else relu_masks[l])
        outs.append(o)
        h = o["x"]
    return h, outs


def diff_pool(sd, pre, A, X):
    """Diff_Pooling.forward (the class actually in effect, layers.py:492-506; the first definition at :476-490 is
    inside a string literal).  GraphConv_base without bias (layers.py:38-45): bmm(A, X) then mm with the weight.
    The returned A_update goes through an always-on F.dropout(p=0.3) in the reference but is never used by
    models.py:104-106, so only X_feature is restated."""
    ax = torch.bmm(A, X)                                                  # layers.py:39
    feat = F.relu(ax.matmul(sd[pre + "feature_layer.weight"]))            # layers.py:499
    S = F.softmax(ax.matmul(sd[pre + "adjacent_layer.weight"]), dim=2)    # layers.py:500
    return F.relu(torch.bmm(S.transpose(1, 2), feat))                     # layers.py:501-503


def head_forward(sd, x_atoms, sizes, training, p=0.0, keep=None, molfp_mode="sum", x0=None, relu_masks=None,
                 A_last=None):
    """Read-out + dense head (models.py:104-121), molfp_mode in {'sum','ave','pool'}.  x0: start from a given
    read-out [B,F] instead of x_atoms.  relu_masks: optional (mask1 [B,D1], mask2 [B,D2]) replacing the two ReLU
    decisions (see block_forward).  A_last: the last layer's normalised averaged attention [B,N,N] ('pool')."""
    if x0 is None:
        if molfp_mode == "pool":
            x = diff_pool(sd, "pool1.", A_last, x_atoms).sum(1)           # models.py:104-106
        else:
            x = x_atoms.sum(1)                                            # models.py:108
            if molfp_mode == "ave":
                x = x / sizes.view(-1, 1).to(x.dtype)                     # models.py:110-111
    else:
        x = x0

    def bn(x, pre):
        if training:
            mean = x.mean(0); var = x.var(0, unbiased=False)
        else:
            mean = sd[pre + "running_mean"]; var = sd[pre + "running_var"]
        return (x - mean) / torch.sqrt(var + EPS_BN) * sd[pre + "weight"] + sd[pre + "bias"]

    x = bn(x, "Graph_BN.")                                                # models.py:112
    x = x.mm(sd["den1.weight"])                                           # models.py:114
    z1 = bn(x, "bn_den1.")
    x = F.relu(z1) if relu_masks is None else z1 * relu_masks[0]          # models.py:115
    if training and p > 0.0:                                              # models.py:116
        x = x * keep / (1.0 - p)
    x = x.mm(sd["den2.weight"])                                           # models.py:117
    g = x
    z2 = bn(x, "bn_den2.")
    x = F.relu(z2) if relu_masks is None else z2 * relu_masks[1]          # models.py:119
    x = x.mm(sd["den3.weight"])                                           # models.py:120
    return x, g


# --------------------------------------------------------------------------------------------
# helpers for tests
# --------------------------------------------------------------------------------------------
def clone_sd(module_or_sd, dtype=None, requires_grad=False):
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().cpu()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            t.requires_grad_(requires_grad)
        out[k] = t
    return out


def rel_err(x: torch.Tensor, ref: torch.Tensor) -> float:
    """max|x-ref| / max|ref| -- the parity metric (SURVEY.md 7 'hard parts': never element-wise
    relative on near-zero post-ReLU values)."""
    denom = float(ref.abs().max())
    if denom == 0.0:
        return float((x - ref).abs().max())
    return float((x.double() - ref.double()).abs().max()) / denom


# --------------------------------------------------------------------------------------------
# reference-cost form (cpu_baseline of bench.py): the SAME op sequence the reference executes --
# 1x1 Conv2d over the dense one-hot planes, expanded masks, bmm + mm, BatchNorm1d on the permuted
# tensor, F.dropout from torch's generator -- so its timing stands in for the reference's CPU path
# on machines where /root/reference does not exist (the GPU box).
# --------------------------------------------------------------------------------------------
def block_forward_conv(sd, pre, adj, afm, rel, mask_tiny, mask2, identity, training, p):
    a = sd[pre + "att.weight"]
    B, N = adj.shape[0], adj.shape[1]
    A1 = F.conv2d(rel.float(), a).view(B, N, -1)                                      # layers.py:82
    A1 = torch.sigmoid(A1) * adj                                                      # layers.py:83
    A = A1 + torch.sigmoid(sd[pre + "self_r"]) * identity + mask_tiny                 # layers.py:84
    A_rowsum = torch.sum(A, dim=2, keepdim=True).expand(B, N, N)                      # layers.py:87
    A = (A / A_rowsum) * mask2                                                        # layers.py:90
    W = sd[pre + "graph_conv.weight"]
    support = torch.bmm(A, afm)                                                       # layers.py:39
    x = torch.mm(support.view(-1, W.shape[0]), W).view(-1, N, W.shape[1]) + sd[pre + "graph_conv.bias"]
    x = x.permute(0, 2, 1)                                                            # layers.py:409-411
    x = F.batch_norm(x.contiguous(), sd[pre + "batch_norm.bn.running_mean"], sd[pre + "batch_norm.bn.running_var"],
                     sd[pre + "batch_norm.bn.weight"], sd[pre + "batch_norm.bn.bias"], training, MOM_BN, EPS_BN)
    x = F.relu(x.permute(0, 2, 1))                                                    # layers.py:93
    x = F.dropout(x, p=p, training=training)                                          # layers.py:94
    return x, A1


def layer_forward_conv(sd, pre, adj, afm, rels, training, p=0.0):
    """GraphConv_Layer.forward in the reference's own op sequence (layers.py:293-325, 'Concate')."""
    B, N = adj.shape[0], adj.shape[1]
    mask_tiny = (1.0 - adj) * TINY                                                    # layers.py:294
    mask_blank, _ = adj.max(dim=2, keepdim=True)                                      # layers.py:295
    mask2 = mask_blank.expand(B, N, N)
    identity = mask_blank.expand(B, N, N) * torch.eye(N)                              # layers.py:302-304
    xs, As = [], []
    for v, rel in enumerate(rels):
        x, A1 = block_forward_conv(sd, f"{pre}block{v + 1}.", adj, afm, rel, mask_tiny, mask2, identity, training, p)
        xs.append(x); As.append(A1)
    x = torch.cat(xs, dim=2)
    x = x * mask_blank.expand(B, N, x.shape[2])                                       # layers.py:313
    return x, torch.stack(As, 0)                                                      # layers.py:318


def model_forward_conv(sd, adj, afm, rels, sizes, n_layers, training, p=0.0):
    """n_layers x GraphConv_Layer + sum read-out + head (models.py:96-121) with the reference's op sequence."""
    h = afm
    for l in range(n_layers):
        h, _ = layer_forward_conv(sd, f"layer{l + 1}.", adj, h, rels, training, p)
    _ = h.data.cpu()                                                                  # models.py:102
    x = torch.sum(h, 1)                                                               # models.py:108

    def bn(x, pre):
        return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"],
                            sd[pre + "bias"], training, MOM_BN, EPS_BN)
    x = bn(x, "Graph_BN.")
    x = F.relu(bn(x.mm(sd["den1.weight"]), "bn_den1."))
    x = F.dropout(x, p=p, training=training)
    x = x.mm(sd["den2.weight"])
    x = F.relu(bn(x, "bn_den2."))
    return x.mm(sd["den3.weight"])


# --------------------------------------------------------------------------------------------
# loss of the classification branch, restated loop for loop (utils.py:653-679, train.py:326-331)
# --------------------------------------------------------------------------------------------
def weight_tensor_loop(weights, labels):
    """utils.weight_tensor: Python double loop over (molecule, task)."""
    out = []
    for i in range(labels.shape[0]):
        for j in range(labels.shape[1]):
            try:
                v = int(labels[i][j])
                if v == 1:
                    out.append(weights[j][0])
                elif v == 0:
                    out.append(weights[j][1])
                else:
                    out.append(0)
            except ValueError:                      # int(nan)
                out.append(0)
    return torch.tensor(out, dtype=torch.float32)


def bce_loss_loop(outputs, labels, weights):
    w = weight_tensor_loop(weights, labels)
    non_nan = ((labels == 1).sum() + (labels == 0).sum()).float()
    target = torch.nan_to_num(labels.float(), nan=0.0)
    return F.binary_cross_entropy_with_logits(outputs.view(-1), target.view(-1), weight=w, reduction="sum") / non_nan
