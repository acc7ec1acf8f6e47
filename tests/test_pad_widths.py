"""CPU: the width-padding algebra of ``functional.pad_layer_args`` (layers whose widths are not multiples of 4 floats,
e.g. HIV layer 2: five views of 250 -- SURVEY.md 8(a) a7).  The CUDA core is replaced by a small torch-CPU core with the
same interface and the same channel-wise structure (projection, per-channel training-mode BatchNorm over ALL positions
with running-statistic updates, ReLU): padding must leave the real channels, every gradient and the running statistics
unchanged, and padded channels must stay finite."""
import torch

from eagcn_b200 import functional as EF


def _toy_core(cfg, H, params, buffers, m_total):
    """X = relu(BN_v(A (H W_v) + b_v)) per view on packed rows; statistics over m_total positions of which the rows
    of H are the non-constant ones (the padded rows of the reference hold Y = b exactly: layers.py:408-412)."""
    T = H.shape[0]
    A = torch.tril(torch.ones(T, T)) / torch.arange(1, T + 1).view(-1, 1)      # some fixed row-stochastic mixing
    outs = []
    for v, fo in enumerate(cfg.fo):
        a, r, W, b, g, be = params[6 * v: 6 * v + 6]
        rm, rv, nbt = buffers[3 * v: 3 * v + 3]
        Y = A @ (H @ W) + b
        if cfg.training:
            D = Y - b
            m1 = D.sum(0) / m_total
            var = (D * D).sum(0) / m_total - m1 * m1
            mean = b + m1
            with torch.no_grad():
                rm.mul_(1 - cfg.momentum).add_(cfg.momentum * mean)
                rv.mul_(1 - cfg.momentum).add_(cfg.momentum * var * m_total / (m_total - 1))
        else:
            mean, var = rm, rv
        outs.append(torch.relu((Y - mean) / torch.sqrt(var + cfg.eps) * g + be))
    return torch.cat(outs, 1)


def _make(fin, fo, seed):
    g = torch.Generator().manual_seed(seed)
    params, buffers = [], []
    for f in fo:
        params += [torch.randn(3, generator=g), torch.randn(1, generator=g),
                   (torch.randn(fin, f, generator=g) * 0.3).requires_grad_(True),
                   (torch.randn(f, generator=g) * 0.1).requires_grad_(True),
                   (torch.rand(f, generator=g) + 0.5).requires_grad_(True),
                   (torch.randn(f, generator=g) * 0.1).requires_grad_(True)]
        buffers += [torch.randn(f, generator=g) * 0.1, torch.rand(f, generator=g) + 0.5, torch.zeros((), dtype=torch.long)]
    return params, buffers


def _run(training, padded):
    fin, fo, T, m_total = 10, (250 % 7 + 3, 5, 6, 9, 2), 37, 64          # widths 8, 5, 6, 9, 2; fin 10 -> 12
    params, buffers = _make(fin, fo, 1)
    H = torch.randn(T, fin, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    cfg = EF.LayerConfig(fin=fin, fo=fo, training=training)
    if padded:
        cfg_p, H_p, params_p, buffers_p, finish = EF.pad_layer_args(cfg, H, params, buffers)
        assert cfg_p.fin % 4 == 0 and all(f % 4 == 0 for f in cfg_p.fo)
        X_p = _toy_core(cfg_p, H_p, params_p, buffers_p, m_total)
        assert torch.isfinite(X_p).all()
        X = finish(X_p)
    else:
        X = _toy_core(cfg, H, params, buffers, m_total)
    R = torch.randn(X.shape, generator=torch.Generator().manual_seed(3))
    (X * R).sum().backward()
    grads = [p.grad for p in params if p.requires_grad] + [H.grad]
    return X.detach(), grads, [b for b in buffers if b.is_floating_point()]


def test_padding_preserves_values_gradients_and_running_stats():
    for training in (True, False):
        X0, g0, b0 = _run(training, padded=False)
        X1, g1, b1 = _run(training, padded=True)
        assert X0.shape == X1.shape
        assert torch.allclose(X0, X1, rtol=1e-6, atol=1e-6)
        scale = max(float(a.abs().max()) for a in g0)        # d bias through a training-mode BatchNorm is rounding noise
        for a, b in zip(g0, g1):
            assert a.shape == b.shape and float((a - b).abs().max()) <= 2e-5 * scale
        for a, b in zip(b0, b1):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)


def test_no_padding_needed_is_identity():
    params, buffers = _make(8, (4, 8), 5)
    H = torch.randn(9, 8)
    cfg = EF.LayerConfig(fin=8, fo=(4, 8), training=True)
    cfg_p, H_p, params_p, buffers_p, finish = EF.pad_layer_args(cfg, H, params, buffers)
    assert cfg_p.fin == 8 and cfg_p.fo == (4, 8) and H_p is H
    for a, b in zip(params, params_p):
        assert a.shape == b.shape
