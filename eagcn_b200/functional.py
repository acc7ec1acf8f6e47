"""torch.autograd bindings of the CUDA layer path (the only callers of the C ABI's compute entry points).

``graph_conv_layer`` is the packed-row core of GraphConv_Layer.forward (reference layers.py:293-325):
H [t_cap, fin] -> X [t_cap, sum_v fo_v].  Everything runs on torch's current CUDA stream, allocates
through torch's caching allocator and never synchronises, so a whole training step can be captured in
one CUDA graph.
"""
from __future__ import annotations

import ctypes
import dataclasses
import os
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import EagcnError, LayerStruct, WorkStruct, check, lib, ptr
from .plan import GraphPlan, _stream

_F32 = torch.float32


class Overlap:
    """Independent branches of a step on a side stream (captured into a CUDA graph as parallel branches).

    A training step is a chain of ~40 small kernels; three groups of them do not depend on their stream neighbours:
      * the per-step parameter preparation (W_all concat / hi-lo split / sigmoid tables) and the dropout-generator
        fork depend only on the parameters -- they run beside the graph-plan packing (``prefetch``);
      * dW_all = H^T Q (+ the attention-gradient sums) and dH = Q W_all^T both depend only on Q -- dW runs beside dH and
        beside the next layer's backward;
      * dW = x^T dy and dx = dy W^T of a Dense layer.
    Fork = the side stream waits for the current stream; join = the current stream waits for the side stream.  The
    join of the backward branches is deferred to the end of the backward pass (autograd engine callback) when the
    parameters receive fresh gradients (``p.grad is None``: autograd adopts the returned tensor without touching it);
    otherwise it happens before the layer's backward returns.  Every tensor a side-stream kernel touches is kept
    alive until the join, so the caching allocator cannot hand its memory to a main-stream kernel early.
    ``Overlap.enabled = False`` (or EAGCN_OVERLAP=0) puts everything back on one stream."""
    enabled = os.environ.get("EAGCN_OVERLAP", "1") != "0"
    defer_param_grads = True
    _streams = {}
    _pending = {}          # device index -> list of keep-alive tuples of the un-joined backward branches
    _cb_armed = {}

    @classmethod
    def side(cls, dev):
        key = torch.device(dev).index or 0
        if key not in cls._streams:
            cls._streams[key] = torch.cuda.Stream(device=dev)
        return cls._streams[key]

    @classmethod
    def fork(cls, dev):
        """The side stream, ordered after everything enqueued on the current stream so far."""
        s = cls.side(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        return s

    @classmethod
    def fork_point(cls, dev):
        """Mark a branch point on the current stream; work enqueued on it afterwards is NOT waited for by fork_from."""
        e = torch.cuda.Event()
        e.record(torch.cuda.current_stream(dev))
        return e

    @classmethod
    def fork_from(cls, dev, event):
        s = cls.side(dev)
        s.wait_event(event)
        return s

    @classmethod
    def join(cls, dev):
        key = torch.device(dev).index or 0
        torch.cuda.current_stream(dev).wait_stream(cls.side(dev))
        cls._pending.pop(key, None)
        cls._cb_armed[key] = False

    _fwd_open = {}

    @classmethod
    def fork_fwd(cls, dev):
        """Side stream for the forward prefetch (parameter preparation + dropout-generator fork)."""
        cls.join_pending(dev)
        cls._fwd_open[torch.device(dev).index or 0] = True
        return cls.fork(dev)

    @classmethod
    def join_fwd(cls, dev):
        """First consumer of prefetched data: the current stream waits for the side stream (once per prefetch)."""
        key = torch.device(dev).index or 0
        if cls._fwd_open.get(key):
            cls._fwd_open[key] = False
            torch.cuda.current_stream(dev).wait_stream(cls.side(dev))

    @classmethod
    def join_pending(cls, dev):
        """Join now if a backward branch is still open (start of a new step after an aborted backward)."""
        if cls._pending.get(torch.device(dev).index or 0):
            cls.join(dev)

    @classmethod
    def defer_join(cls, dev, keep):
        """Join at the end of the running backward pass; ``keep`` stays referenced until then."""
        key = torch.device(dev).index or 0
        cls._pending.setdefault(key, []).append(keep)
        if not cls._cb_armed.get(key):
            cls._cb_armed[key] = True
            torch.autograd.Variable._execution_engine.queue_callback(lambda: cls.join(dev))

    # ---- how many autograd functions of the recorded graph(s) produce a gradient for a parameter -------------------
    # A parameter with TWO producers (models.py:104-106 / layers.py:318: the attention weights feed the layer AND the
    # returned dense attention of the 'pool' read-out; user-side weight tying) gets its gradients SUMMED by the autograd
    # engine's input buffer the moment the second one arrives -- a main-stream kernel reading the first one, which a
    # deferred join may not have written yet.  Every function of this library that hands out parameter gradients
    # registers its parameters at forward time; a parameter seen more than once since the last backward pass makes
    # ``fresh`` false (the join then precedes the function's return).
    _uses = {}             # parameter storage pointer -> number of gradient producers recorded since the last backward
    _bwd_seen = False

    @classmethod
    def note_use(cls, params):
        """Forward side: these parameters get a gradient producer (call only when the function will be differentiated)."""
        if cls._bwd_seen:                     # first forward after a backward pass: a new set of graphs starts
            cls._uses.clear()
            cls._bwd_seen = False
        for p in params:
            if p is not None and p.requires_grad:
                k = p.data_ptr()
                cls._uses[k] = cls._uses.get(k, 0) + 1

    @classmethod
    def note_backward(cls):
        cls._bwd_seen = True

    @staticmethod
    def fresh(params):
        """True when autograd will adopt the returned gradient tensors as-is, i.e. when NOTHING launches a kernel on
        them before the end of the backward pass: every parameter is a leaf without a gradient yet (AccumulateGrad
        steals the tensor), has exactly ONE gradient producer in the recorded graphs (``note_use``; two producers are
        summed by the engine on arrival), carries no tensor hooks and no post-accumulate-grad hooks
        (optimizer-in-backward, FSDP), and the backward pass is not itself being recorded (``create_graph=True`` keeps
        grad mode on and makes AccumulateGrad clone).  Hooks on the AccumulateGrad NODES (torch DDP's bucket hooks)
        cannot be seen from here: wrap the model in DDP only with ``Overlap.defer_param_grads = False`` -- the join then
        precedes the layer's return (eagcn_b200.parallel's own FlatGradBucket reduces after backward() and needs
        nothing)."""
        if not Overlap.defer_param_grads or torch.is_grad_enabled():
            return False
        for p in params:
            if p is None or not p.requires_grad:
                continue
            if not p.is_leaf or p.grad is not None or p._backward_hooks:
                return False
            if getattr(p, "_post_accumulate_grad_hooks", None):
                return False
            if Overlap._uses.get(p.data_ptr(), 0) != 1:      # unregistered (0) or shared (> 1): do not defer
                return False
        return True


class GradArena:
    """All parameter gradients of a backward pass in ONE flat fp32 buffer, written there by the kernels themselves.

    The layer / dense / BatchNorm backward functions allocate their gradient outputs (dW_all, d bias / gamma / beta,
    d att / d self_r, head dW, dgamma, dbeta) through ``GradArena.empty``; with an arena installed (``GradArena.current``)
    those are consecutive 64-byte aligned slices of ``flat`` in backward order (head first, layer 1 last) and autograd
    hands the parameters views of it -- so the data-parallel gradient exchange is ONE collective over ``flat[:off]``
    with no packing copy, and ``flat[:mark]`` -- everything above layer 1 -- can be reduced while
    layer 1's backward still runs (eagcn_b200.parallel.ArenaAllReduce).  Offsets are a pure function of the model, so a
    captured CUDA graph replays onto the same addresses; call ``reset()`` at the start of every step."""
    current = None

    def __init__(self, nfloats, device):
        self.flat = torch.zeros(int(nfloats), dtype=_F32, device=device)
        self.off = 0
        self.device = torch.device(device)

    def reset(self):
        self.off = 0

    def take(self, n):
        n = int(n)
        if self.off + n > self.flat.numel():
            raise EagcnError(f"GradArena of {self.flat.numel()} floats is too small (need {self.off + n}); size it with GradArena.measure")
        t = self.flat[self.off:self.off + n]
        self.off += (n + 15) & ~15
        return t

    @staticmethod
    def empty(n, device):
        a = GradArena.current
        if a is not None and a.device == torch.device(device):
            return a.take(n)
        return torch.empty(int(n), dtype=_F32, device=device)

    @classmethod
    def measure(cls, run_backward, device):
        """Floats one backward pass allocates (run under a counting arena)."""
        class _Count:
            def __init__(self):
                self.off, self.device = 0, torch.device(device)

            def take(self, n):
                self.off += (int(n) + 15) & ~15
                return torch.empty(int(n), dtype=_F32, device=device)
        prev, cls.current = cls.current, _Count()
        try:
            run_backward()
            return cls.current.off
        finally:
            cls.current = prev

    def holds(self, t):
        """True when tensor ``t`` lives inside the arena (its gradient needs no separate exchange)."""
        lo = self.flat.data_ptr()
        return t is not None and lo <= t.data_ptr() < lo + self.flat.numel() * 4


@dataclass
class LayerConfig:
    """Static description of one GraphConv_Layer call."""
    fin: int
    fo: tuple
    training: bool
    p_drop: float = 0.0
    eps: float = 1e-5            # layers.py:399
    momentum: float = 0.1        # layers.py:399
    rng_stream: int = 0
    # BatchNorm statistics across data-parallel ranks: callable(tensor) -> None that
    # all-reduces (sum) an fp64 [2, fo_tot] tensor in place, or None for per-replica statistics
    stat_allreduce: object = None
    # also return z_pad [fo_tot]: the post-BatchNorm pre-activation every bond-less / padded row holds (Y = b there).
    # 'Weighted_sum' does not mask those rows (layers.py:314-316), so their value and its gradient matter.
    want_pad: bool = False
    # LayerPrep from prepare_layer(): wall / wallT / wsplit / ball / sig already filled (possibly on the side stream)
    prep: object = None


class RngState:
    """Device-resident philox (seed, offset); advancing it is a captured device op (graph-replay safe)."""
    _states = {}

    def __init__(self, device, seed=0):
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self._inc = torch.tensor([0, 1 << 20], dtype=torch.int64, device=device)
        self._queue = []
        self._queue_ready = None

    @classmethod
    def get(cls, device):
        key = torch.device(device).index or 0
        if key not in cls._states:
            cls._states[key] = cls(device, seed=torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        return cls._states[key]

    def seed(self, seed):
        self.state.copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))
        self._queue = []
        self._queue_ready = None

    def advance(self):
        self.state.add_(self._inc)

    def fork(self):
        """Snapshot for one dropout call site + advance of the live state, in one (graph-capturable) launch."""
        if self._queue:
            if self._queue_ready is not None:        # the snapshots were written on a side stream: wait for that launch only
                torch.cuda.current_stream(self.state.device).wait_event(self._queue_ready)
                self._queue_ready = None
            elif not _prep_events_enabled:
                Overlap.join_fwd(self.state.device)
            return self._queue.pop(0)
        snap = torch.empty_like(self.state)
        check(lib().eagcn_rng_fork(ptr(self.state), ptr(snap), 1 << 20, _stream(self.state.device)), "eagcn_rng_fork")
        return snap

    def prefork(self, n, stream=None):
        """The next ``n`` fork() calls in ONE launch (on ``stream``): they then only hand out the snapshots.  Same
        (seed, offset) sequence as n separate forks; snapshots left over from an earlier prefork are dropped."""
        snaps = torch.empty(n, 2, dtype=torch.int64, device=self.state.device)
        check(lib().eagcn_rng_fork_n(ptr(self.state), ptr(snaps), n, 1 << 20,
                                     _stream(self.state.device) if stream is None else ctypes.c_void_p(stream.cuda_stream)),
              "eagcn_rng_fork_n")
        self._queue = [snaps[i] for i in range(n)]
        self._queue_ready = None
        if stream is not None and _prep_events_enabled:
            self._queue_ready = torch.cuda.Event()
            self._queue_ready.record(stream)


def manual_seed(seed, device=None):
    """Seed the dropout generator of the CUDA path (independent of torch's generator, layers.py:94)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    RngState.get(dev).seed(seed)


def _layer_struct(plan: GraphPlan, cfg: LayerConfig, params, buffers):
    """params: per view (att_w, self_r, W, bias, gamma, beta); buffers: per view (run_mean, run_var, nbt)."""
    s = LayerStruct()
    s.fin, s.V = cfg.fin, plan.V
    off = 0
    for v in range(plan.V):
        s.fo[v] = cfg.fo[v]
        s.off[v] = off
        off += cfg.fo[v]
        a, r, W, b, g, be = params[6 * v: 6 * v + 6]
        rm, rv, nbt = buffers[3 * v: 3 * v + 3]
        s.att_w[v], s.self_r[v], s.W[v], s.bias[v] = a.data_ptr(), r.data_ptr(), W.data_ptr(), b.data_ptr()
        s.gamma[v], s.beta[v] = g.data_ptr(), be.data_ptr()
        s.run_mean[v], s.run_var[v] = rm.data_ptr(), rv.data_ptr()
        s.nbt[v] = nbt.data_ptr() if nbt is not None else None
    for v in range(plan.V, _lib.MAX_VIEWS + 1):
        s.off[v] = off
    s.fo_tot = off
    return s


def _check_params(plan, cfg, params, buffers):
    dev = plan.device
    for v in range(plan.V):
        a, r, W, b, g, be = params[6 * v: 6 * v + 6]
        fo = cfg.fo[v]
        exp = [(a, plan.channels[v]), (r, 1), (W, cfg.fin * fo), (b, fo), (g, fo), (be, fo),
               (buffers[3 * v], fo), (buffers[3 * v + 1], fo)]
        for t, n in exp:
            if t.device != dev or t.dtype != _F32 or t.numel() != n or not t.is_contiguous():
                raise ValueError("GraphConv_Layer parameter with unexpected device/dtype/shape "
                                 f"(view {v + 1}: expected {n} contiguous float32 on {dev}, got {tuple(t.shape)} "
                                 f"{t.dtype} on {t.device})")


class _GraphConvLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, cfg, buffers, H, *params):
        L = lib()
        dev = plan.device
        Overlap.join_pending(dev)           # a backward pass that aborted after forking: join before anything else
        if H.device != dev or H.dtype != _F32 or H.dim() != 2 or H.shape[0] != plan.t_cap or H.shape[1] != cfg.fin:
            raise ValueError(f"packed input must be float32 [{plan.t_cap}, {cfg.fin}] on {dev}, "
                             f"got {tuple(H.shape)} {H.dtype} on {H.device}")
        H = H.contiguous()
        ctx.param_refs = params
        if any(ctx.needs_input_grad[4:]):
            Overlap.note_use(params)
        params = tuple(p.detach() for p in params)
        _check_params(plan, cfg, params, buffers)
        ls = _layer_struct(plan, cfg, params, buffers)
        C = int(ls.fo_tot)
        T = plan.t_cap
        f32 = dict(dtype=_F32, device=dev)
        w = WorkStruct()
        Z = torch.empty(T, C, **f32)
        Y = torch.empty(T, C, **f32)
        X = torch.empty(T, C, **f32)
        invR = torch.empty(plan.V, T, **f32)
        prep = cfg.prep
        if prep is not None and not prep.matches(cfg, params, plan):
            prep = None
        if prep is not None:
            prep.join()                                 # the side stream filled these buffers
            wall, wallT, wsplit, ball, sig = prep.wall, prep.wallT, prep.wsplit, prep.ball, prep.sig
        else:
            wall, wallT, wsplit, ball, sig = _prep_buffers(cfg.fin, C, plan.V, dev)
        partial = torch.empty(int(L.eagcn_partial_floats(T, C, plan.V)), **f32)
        sums = torch.empty(2, C, dtype=torch.float64, device=dev)
        mean = torch.empty(C, **f32)
        invstd = torch.empty(C, **f32)
        rng = RngState.get(dev)
        rng_snapshot = None
        if cfg.training and cfg.p_drop > 0.0:
            rng_snapshot = rng.fork()               # backward regenerates the same keep mask from it
        w.H, w.Z, w.Y, w.X, w.invR = ptr(H), ptr(Z), ptr(Y), ptr(X), ptr(invR)
        w.wall, w.ball, w.sig, w.partial, w.sums = ptr(wall), ptr(ball), ptr(sig), ptr(partial), ptr(sums)
        w.wallT, w.wsplit = ptr(wallT), ptr(wsplit)
        w.mean, w.invstd = ptr(mean), ptr(invstd)
        w.rng = ptr(rng_snapshot)
        no_bwd = not any(ctx.needs_input_grad[3:])     # inference: the fused forward kernel then skips the store of Z
        w.training = ((1 if cfg.training else 0) | (2 if (cfg.training and cfg.stat_allreduce is not None) else 0) |
                      (4 if prep is not None else 0) | (16 if no_bwd else 0))
        w.rng_stream = int(cfg.rng_stream)
        w.m_total, w.n_pad = int(plan.m_total), int(plan.n_pad)
        w.p_drop, w.eps, w.momentum = float(cfg.p_drop), float(cfg.eps), float(cfg.momentum)
        st = _stream(dev)
        check(L.eagcn_layer_forward_a(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_forward_a")
        if cfg.training and cfg.stat_allreduce is not None:
            cfg.stat_allreduce(sums)
        check(L.eagcn_layer_forward_b(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_forward_b")
        ctx.plan, ctx.cfg, ctx.buffers, ctx.params = plan, cfg, buffers, params
        ctx.saved = (H, Z, Y, invR, wall, wsplit, ball, sig, mean, invstd, rng_snapshot)
        if cfg.want_pad:
            # ball rows: bias | gamma | beta.  mean / invstd are the statistics the kernels used (batch or running).
            z_pad = ball[1] * (ball[0] - mean) * invstd + ball[2]
            return X, z_pad
        return X

    @staticmethod
    def backward(ctx, dX, gz=None):
        L = lib()
        Overlap.note_backward()
        plan, cfg, buffers, params = ctx.plan, ctx.cfg, ctx.buffers, ctx.params
        H, Z, Y, invR, wall, wsplit, ball, sig, mean, invstd, rng_snapshot = ctx.saved
        dev = plan.device
        ls = _layer_struct(plan, cfg, params, buffers)
        C = int(ls.fo_tot)
        T = plan.t_cap
        f32 = dict(dtype=_F32, device=dev)
        dX = dX.contiguous() if dX is not None else torch.zeros(T, C, **f32)
        # gradient reaching the padded rows' common pre-activation: they are part of the BatchNorm population, so their
        # share S1_pad = sum g, S2_pad = sum g * xhat_pad of the two backward sums joins the active rows' sums
        pad_sums = None
        if gz is not None:
            xhat_pad = ((ball[0] - mean) * invstd).double()
            pad_sums = torch.stack((gz.double(), gz.double() * xhat_pad))
        w = WorkStruct()
        dY = torch.empty(T, C, **f32)
        Q = torch.empty(T, C, **f32)
        need_dH = ctx.needs_input_grad[3]
        dH = torch.empty(T, cfg.fin, **f32) if need_dH else None
        dwall = GradArena.empty(cfg.fin * C, dev)      # view-blocked: each view's [fin, fo_v] block contiguous
        dvec = GradArena.empty(3 * C, dev).view(3, C)
        datt = GradArena.empty(plan.V * _lib.SIG_STRIDE, dev).view(plan.V, _lib.SIG_STRIDE)
        bsums = torch.empty(2, C, dtype=torch.float64, device=dev)
        partial = torch.empty(int(L.eagcn_partial_floats(T, C, plan.V)), **f32)
        ws_bytes = int(L.eagcn_gemm_workspace_bytes(cfg.fin, C, T))
        gemm_ws = torch.empty(max(ws_bytes // 4, 1), **f32)
        w.H, w.Z, w.Y, w.invR = ptr(H), ptr(Z), ptr(Y), ptr(invR)
        w.wall, w.ball, w.sig, w.partial = ptr(wall), ptr(ball), ptr(sig), ptr(partial)
        w.wsplit = ptr(wsplit)
        w.mean, w.invstd, w.rng = ptr(mean), ptr(invstd), ptr(rng_snapshot)
        w.training = (1 if cfg.training else 0) | (2 if (cfg.training and cfg.stat_allreduce is not None) else 0)
        w.rng_stream = int(cfg.rng_stream)
        w.m_total, w.n_pad = int(plan.m_total), int(plan.n_pad)
        w.p_drop, w.eps, w.momentum = float(cfg.p_drop), float(cfg.eps), float(cfg.momentum)
        w.dX, w.dY, w.Q, w.dH, w.dwall, w.dvec, w.datt = ptr(dX), ptr(dY), ptr(Q), ptr(dH), ptr(dwall), ptr(dvec), ptr(datt)
        w.bsums, w.gemm_ws, w.gemm_ws_bytes = ptr(bsums), ptr(gemm_ws), ws_bytes
        if _bwd_tickets_enabled:
            w.tickets = ptr(_ticket_buf(dev, 2, C // 128 + 2))     # BatchNorm backward sums reduced by the last CTAs
        st = _stream(dev)
        check(L.eagcn_layer_backward_a(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_backward_a")
        if cfg.training and cfg.stat_allreduce is not None:
            cfg.stat_allreduce(bsums)
        if pad_sums is not None and cfg.training:
            bsums.add_(pad_sums)
        if Overlap.enabled and need_dH and pad_sums is None:
            # dW_all = H^T Q (+ parameter sums) on the side stream, beside dH = Q W_all^T (and, when the parameters
            # take fresh gradients, beside everything that follows in this backward pass)
            w.phase = 1
            check(L.eagcn_layer_backward_b(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_backward_b")
            forked = Overlap.fork_point(dev)            # the branch point: after the aggregation backward
            w.phase = 2                                 # dH first: it is the one on the critical path (next layer's input)
            check(L.eagcn_layer_backward_b(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_backward_b")
            side = Overlap.fork_from(dev, forked)
            w.phase = 4
            check(L.eagcn_layer_backward_b(plan.ref, ctypes.byref(ls), ctypes.byref(w), ctypes.c_void_p(side.cuda_stream)),
                  "eagcn_layer_backward_b")
            if Overlap.fresh(ctx.param_refs):
                Overlap.defer_join(dev, (H, Q, gemm_ws, dwall, partial, datt, Z, Y, dY))
            else:
                Overlap.join(dev)
        else:
            check(L.eagcn_layer_backward_b(plan.ref, ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_backward_b")
        if pad_sums is not None:                      # dvec rows: d bias | d gamma | d beta
            dvec[1].add_(pad_sums[1].float())
            dvec[2].add_(pad_sums[0].float())
            if not cfg.training:
                dvec[0].add_((ball[1] * invstd).mul(pad_sums[0].float()))
        grads = []
        off = 0
        for v in range(plan.V):
            fo = cfg.fo[v]
            a, r, W, b, g, be = params[6 * v: 6 * v + 6]
            grads += [datt[v, :plan.channels[v]].reshape(a.shape), datt[v, 256:257].reshape(r.shape),
                      dwall[cfg.fin * off: cfg.fin * (off + fo)].view(cfg.fin, fo),
                      dvec[0, off:off + fo], dvec[1, off:off + fo], dvec[2, off:off + fo]]
            off += fo
        return (None, None, None, dH, *grads)


def _prep_buffers(fin, C, V, dev):
    f32 = dict(dtype=_F32, device=dev)
    return (torch.empty(fin, C, **f32), torch.empty(2, C, fin, **f32), torch.empty(2, fin, C, **f32),
            torch.empty(4, C, **f32), torch.empty(V, _lib.SIG_STRIDE, **f32))


class LayerPrep:
    """Per-step parameter-derived buffers of one layer (W_all, its transposed / hi-lo split copies, bias / gamma / beta
    rows, sigmoid tables), filled by eagcn_layer_prepare -- on the side stream when ``stream`` is given, so that it runs
    beside the packing of the batch.  Valid until the parameters change (one optimiser step): ``matches`` compares the
    parameters' storage and ``_version`` counters, so an in-place update through ``.data`` (the reference's
    ``weights_init``, utils.py:702-708) between a prefetch and the forward pass it was made for is NOT detected -- prefetch
    right before the forward call (EAGCNStack.forward does) and after any such initialisation."""

    def __init__(self, fin, fo, channels, params, dev, stream=None):
        self.fin, self.fo, self.channels = int(fin), tuple(int(f) for f in fo), tuple(int(c) for c in channels)
        self.keys = tuple((p.data_ptr(), p._version) for p in params)
        C, V = sum(self.fo), len(self.fo)
        self.wall, self.wallT, self.wsplit, self.ball, self.sig = _prep_buffers(self.fin, C, V, dev)
        self.dev, self.side = dev, stream is not None
        ps = _lib.PlanStruct()
        ps.V = V
        for v, c in enumerate(self.channels):
            ps.chan[v] = c
        ls = LayerStruct()
        ls.fin, ls.V, ls.fo_tot = self.fin, V, C
        off = 0
        for v in range(V):
            ls.fo[v], ls.off[v] = self.fo[v], off
            off += self.fo[v]
            a, r, W, b, g, be = params[6 * v: 6 * v + 6]
            ls.att_w[v], ls.self_r[v], ls.W[v], ls.bias[v] = a.data_ptr(), r.data_ptr(), W.data_ptr(), b.data_ptr()
            ls.gamma[v], ls.beta[v] = g.data_ptr(), be.data_ptr()
        for v in range(V, _lib.MAX_VIEWS + 1):
            ls.off[v] = off
        w = WorkStruct()
        w.wall, w.wallT, w.wsplit, w.ball, w.sig = ptr(self.wall), ptr(self.wallT), ptr(self.wsplit), ptr(self.ball), ptr(self.sig)
        st = _stream(dev) if stream is None else ctypes.c_void_p(stream.cuda_stream)
        check(lib().eagcn_layer_prepare(ctypes.byref(ps), ctypes.byref(ls), ctypes.byref(w), st), "eagcn_layer_prepare")
        # the consumer waits for THIS layer's preparation only (not for the whole side stream): layer 2's preparation then
        # runs beside layer 1's forward pass instead of ahead of it
        self.ready = None
        if stream is not None and _prep_events_enabled:
            self.ready = torch.cuda.Event()
            self.ready.record(stream)

    def matches(self, cfg, params, plan):
        return (self.fin == cfg.fin and self.fo == tuple(cfg.fo) and self.channels == tuple(plan.channels[:plan.V]) and
                self.keys == tuple((p.data_ptr(), p._version) for p in params) and self.dev == plan.device)

    def join(self):
        if self.side:
            if self.ready is not None:
                torch.cuda.current_stream(self.dev).wait_event(self.ready)
            else:
                Overlap.join_fwd(self.dev)
            self.side = False


def pad_layer_args(cfg: LayerConfig, H, params, buffers):
    """Widths that are not multiples of 4 floats (HIV layer 2: five views of 250, row stride 1 250) cannot use the TMA /
    tcgen05 GEMM (16-byte row alignment) nor the float4 aggregation and BatchNorm kernels.  This builds the equivalent
    layer call on widths rounded up to 4: zero weight columns (and zero weight rows for padded input channels), zero bias,
    gamma 1, beta 0, running statistics 0 / 1 -- a padded channel is the constant 0 through projection, aggregation,
    BatchNorm (variance 0, finite inverse std), ReLU and dropout, so it neither changes the real channels nor produces
    non-finite values.  Returns (cfg_p, H_p, params_p, buffers_p, finish); ``finish(X_p)`` cuts the real channels out of
    the padded output (autograd routes the gradients of the padded parameters back through ``F.pad``) and copies the
    updated running statistics into the module's buffers.  Pure tensor code: shared by the CUDA path and its CPU test."""
    import torch.nn.functional as F
    V = len(cfg.fo)
    pad4 = lambda n: (int(n) + 3) & ~3
    fin_p, fo_p = pad4(cfg.fin), tuple(pad4(f) for f in cfg.fo)
    e = fin_p - cfg.fin
    H_p = F.pad(H, (0, e)) if e else H
    params_p, buffers_p = [], []
    for v in range(V):
        a, r, W, b, g, be = params[6 * v: 6 * v + 6]
        rm, rv, nbt = buffers[3 * v: 3 * v + 3]
        d = fo_p[v] - cfg.fo[v]
        params_p += [a, r, F.pad(W, (0, d, 0, e)), F.pad(b, (0, d)), F.pad(g, (0, d), value=1.0), F.pad(be, (0, d))]
        buffers_p += [F.pad(rm, (0, d)), F.pad(rv, (0, d), value=1.0), nbt]
    cfg_p = dataclasses.replace(cfg, fin=fin_p, fo=fo_p, prep=None)

    def finish(X_p):
        outs, off = [], 0
        for v in range(V):
            outs.append(X_p[:, off: off + cfg.fo[v]])
            off += fo_p[v]
            if cfg.training:
                with torch.no_grad():
                    buffers[3 * v].copy_(buffers_p[3 * v][:cfg.fo[v]])
                    buffers[3 * v + 1].copy_(buffers_p[3 * v + 1][:cfg.fo[v]])
        return torch.cat(outs, 1)
    return cfg_p, H_p, params_p, buffers_p, finish


def graph_conv_layer(plan: GraphPlan, cfg: LayerConfig, H, params, buffers):
    """X = GraphConv_Layer(H) on packed rows.  params: flat per-view (att_w, self_r, W, bias, gamma, beta);
    buffers: flat per-view (running_mean, running_var, num_batches_tracked)."""
    return _GraphConvLayerFn.apply(plan, cfg, tuple(buffers), H, *params)


# ------------------------------------------------------------------------------------------------
class _GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, dense):
        if dense.device != plan.device:
            raise _lib.EagcnError(f"eagcn_b200 is CUDA-only (no CPU fallback): features are on {dense.device}")
        if dense.dim() != 3 or dense.shape[0] != plan.B or dense.shape[1] != plan.N:
            raise ValueError(f"afms must be [B={plan.B}, N={plan.N}, F], got {tuple(dense.shape)}")
        ctx.plan = plan
        return plan.gather(dense.float())

    @staticmethod
    def backward(ctx, g):
        # d(dense): gradient of the gathered rows; padded / bond-less rows get exact zeros (their features
        # reach the output only through the dropped 1e-9 mask_tiny terms)
        return None, ctx.plan.scatter(g)


class _ScatterRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, packed):
        ctx.plan = plan
        return plan.scatter(packed)

    @staticmethod
    def backward(ctx, g):
        return None, ctx.plan.gather(g)


def gather_rows(plan, dense):
    return _GatherRowsFn.apply(plan, dense)


def scatter_rows(plan, packed):
    return _ScatterRowsFn.apply(plan, packed)


class _ReadoutSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, packed):
        ctx.plan = plan
        F = packed.shape[1]
        out = torch.empty(plan.B, F, dtype=_F32, device=plan.device)
        check(lib().eagcn_readout_sum(plan.ref, ptr(packed.contiguous()), ptr(out), F, _stream(plan.device)), "eagcn_readout_sum")
        return out

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        F = g.shape[1]
        out = torch.empty(plan.t_cap, F, dtype=_F32, device=plan.device)
        check(lib().eagcn_readout_sum_bwd(plan.ref, ptr(g.contiguous()), ptr(out), F, _stream(plan.device)),
              "eagcn_readout_sum_bwd")
        return None, out


def readout_sum(plan, packed):
    """torch.sum(x, 1) of models.py:108 on packed rows -> [B, F]."""
    return _ReadoutSumFn.apply(plan, packed)


# ------------------------------------------------------------------------------------------------
class _AttentionDenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *att_ws):
        s = LayerStruct()
        s.V = plan.V
        if any(ctx.needs_input_grad[1:]):
            Overlap.note_use(att_ws)                    # the layer's own function produces a gradient for these too
        att = tuple(a.detach().contiguous() for a in att_ws)
        for v in range(plan.V):
            s.att_w[v] = att[v].data_ptr()
        A = torch.empty(plan.V, plan.B, plan.N, plan.N, dtype=_F32, device=plan.device)
        check(lib().eagcn_attention_dense(plan.ref, ctypes.byref(s), ptr(A), _stream(plan.device)), "eagcn_attention_dense")
        ctx.plan, ctx.att = plan, att
        return A

    @staticmethod
    def backward(ctx, dA):
        plan, att = ctx.plan, ctx.att
        s = LayerStruct()
        s.V = plan.V
        for v in range(plan.V):
            s.att_w[v] = att[v].data_ptr()
        datt = torch.empty(plan.V, _lib.SIG_STRIDE, dtype=_F32, device=plan.device)
        check(lib().eagcn_attention_dense_bwd(plan.ref, ctypes.byref(s), ptr(dA.contiguous()), ptr(datt), _stream(plan.device)),
              "eagcn_attention_dense_bwd")
        return (None, *[datt[v, :plan.channels[v]].reshape(att[v].shape) for v in range(plan.V)])


def attention_dense(plan, att_ws):
    """A_weight = stack_v(sigmoid(att_v(rel_v)) * adj)  [V,B,N,N]  (layers.py:83,318)."""
    return _AttentionDenseFn.apply(plan, *att_ws)


def dropout_keep_mask(plan, cfg: LayerConfig, fo_tot, rng_state):
    """Test hook: the keep mask eagcn_layer_forward_b draws for (rng_state, cfg.rng_stream): u8 [t_cap, fo_tot]."""
    w = WorkStruct()
    w.rng = ptr(rng_state)
    w.p_drop, w.rng_stream = float(cfg.p_drop), int(cfg.rng_stream)
    keep = torch.empty(plan.t_cap, fo_tot, dtype=torch.uint8, device=plan.device)
    check(lib().eagcn_dropout_mask(plan.ref, ctypes.byref(w), fo_tot, ptr(keep), _stream(plan.device)), "eagcn_dropout_mask")
    return keep


# ------------------------------------------------------------------------------------------------
class _BnActFn(torch.autograd.Function):
    """dropout(relu(BatchNorm1d(x))) in one kernel per direction (models.py:112, 114-116, 119)."""

    @staticmethod
    def forward(ctx, cfg, bufs, x, gamma, beta):
        training, relu, p_drop, rng_stream, eps, momentum = cfg
        if not x.is_cuda:
            raise EagcnError("eagcn_b200.bn_act is CUDA-only (sm_100a, no CPU fallback)")
        x = x.contiguous()
        B, C = x.shape
        rm, rv, nbt = bufs
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=_F32, device=x.device)
        invstd = torch.empty(C, dtype=_F32, device=x.device)
        rng = RngState.get(x.device)
        snap = rng.fork() if (training and p_drop > 0.0) else None
        g, b = gamma.detach().contiguous(), beta.detach().contiguous()
        check(lib().eagcn_bn_act_forward(ptr(x), ptr(y), ptr(g), ptr(b), ptr(rm), ptr(rv),
                                         ptr(nbt) if nbt is not None else None, ptr(mean), ptr(invstd), B, C,
                                         int(training), int(relu), float(p_drop), ptr(snap) if snap is not None else None,
                                         int(rng_stream), float(momentum), float(eps), _stream(x.device)), "eagcn_bn_act_forward")
        ctx.cfg = cfg
        ctx.saved = (x, g, b, mean, invstd, snap)
        return y

    @staticmethod
    def backward(ctx, dy):
        training, relu, p_drop, rng_stream, eps, momentum = ctx.cfg
        x, g, b, mean, invstd, snap = ctx.saved
        B, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg, db = GradArena.empty(C, x.device), GradArena.empty(C, x.device)
        check(lib().eagcn_bn_act_backward(ptr(x), ptr(dy), ptr(g), ptr(b), ptr(mean), ptr(invstd), ptr(dx), ptr(dg),
                                          ptr(db), B, C, int(training), int(relu), float(p_drop),
                                          ptr(snap) if snap is not None else None, int(rng_stream), _stream(x.device)),
              "eagcn_bn_act_backward")
        return None, None, dx, dg, db


_tickets = {}
_prep_events_enabled = os.environ.get("EAGCN_PREP_EVENTS", "1") != "0"   # 0: a prefetch consumer joins the whole side stream
_bwd_tickets_enabled = os.environ.get("EAGCN_BWD_TICKETS", "1") != "0"   # 0: separate stat_reduce launch (A/B measurements)


def _ticket_buf(dev, lane, n):
    """Persistent zero-initialised int32 ticket array of the tile GEMM (the kernel leaves it zero).  One array per
    (device, lane): lane 0 serves the calls on the current stream, lane 1 those on the side stream -- calls within a
    lane are stream-ordered, calls of different lanes may run concurrently."""
    key = (torch.device(dev).index or 0, lane)
    t = _tickets.get(key)
    if t is None or t.numel() < n:
        t = torch.zeros(max(int(n), 1 << 14), dtype=torch.int32, device=dev)
        _tickets[key] = t
    return t


def _mm_tile_bufs(A, transA, B, transB, lane, grad=False):
    """Output, split-K workspace and ticket array of one ``_mm_tile`` call, allocated (and, for a new ticket array,
    zero-filled) on the CURRENT stream: a side-stream call gets them BEFORE its fork point, so the fork orders the
    allocation / zero-fill ahead of the side-stream kernel."""
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    L = lib()
    # a weight gradient goes to the gradient arena when one is installed
    C = GradArena.empty(M * N, A.device).view(M, N) if grad else torch.empty(M, N, dtype=_F32, device=A.device)
    nbytes = int(L.eagcn_mm_tile_workspace_bytes(M, N, K))
    ws = torch.empty(nbytes // 4, dtype=_F32, device=A.device) if nbytes else None
    tk = _ticket_buf(A.device, lane, int(L.eagcn_mm_tile_tickets(M, N))) if nbytes else None
    return C, ws, tk, nbytes


def _mm_tile(A, transA, B, transB, stream=None, bufs=None):
    """op(A) @ op(B) on the small-matrix tile kernel (split-K combined inside the launch; bit-reproducible)."""
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    if (B.shape[1] if transB else B.shape[0]) != K:
        raise ValueError("mm: inner dimensions differ")
    L = lib()
    C, ws, tk, nbytes = bufs if bufs is not None else _mm_tile_bufs(A, transA, B, transB, 0 if stream is None else 1)
    st = _stream(A.device) if stream is None else ctypes.c_void_p(stream.cuda_stream)
    check(L.eagcn_mm_tile(ptr(A), A.shape[1], int(transA), ptr(B), B.shape[1], int(transB), ptr(C), M, N, K,
                          ptr(ws), nbytes, ptr(tk), st), "eagcn_mm_tile")
    return C, ws


class _DenseMmFn(torch.autograd.Function):
    """y = x @ W (layers.py:382-388) with dX = dY @ W^T and dW = x^T @ dY on mm_tile.cu, the weight gradient on the
    side stream beside dX (Overlap)."""

    @staticmethod
    def forward(ctx, x, W, engine):
        if not x.is_cuda:
            raise EagcnError("eagcn_b200.dense_mm is CUDA-only (sm_100a, no CPU fallback)")
        ctx.w_ref = W
        if ctx.needs_input_grad[1]:
            Overlap.note_use((W,))
        x, W = x.contiguous(), W.detach().contiguous()
        ctx.save_for_backward(x, W)
        if engine != "tile":
            raise ValueError("dense_mm engines: 'tile' (mm_tile.cu)")
        return _mm_tile(x, False, W, False)[0]

    @staticmethod
    def backward(ctx, dy):
        Overlap.note_backward()
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = dW = None
        if need_dw and need_dx and Overlap.enabled:
            # every buffer the side-stream product touches exists (tickets zero-filled) before the branch point
            side_bufs = _mm_tile_bufs(x, True, dy, False, lane=1, grad=True)
            forked = Overlap.fork_point(dy.device)
            dx, ws_dx = _mm_tile(dy, False, W, True)      # first: the rest of the backward pass waits for it
            side = Overlap.fork_from(dy.device, forked)
            dW, ws = _mm_tile(x, True, dy, False, stream=side, bufs=side_bufs)
            if Overlap.fresh((ctx.w_ref,)):
                # dW itself is NOT kept: an extra reference would make autograd copy it instead of adopting it.
                # ws_dx stays referenced too: freed early, the allocator could hand it to a side-stream kernel
                Overlap.defer_join(dy.device, (x, dy, ws, ws_dx))
            else:
                Overlap.join(dy.device)
            return dx, dW, None
        if need_dx:
            dx = _mm_tile(dy, False, W, True)[0]
        if need_dw:
            dW = _mm_tile(x, True, dy, False, bufs=_mm_tile_bufs(x, True, dy, False, lane=0, grad=True))[0]
        return dx, dW, None


def dense_mm(x, W, engine="tile"):
    return _DenseMmFn.apply(x, W, engine)


def bn_act(x, bn, training, relu=False, p_drop=0.0, rng_stream=1000):
    """``F.dropout(F.relu(bn(x)), p_drop, training)`` (each stage optional) for an ``nn.BatchNorm1d`` ``bn`` on [B, C]."""
    momentum = bn.momentum if bn.momentum is not None else 0.1
    cfg = (bool(training), bool(relu), float(p_drop), int(rng_stream), float(bn.eps), float(momentum))
    return _BnActFn.apply(cfg, (bn.running_mean, bn.running_var, bn.num_batches_tracked), x, bn.weight, bn.bias)


def dropout_keep_mask_flat(rng_state, rng_stream, p_drop, total):
    """Test hook: keep mask of the flat index range [0, total) for (rng_state, rng_stream): u8 [total]."""
    keep = torch.empty(total, dtype=torch.uint8, device=rng_state.device)
    check(lib().eagcn_dropout_mask_flat(ptr(rng_state), int(rng_stream), float(p_drop), int(total), ptr(keep), _stream(rng_state.device)),
          "eagcn_dropout_mask_flat")
    return keep


def gemm_nt(A, B, m_dev, engine=0):
    """C = A @ B.T on the projection GEMM engine (0: tcgen05 3xTF32, 1: FFMA); rows >= m_dev[0] come back zero."""
    A, B = A.contiguous(), B.contiguous()
    C = torch.empty(A.shape[0], B.shape[0], dtype=_F32, device=A.device)
    check(lib().eagcn_gemm_nt(ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C), B.shape[0], A.shape[0], B.shape[0],
                              A.shape[1], ptr(m_dev), engine, _stream(A.device)), "eagcn_gemm_nt")
    return C


def gemm_tn(A, B, k_dev, engine=0):
    """C = A.T @ B over the first k_dev[0] rows (rows beyond must be zero) on the projection GEMM engine."""
    A, B = A.contiguous(), B.contiguous()
    M, N, K = A.shape[1], B.shape[1], A.shape[0]
    C = torch.empty(M, N, dtype=_F32, device=A.device)
    nbytes = int(lib().eagcn_gemm_workspace_bytes(M, N, K))
    ws = torch.empty(max(nbytes // 4, 1), dtype=_F32, device=A.device)
    check(lib().eagcn_gemm_tn(ptr(A), M, ptr(B), N, ptr(C), M, N, K, ptr(k_dev), ptr(ws), nbytes, engine, _stream(A.device)),
          "eagcn_gemm_tn")
    return C


def set_gemm_engine(name: str):
    """'tcgen05' (default: tensor cores with 3xTF32 compensation where the layout allows) or 'ffma'."""
    check(lib().eagcn_set_gemm_mode({"tcgen05": 0, "ffma": 1, "tcgen05-nt": 2}[name]), "eagcn_set_gemm_mode")


def set_tc_precision(name: str):
    """'fp32x3' (default: hi/lo-compensated 3xTF32, fp32-faithful, the parity mode) or 'tf32' (ONE TF32 tensor-core pass on
    the raw operands: ~1e-3 relative error, NOT within the 1e-5 parity bar -- the analogue of BASELINE.json's "bf16"
    configuration, for throughput reporting beside the strict mode)."""
    check(lib().eagcn_set_tc_passes({"fp32x3": 3, "tf32": 1}[name]), "eagcn_set_tc_passes")


def set_agg_engine(name: str):
    """'tile' (default: shared-memory tile aggregation kernels, BatchNorm backward folded into the backward one) or
    'generic' (warp-per-row kernels reading neighbours through L2, dY materialised)."""
    check(lib().eagcn_set_agg_mode({"tile": 0, "generic": 1}[name]), "eagcn_set_agg_mode")
