"""GraphPlan: the packed form of one padded molecular batch, shared by all layers of a step.

Built once per batch from what the reference's collate hands over (adj [B,N,N] + the V one-hot
relation tensors, utils.py:560-566) -- or from the compact uint8 edge-code layout of the packed data
boundary -- and then reused by every GraphConv_Layer forward and backward call (the reference instead
rebuilds masks and re-reads all one-hot planes in each of its 4 layers, layers.py:294-304, :82).

No host synchronisation when capacities are supplied (``t_cap`` / ``e_cap``), which keeps the whole
training step CUDA-graph capturable; without them one tiny D2H read sizes the plan exactly.
Input validation (0/1 adjacency, symmetry, one-hot relations) happens on the device and is surfaced
lazily through ``check()`` / ``poll()``.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import PlanStruct, ROW_TILE, check, lib, ptr


def _round_up(x, m):
    return (int(x) + m - 1) // m * m


def _stream(device=None):
    """torch's current stream ON ``device`` (the plan's / tensor's device, not whatever device happens to be current)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class GraphPlan:
    def __init__(self, B, N, channels, device):
        self.B, self.N = int(B), int(N)
        self.channels = tuple(int(c) for c in channels)
        self.V = len(self.channels)
        if not (1 <= self.V <= _lib.MAX_VIEWS):
            raise ValueError(f"1..{_lib.MAX_VIEWS} views supported, got {self.V}")
        if any(c < 1 or c > 254 for c in self.channels):
            raise ValueError("relation channel counts must be in 1..254")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EagcnError("eagcn_b200 is CUDA-only (sm_100a); got tensors on %s" % self.device)
        self.t_cap = 0
        self.e_cap = 0
        self.m_total = self.B * self.N      # BatchNorm population (override for global-batch DP)
        self.n_pad = self.N                 # padded width in the (N - deg) * 1e-9 normaliser term
        self.struct = PlanStruct()
        self._pending = None
        P = self.B * self.N
        i32 = dict(dtype=torch.int32, device=self.device)
        self.counts = torch.zeros(8, **i32)
        self.deg = torch.empty(P, **i32)
        self.blk = torch.empty(2 * ((P + 31) // 32) + 2, **i32)
        self.pos_row = torch.empty(P, **i32)
        self.mol_ptr = torch.empty(self.B + 1, **i32)

    # ------------------------------------------------------------------------------------
    def _alloc(self, t_cap, e_cap):
        self.t_cap = max(ROW_TILE, _round_up(t_cap, ROW_TILE))
        self.e_cap = max(1, int(e_cap))
        i32 = dict(dtype=torch.int32, device=self.device)
        self.row_pos = torch.empty(self.t_cap, **i32)
        self.row_ptr = torch.empty(self.t_cap + 1, **i32)
        self.col = torch.empty(self.e_cap, **i32)
        self.colpos = torch.empty(self.e_cap, **i32)
        self.rev = torch.empty(self.e_cap, **i32)
        self.code = torch.empty(self.V, self.e_cap, dtype=torch.uint8, device=self.device)
        self.rcode = torch.empty(self.V, self.e_cap, dtype=torch.uint8, device=self.device)
        self.tile_row = torch.empty(self.t_cap // 32 + 2, **i32)      # molecule-aligned row tiles (fused layer kernel)
        self._fill_struct()

    def _fill_struct(self):
        s = self.struct
        s.B, s.N, s.V, s.t_cap, s.e_cap = self.B, self.N, self.V, self.t_cap, self.e_cap
        for v in range(_lib.MAX_VIEWS):
            s.chan[v] = self.channels[v] if v < self.V else 0
        for name in ("counts", "deg", "blk", "pos_row", "mol_ptr"):
            setattr(s, name, getattr(self, name).data_ptr())
        for name in ("row_pos", "row_ptr", "col", "colpos", "rev", "code", "rcode", "tile_row"):
            t = getattr(self, name, None)
            setattr(s, name, t.data_ptr() if t is not None else None)

    @property
    def ref(self):
        return ctypes.byref(self.struct)

    # ------------------------------------------------------------------------------------
    @classmethod
    def build(cls, adj, rels, t_cap=None, e_cap=None):
        """From the dense layout of the reference collate: adj [B,N,N] f32, rels: V x [B,C_v,N,N] f32.

        The relation tensors may also be PINNED HOST tensors (``tensor.pin_memory()``): the packer only reads them
        at bonded pairs, so it gathers those ~E*sum(C_v) values straight out of page-locked host memory over PCIe
        (zero-copy, unified addressing) instead of first copying the whole 4*(Kb+10)*N^2 bytes per molecule to the
        device -- ~25x fewer bytes across the bus for a Tox21 batch.  ``adj`` must be on the device."""
        if not adj.is_cuda:
            raise _lib.EagcnError("eagcn_b200 is CUDA-only (no CPU fallback): adj is on %s" % adj.device)
        if adj.dim() != 3 or adj.shape[1] != adj.shape[2]:
            raise ValueError("adj must be [B,N,N]")
        B, N = adj.shape[0], adj.shape[1]
        rels = [r if r.dtype == torch.float32 else r.float() for r in rels]   # layers.py:82 '.float()'
        for r in rels:
            on_dev = r.device == adj.device
            zero_copy = r.device.type == "cpu" and r.is_pinned()
            if r.dim() != 4 or r.shape[0] != B or r.shape[2] != N or r.shape[3] != N or not (on_dev or zero_copy):
                raise ValueError("relation tensors must be [B,C_v,N,N] on the adjacency's device "
                                 "(or pinned host tensors for the zero-copy gather)")
        adj = adj.contiguous() if adj.dtype == torch.float32 else adj.float().contiguous()
        rels = [r.contiguous() for r in rels]
        self = cls(B, N, [r.shape[1] for r in rels], adj.device)
        self._src = (adj, rels)
        self._run(adj, rels, None, t_cap, e_cap)
        return self

    @classmethod
    def from_codes(cls, codes, channels, t_cap=None, e_cap=None):
        """From the packed data boundary: codes u8 [B,V,N,N], 255 = no bond (eagcn_b200.data.MolBatch.codes)."""
        if not codes.is_cuda:
            raise _lib.EagcnError("eagcn_b200 is CUDA-only (no CPU fallback): codes are on %s" % codes.device)
        if codes.dtype != torch.uint8 or codes.dim() != 4 or codes.shape[2] != codes.shape[3]:
            raise ValueError("codes must be uint8 [B,V,N,N]")
        if codes.shape[1] != len(channels):
            raise ValueError("codes.shape[1] must equal the number of views")
        codes = codes.contiguous()
        self = cls(codes.shape[0], codes.shape[2], channels, codes.device)
        self._src = (codes,)
        self._run(None, None, codes, t_cap, e_cap)
        return self

    def _run(self, adj, rels, codes, t_cap, e_cap):
        L = lib()
        st = _stream(self.device)
        sync_sizes = t_cap is None or e_cap is None
        if sync_sizes:
            # phase 1 with unlimited capacities, read (T, E) once, allocate exactly
            self.t_cap, self.e_cap = _round_up(self.B * self.N, ROW_TILE), 2 ** 31 - 2
            self._fill_struct()
        else:
            self._alloc(t_cap, e_cap)
        if codes is None:
            check(L.eagcn_pack_count(self.ref, ptr(adj), st), "eagcn_pack_count")
        else:
            check(L.eagcn_pack_count_codes(self.ref, ptr(codes), st), "eagcn_pack_count_codes")
        if sync_sizes:
            T, E = (int(x) for x in self.counts[:2].tolist())
            self._alloc(max(T, 1) if t_cap is None else t_cap, max(E, 1) if e_cap is None else e_cap)
        if codes is None:
            arr = (ctypes.c_void_p * self.V)(*[r.data_ptr() for r in rels])
            check(L.eagcn_pack_fill(self.ref, ptr(adj), arr, st), "eagcn_pack_fill")
        else:
            check(L.eagcn_pack_fill_codes(self.ref, ptr(codes), st), "eagcn_pack_fill_codes")

    # ------------------------------------------------------------------------------------
    def check(self):
        """Synchronous validation of the packed inputs; raises ValueError on a malformed batch."""
        st = int(self.counts[2].item())
        if st:
            raise ValueError("malformed molecular batch: " +
                             "; ".join(txt for bit, txt in _lib.STATUS_TEXT.items() if st & bit))
        return self

    def check_async(self):
        """Start a non-blocking copy of the status word; ``poll()`` raises later if it was bad."""
        host = torch.empty(8, dtype=torch.int32, pin_memory=True)
        host.copy_(self.counts, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending = (host, ev)

    def poll(self, wait=False):
        if self._pending is None:
            return
        host, ev = self._pending
        if wait:
            ev.synchronize()
        if ev.query():
            self._pending = None
            st = int(host[2])
            if st:
                raise ValueError("malformed molecular batch (reported late): " +
                                 "; ".join(txt for bit, txt in _lib.STATUS_TEXT.items() if st & bit))

    @property
    def n_rows(self):      # synchronises
        return int(self.counts[0].item())

    @property
    def n_edges(self):     # synchronises
        return int(self.counts[1].item())

    # ------------------------------------------------------------------------------------
    def gather(self, dense):
        """dense [B,N,F] -> packed rows [t_cap,F] (no autograd; see functional.gather_rows)."""
        F = dense.shape[-1]
        out = torch.empty(self.t_cap, F, dtype=torch.float32, device=self.device)
        check(lib().eagcn_rows_gather(self.ref, ptr(dense.contiguous()), ptr(out), F, _stream(self.device)), "eagcn_rows_gather")
        return out

    def scatter(self, packed):
        F = packed.shape[-1]
        out = torch.empty(self.B, self.N, F, dtype=torch.float32, device=self.device)
        check(lib().eagcn_rows_scatter(self.ref, ptr(packed.contiguous()), ptr(out), F, _stream(self.device)),
              "eagcn_rows_scatter")
        return out

    def unpack_view(self, v):
        """(one-hot relation tensor of view v, adjacency) re-expanded from the plan (bit-exactness check)."""
        rel = torch.empty(self.B, self.channels[v], self.N, self.N, dtype=torch.float32, device=self.device)
        adj = torch.empty(self.B, self.N, self.N, dtype=torch.float32, device=self.device)
        check(lib().eagcn_unpack_view(self.ref, v, ptr(rel), ptr(adj), _stream(self.device)), "eagcn_unpack_view")
        return rel, adj

    def row_mask(self):
        """m[b,i] of layers.py:295 as a float tensor [B,N] (1 = atom with at least one bond)."""
        return (self.pos_row >= 0).to(torch.float32).view(self.B, self.N)
