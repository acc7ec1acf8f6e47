"""ctypes binding of libeagcn_sm100.so (the C ABI declared in include/eagcn_b200.h).

The library is built in-tree by ``eagcn_b200.build.build()`` (plain ``nvcc -gencode
arch=compute_100a,code=sm_100a``; no torch headers).  There is NO fallback: if the shared object is
missing or cannot be loaded the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeagcn_sm100.so")
ABI_VERSION = 18
MAX_VIEWS = 16
ROW_TILE = 128
SIG_STRIDE = 257

ST_ADJ_NOT_01, ST_NOT_ONEHOT, ST_ASYMMETRIC, ST_EDGE_CAP, ST_ROW_CAP = 1, 2, 4, 8, 16
STATUS_TEXT = {
    ST_ADJ_NOT_01: "adjacency holds values other than 0.0/1.0",
    ST_NOT_ONEHOT: "a relation tensor is not one-hot 0/1 on a bonded pair",
    ST_ASYMMETRIC: "adjacency is not symmetric",
    ST_EDGE_CAP: "more directed edges than the plan's edge capacity (e_cap)",
    ST_ROW_CAP: "more active atom rows than the plan's row capacity (t_cap)",
}

_PV = c_void_p * MAX_VIEWS
_IV = c_int64 * MAX_VIEWS


class PlanStruct(ctypes.Structure):
    _fields_ = [("B", c_int64), ("N", c_int64), ("V", c_int64), ("t_cap", c_int64), ("e_cap", c_int64),
                ("chan", _IV),
                ("counts", c_void_p), ("deg", c_void_p), ("blk", c_void_p), ("pos_row", c_void_p),
                ("row_pos", c_void_p), ("row_ptr", c_void_p), ("mol_ptr", c_void_p), ("col", c_void_p),
                ("colpos", c_void_p), ("rev", c_void_p), ("code", c_void_p), ("rcode", c_void_p), ("tile_row", c_void_p)]


class LayerStruct(ctypes.Structure):
    _fields_ = [("fin", c_int64), ("fo_tot", c_int64), ("V", c_int64), ("fo", _IV),
                ("off", c_int64 * (MAX_VIEWS + 1)),
                ("att_w", _PV), ("self_r", _PV), ("W", _PV), ("bias", _PV), ("gamma", _PV), ("beta", _PV),
                ("run_mean", _PV), ("run_var", _PV), ("nbt", _PV)]


class WorkStruct(ctypes.Structure):
    _fields_ = [("H", c_void_p), ("Z", c_void_p), ("Y", c_void_p), ("X", c_void_p), ("invR", c_void_p),
                ("wall", c_void_p), ("ball", c_void_p), ("sig", c_void_p), ("partial", c_void_p),
                ("sums", c_void_p), ("mean", c_void_p), ("invstd", c_void_p), ("rng", c_void_p),
                ("training", c_int64), ("rng_stream", c_int64), ("m_total", c_int64), ("n_pad", c_int64),
                ("p_drop", c_double), ("eps", c_double), ("momentum", c_double),
                ("dX", c_void_p), ("dY", c_void_p), ("Q", c_void_p), ("dH", c_void_p), ("dwall", c_void_p),
                ("dvec", c_void_p), ("datt", c_void_p), ("bsums", c_void_p), ("gemm_ws", c_void_p),
                ("gemm_ws_bytes", c_int64), ("wallT", c_void_p), ("wsplit", c_void_p), ("phase", c_int64),
                ("tickets", c_void_p)]


_PROTOS = {
    "eagcn_version": (c_int, []),
    "eagcn_sizeof": (c_int64, [c_int]),
    "eagcn_stat_tiles": (c_int64, [c_int64]),
    "eagcn_partial_floats": (c_int64, [c_int64, c_int64, c_int64]),
    "eagcn_gemm_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "eagcn_pack_count": (c_int, [c_void_p, c_void_p, c_void_p]),
    "eagcn_pack_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_pack_count_codes": (c_int, [c_void_p, c_void_p, c_void_p]),
    "eagcn_pack_fill_codes": (c_int, [c_void_p, c_void_p, c_void_p]),
    "eagcn_unpack_view": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "eagcn_rows_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "eagcn_rows_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "eagcn_readout_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "eagcn_readout_sum_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "eagcn_layer_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_layer_forward_a": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_layer_forward_b": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_layer_backward_a": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_layer_backward_b": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_attention_dense": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_attention_dense_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "eagcn_dropout_mask": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "eagcn_set_gemm_mode": (c_int, [c_int]),
    "eagcn_get_gemm_mode": (c_int, []),
    "eagcn_bn_act_forward": (c_int, [c_void_p] * 9 + [c_int64, c_int64, c_int, c_int, c_double, c_void_p, c_int64,
                                     c_double, c_double, c_void_p]),
    "eagcn_bn_act_backward": (c_int, [c_void_p] * 9 + [c_int64, c_int64, c_int, c_int, c_double, c_void_p, c_int64,
                                      c_void_p]),
    "eagcn_set_bn_act_mode": (c_int, [c_int]),
    "eagcn_get_bn_act_mode": (c_int, []),
    "eagcn_mm_tile_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "eagcn_mm_tile_tickets": (c_int64, [c_int64, c_int64]),
    "eagcn_mm_tile": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                              c_int64, c_void_p, c_void_p]),
    "eagcn_gemm_trace": (c_int, [c_void_p, c_int64]),
    "eagcn_gemm_trace_stride": (c_int64, []),
    "eagcn_rng_fork": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "eagcn_rng_fork_n": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "eagcn_set_fwd_fused": (c_int, [c_int]),
    "eagcn_get_fwd_fused": (c_int, []),
    "eagcn_set_fuse_mode": (c_int, [c_int]),
    "eagcn_get_fuse_mode": (c_int, []),
    "eagcn_set_tc_bk": (c_int, [c_int]),
    "eagcn_set_tc_passes": (c_int, [c_int]),
    "eagcn_get_tc_passes": (c_int, []),
    "eagcn_set_tc_a_tmem": (c_int, [c_int]),
    "eagcn_get_tc_a_tmem": (c_int, []),
    "eagcn_set_pdl": (c_int, [c_int]),
    "eagcn_get_pdl": (c_int, []),
    "eagcn_set_agg_mode": (c_int, [c_int]),
    "eagcn_get_agg_mode": (c_int, []),
    "eagcn_gemm_nt": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                              c_void_p, c_int, c_void_p]),
    "eagcn_gemm_tn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                              c_void_p, c_int64, c_int, c_void_p]),
    "eagcn_dropout_mask_flat": (c_int, [c_void_p, c_int64, c_double, c_int64, c_void_p, c_void_p]),
    "eagcn_launch_count": (c_int64, []),
    "eagcn_spin": (c_int, [c_int64, c_void_p]),
    "eagcn_profile": (c_int, [c_int]),
    "eagcn_profile_report": (c_int64, [ctypes.c_char_p, c_int64]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


class EagcnError(RuntimeError):
    pass


def lib():
    """The loaded shared library (loads on first use; raises if it is not built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise EagcnError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). eagcn_b200 has no CPU / PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        v = L.eagcn_version()
        if v != ABI_VERSION:
            raise EagcnError(f"libeagcn_sm100.so ABI {v} != python binding {ABI_VERSION}: rebuild")
        for which, cls in enumerate((PlanStruct, LayerStruct, WorkStruct)):
            if L.eagcn_sizeof(which) != ctypes.sizeof(cls):
                raise EagcnError(f"{cls.__name__}: ctypes mirror is {ctypes.sizeof(cls)} bytes, the library's struct "
                                 f"{L.eagcn_sizeof(which)}: rebuild")
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise EagcnError(f"{what}: invalid argument (code {rc})")
    raise EagcnError(f"{what}: CUDA error {rc}")


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(lib().eagcn_launch_count())


def profile(enable: bool):
    lib().eagcn_profile(1 if enable else 0)


def profile_report() -> dict:
    """{kernel: (launches, total_ms)} recorded since profile(True)."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().eagcn_profile_report(buf, len(buf))
    return {k: (int(v[0]), float(v[1])) for k, v in json.loads(buf.value.decode() or "{}").items()} if n else {}
