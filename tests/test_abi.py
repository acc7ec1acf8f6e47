"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/eagcn_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "eagcn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eagcn_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    from eagcn_b200 import build, _lib
    path = build.build()
    assert os.path.isfile(path)
    L = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/eagcn_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared            # python binding covers exactly the header
    assert _lib.lib().eagcn_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    """All ABI structs are made of 8-byte fields; sizes must equal the C side's."""
    from eagcn_b200 import _lib
    assert ctypes.sizeof(_lib.PlanStruct) == 8 * (5 + 16 + 13)
    assert ctypes.sizeof(_lib.LayerStruct) == 8 * (3 + 16 + 17 + 9 * 16)
    assert ctypes.sizeof(_lib.WorkStruct) == 8 * 34
    L = _lib.lib()
    for which, cls in enumerate((_lib.PlanStruct, _lib.LayerStruct, _lib.WorkStruct)):
        assert L.eagcn_sizeof(which) == ctypes.sizeof(cls), cls.__name__


def test_size_helpers():
    from eagcn_b200 import _lib
    L = _lib.lib()
    assert L.eagcn_stat_tiles(128) == 4
    assert L.eagcn_partial_floats(128, 700, 5) == max(8 * 2 * 700, 8 * 5 * 257)
    assert L.eagcn_gemm_workspace_bytes(24, 400, 4864) > 0


def test_sass_is_sm100():
    import shutil
    import subprocess
    from eagcn_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_path_has_no_cpu_fallback():
    import torch
    from eagcn_b200 import layers as EL
    from eagcn_b200._lib import EagcnError
    layer = EL.GraphConv_Layer(24, 5, 4, 4, 4, 4, 4, dropout=0.0, structure="Concate")
    B, N = 2, 6
    adj = torch.zeros(B, N, N)
    with pytest.raises(EagcnError):
        layer(adj, torch.zeros(B, N, 24), torch.zeros(B, 5, N, N), torch.zeros(B, 4, N, N),
              torch.zeros(B, 2, N, N), torch.zeros(B, 2, N, N), torch.zeros(B, 2, N, N))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eagcn_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f
