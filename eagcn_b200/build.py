"""In-tree build of libeagcn_sm100.so: one nvcc invocation, sm_100a only, no torch headers."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "eagcn_sm100.cu")
OUT = os.path.join(HERE, "libeagcn_sm100.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _newest_source():
    t = 0.0
    for root in (os.path.join(HERE, "csrc"), os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and os.path.isfile(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    if not os.path.isfile(nvcc):
        if os.path.isfile(OUT):
            return OUT                     # GPU box: prebuilt library travels with the snapshot
        raise RuntimeError("nvcc not found and libeagcn_sm100.so is not built")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT
