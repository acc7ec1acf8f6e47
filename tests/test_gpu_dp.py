"""GPU, needs >= 2 devices: N-rank data parallel with global-batch BatchNorm == single process.
Runs tools/dp_parity.py under torchrun; skipped on single-GPU boxes (the 2-GPU log of round 1 is committed as
profiles/r01_dp2_parity.log)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_global_bn_equals_single_process():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dp_parity.py")],
                       capture_output=True, text=True, timeout=300)
    assert "DP PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
