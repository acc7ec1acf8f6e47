"""GPU, needs >= 2 devices: N-rank data parallel with global-batch BatchNorm == single process.
Runs tools/dp_parity.py under torchrun; skipped on single-GPU boxes (the 2-GPU log of round 1 is committed as
profiles/r01_dp2_parity.log)."""
import os
import signal
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_global_bn_equals_single_process():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    # own process group, killed as a whole on timeout: killing torchrun alone would orphan ranks hung in a collective
    p = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dp_parity.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        pytest.fail("2-rank parity run timed out\n" + out[-2000:] + err[-2000:])
    assert "DP PARITY OK" in out, out[-2000:] + err[-2000:]
