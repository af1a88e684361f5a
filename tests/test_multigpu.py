"""Multi-GPU build parity (needs >= 2 GPUs on the box; skipped otherwise).  The N-GPU build must give
the single-GPU result to FP64 summation-order noise (1e-11) and the reference golden W at 1e-10."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_build_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tools", "dist_build_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
