"""Multi-GPU parity through torchrun (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", [0, 1])
@pytest.mark.parametrize("pou", [0, 1])
def test_two_ranks_match_single_rank_oracle(pou, p2p):
    import ssc_b200
    if ssc_b200.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29533 + 2 * pou + p2p), os.path.join(ROOT, "tests", "mgpu_check.py"), "4", "3", str(pou), str(p2p)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout
