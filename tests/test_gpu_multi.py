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


@pytest.mark.parametrize("pou", [0, 1])
@pytest.mark.parametrize("ranks,slabs,axis", [(2, -3, 0), (2, -6, 2), (4, -4, 0)])
def test_pipelined_host_apply_across_ranks(ranks, slabs, axis, pou):
    """Host-space apply with neighbour ranks: interface pieces of x first, ghost-touching slabs first, the second
    neighbour barrier announced early and collected late, interface pieces of y last (forced on for this small case)."""
    import ssc_b200
    if ssc_b200.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + ranks + pou - slabs), os.path.join(ROOT, "tests", "mgpu_check.py"), "5", "3", str(pou), "1", str(slabs), str(axis)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout and "host_pipeline=%d" % slabs in out.stdout
