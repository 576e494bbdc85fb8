"""Multi-GPU parity as a test: the N-rank tree equals the 1-rank tree through every entry path (whole host matrix, compact host
types, own row block per rank, adopted device arrays, replicated prepare), load errors are agreed across ranks, a slot budget of 1
still builds the whole tree, and the non-default variants run as checked replicas.  Needs >= 2 CUDA devices on the box (skipped
otherwise; the round's runs at N = 2 and N = 8 are logged in profiles/r2_mgpu_check_n*.log)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        from too_many_cells_b200 import _lib
        return _lib.load().tmc_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs at least 2 CUDA devices")
def test_two_rank_tree_equals_single_rank_tree_through_every_entry_path():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "mgpu_check.py"), "tiny", "small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "FAIL" not in r.stdout
