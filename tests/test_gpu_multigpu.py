"""Multi-GPU checks (need >= 2 visible GPUs, skipped otherwise).  The partitioned path -- METIS or brick partition, the
library's own NCCL communicator, NCCL halo and fused peer-memory halo, distributed CG / Newton -- against a SERIAL
assembly / solve of the same global mesh (tests/run_comm_check.py: plain processes + ctypes, no torch.distributed)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_halo_matches_nccl():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "run_peer_check.py"), "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout and "FAIL" not in out.stdout, out.stdout[-2000:]


def test_partitioned_loads_match_serial():
    """body force + tractions on a 2-rank partition (NCCL halo and fused peer halo) against the serial assembly"""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "run_loads_check.py"), "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout and "FAIL" not in out.stdout, out.stdout[-2000:]


def test_library_collective_plane_eight_ranks():
    """the same check on an 8-way METIS split (edge / corner nodes ghosted by several ranks); needs 8 GPUs"""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 8:
        pytest.skip("needs eight GPUs")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "run_comm_check.py"), "8", "8", "metis"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "OK" in out.stdout and "FAIL" not in out.stdout, out.stdout[-3000:]


@pytest.mark.parametrize("how", ["metis", "brick"])
def test_library_collective_plane_matches_serial(how):
    """fecb200_comm_init / halo_sum / halo_update / comm_peer_enable + distributed CG and Newton on 2 ranks, every
    number against the serial handle; Newton iteration counts equal (north_star)."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "run_comm_check.py"), "2", "16", how]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "OK" in out.stdout and "FAIL" not in out.stdout, out.stdout[-3000:]
