"""GPU, >= 2 devices: the EC-sharded path with its NCCL all-reduce against the oracle (tools/multi_gpu_check.py under
torchrun) and the multi-GPU command line (tools/multi_gpu_cli_check.py).  Skipped on single-GPU boxes; the CPU side
of the N > 1 plumbing is covered by tests/test_dist_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(tool, env=None, timeout=600):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(_port()), os.path.join(ROOT, "tools", tool)], capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("peer", ["1", "0"])
def test_sharded_vi_matches_oracle_on_two_gpus(peer):
    """peer = 1: the per-pass all-reduce as a one-shot exchange over NVLink peer memory inside the control kernel
    (csrc/peer.cuh, the default); peer = 0: ncclAllReduce + control kernel.  Same answers, same iteration counts."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun("multi_gpu_check.py", {"MSWB_PEER": peer, "MSWB_PEER_TIMEOUT_S": "60"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU parity ok on 2 GPUs" in r.stdout
    if peer == "0":
        assert "collective: NCCL all-reduce" in r.stdout


def test_peer_exchange_gives_up_on_a_missing_rank():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun("multi_gpu_peer_fault.py", {"MSWB_PEER_TIMEOUT_S": "2"}, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer fault ok" in r.stdout or "nothing to inject" in r.stdout


def test_cli_on_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_cli_check.py"), "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU CLI ok on 2 GPUs" in r.stdout
