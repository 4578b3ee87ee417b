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


def test_sharded_vi_matches_oracle_on_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_port()), os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU parity ok on 2 GPUs" in r.stdout


def test_cli_on_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_cli_check.py"), "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU CLI ok on 2 GPUs" in r.stdout
