"""Process-group plumbing for one-process-per-GPU runs (bench.py, multi-GPU examples).

torch.distributed is used only to get the ranks talking (rendezvous, shipping the NCCL unique id,
max-over-ranks of timings); the data-path collective — the per-pass all-reduce of K+2 doubles — is
issued inside libmsweep_b200 on its own NCCL communicator.  Everything here also works with the
`gloo` backend on CPU, which is how tests/test_dist_gloo.py covers it.
"""
from __future__ import annotations

import os

import numpy as np


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None):
    """Initialises torch.distributed when WORLD_SIZE > 1. Returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload: bytes | None, n_bytes: int, src: int = 0) -> bytes:
    """Ships `n_bytes` raw bytes from rank `src` to everyone (the NCCL unique id of the library)."""
    _, world, _ = env_world()
    if world == 1:
        assert payload is not None
        return payload
    import torch
    import torch.distributed as dist
    buf = torch.zeros(n_bytes, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def barrier() -> None:
    _, world, _ = env_world()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def reduce_max(x: float) -> float:
    _, world, _ = env_world()
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float) -> float:
    _, world, _ = env_world()
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced range of rank `rank` — the same rule as mswb_shard_range."""
    return n * rank // world, n * (rank + 1) // world


def replicate_owner(replicate: int, world: int) -> int:
    """Bootstrap replicate r runs on rank r % world (mswb_bootstrap_run's rule)."""
    return replicate % world


def finalize() -> None:
    _, world, _ = env_world()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
