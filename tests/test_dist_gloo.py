"""CPU, world_size = 2 over gloo: the host-side plumbing of the N > 1 path (rendezvous, shipping the NCCL id
bytes, max-over-ranks timing, EC shard ranges, bootstrap replicate ownership) and the EC-sharded algebra of
one VI pass: per-rank partial sums + a K+1 all-reduce reproduce the single-rank pass."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from msweep_b200 import dist as D
    rank, world, local = D.init("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    # 1. the 128-byte id travels intact from rank 0
    payload = bytes(range(128)) if rank == 0 else None
    got = D.broadcast_bytes(payload, 128)
    assert got == bytes(range(128))
    # 2. timings: max over ranks, sums over ranks
    assert D.reduce_max(1.0 + rank) == 2.0 and D.reduce_sum(3.0) == 6.0
    # 3. shard ranges tile [0, N) contiguously and match the C rule floor(N r / W)
    N = 1001
    lo, hi = D.shard_range(N, rank, world)
    assert (lo, hi) == (N * rank // world, N * (rank + 1) // world)
    edges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(edges, torch.tensor([lo, hi]))
    assert edges[0][0] == 0 and edges[0][1] == edges[1][0] and edges[1][1] == N
    # 4. replicate ownership is a partition of the replicates
    mine = [r for r in range(13) if D.replicate_owner(r, world) == rank]
    cnt = torch.tensor([len(mine)]); dist.all_reduce(cnt); assert cnt.item() == 13
    # 5. EC-sharded EM pass: partial (A_k, sum c log S) + all-reduce == the single-rank pass
    rng = np.random.default_rng(0)
    K = 7
    logl = rng.normal(-4, 2, size=(N, K)); c = rng.integers(1, 9, size=N).astype(float)
    dg = rng.normal(0, 1, size=K)
    def partial(rows):
        x = logl[rows] + dg; m = x.max(1, keepdims=True); s = np.exp(x - m).sum(1, keepdims=True)
        q = np.exp(x - m) / s
        return np.concatenate([(c[rows, None] * q).sum(0), [(c[rows] * (m[:, 0] + np.log(s[:, 0]))).sum()]])
    part = torch.from_numpy(partial(slice(lo, hi)))
    dist.all_reduce(part)
    full = partial(slice(0, N))
    assert np.allclose(part.numpy(), full, rtol=1e-13, atol=0)
    D.barrier(); D.finalize()
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
