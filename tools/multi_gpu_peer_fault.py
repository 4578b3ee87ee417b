#!/usr/bin/env python
"""Fault injection for the peer-memory all-reduce (csrc/peer.cuh), run under torchrun with 2 ranks and a short
MSWB_PEER_TIMEOUT_S: rank 1 never enters the optimiser, so rank 0's control kernel waits for a vector that does not come.
The wait must END (time-out) and the call must fail with the library's message — not hang; the context must still close."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import msweep_b200 as M
from msweep_b200 import dist, synth

rank, world, local = dist.init()
assert world == 2
torch.cuda.set_device(local)
nccl_id = dist.broadcast_bytes(M.nccl_unique_id() if rank == 0 else None, M.NCCL_ID_BYTES)
ctx = M.Context(local, rank, world, nccl_id)
if not ctx.peer_active:
    if rank == 0:
        print("peer exchange not available on this box: nothing to inject")
    dist.barrier(); ctx.close(); dist.finalize(); sys.exit(0)
wl = synth.generate(20000, 600, 30, n_present=4, n_templates=200, seed=3)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
dist.barrier()
if rank == 0:
    t0 = time.time()
    try:
        lik.vi_run(M.ALGO_EM)
        print("FAILED: the run returned although the peer never arrived")
    except RuntimeError as e:
        dt = time.time() - t0
        assert "peer rank did not arrive" in str(e), str(e)
        print(f"peer fault ok: the wait ended after {dt:.1f} s with: {e}")
dist.barrier()
lik.close(); aln.close(); ctx.close()
dist.finalize()
