#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank builds the same (replicated) EC table, keeps its contiguous shard of the likelihood and runs
RCG and EM with the per-pass NCCL all-reduce; rank 0 compares with the oracle.  Then the hash-partitioned
path: each rank gets only the reads whose pattern hash falls in its range; results must be the same."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import msweep_b200 as M
from msweep_b200 import dist, synth

rank, world, local = dist.init()
torch.cuda.set_device(local)
nccl_id = dist.broadcast_bytes(M.nccl_unique_id() if rank == 0 else None, M.NCCL_ID_BYTES) if world > 1 else None
ctx = M.Context(local, rank, world, nccl_id)
wl = synth.generate(60000, 3000, 50, n_present=5, n_templates=400, p_noise=0.02, seed=5)

aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
assert (lik.ec_begin, lik.ec_begin + lik.n_ecs) == dist.shard_range(aln.n_ecs, rank, world), "shard range"
lik_mh = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=50)
res = {}
for name, code in (("rcg", M.ALGO_RCG), ("em", M.ALGO_EM)):
    res[name] = lik.vi_run(code)
res["rcg_mh"] = lik_mh.vi_run(M.ALGO_RCG)
lik_sp = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
res["rcg_sparse"] = lik_sp.vi_run(M.ALGO_RCG)       # sparse RCG, EC-sharded: a rejected step stalls every rank at the same iteration
res["em_sparse"] = lik_sp.vi_run(M.ALGO_EM)
mask_mh, hits_mh = lik_mh.mask(want_hits=True)

# hash-partitioned reads: rank r owns hash range [r, r+1) * 2^64 / world
rp = wl.row_ptr.astype(np.int64)
h = np.array([M.pattern_hash(wl.targets[rp[i]:rp[i + 1]]) for i in range(wl.n_reads)], np.uint64)
owner = (h.astype(np.float64) / 2.0 ** 64 * world).astype(np.int64).clip(0, world - 1)
lens = np.diff(rp)
owner[lens == 0] = np.arange(wl.n_reads)[lens == 0] % world       # unaligned reads: anywhere
mine = np.nonzero(owner == rank)[0]
my_ptr = np.zeros(len(mine) + 1, np.uint64); my_ptr[1:] = np.cumsum(lens[mine])
my_tg = np.concatenate([wl.targets[rp[i]:rp[i + 1]] for i in mine]) if len(mine) else np.zeros(0, np.uint32)
aln_p = M.Alignment(ctx, len(mine), wl.n_targets, my_ptr, my_tg, partitioned=True)
lik_p = M.Likelihood.build(ctx, aln_p, wl.group_of_target, wl.group_sizes)
res["rcg_part"] = lik_p.vi_run(M.ALGO_RCG)
n_total = dist.reduce_sum(aln_p.n_ecs)

if rank == 0:
    from oracle import pyoracle as orc
    ec = orc.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    ref_l = orc.lik_build(ec, wl.group_of_target, wl.group_sizes)
    assert int(n_total) == ec.n_ecs == lik_p.n_ecs_total, (n_total, ec.n_ecs)
    for name in ("rcg", "em", "rcg_sparse", "em_sparse"):
        ref = orc.vi_run(name.split("_")[0], ref_l.logl, ref_l.log_counts)
        got = res[name]
        assert got.iters == ref.iters, (name, got.iters, ref.iters)
        assert np.max(np.abs(got.theta - ref.theta)) < 1e-6 and abs(got.bound - ref.bound) <= 1e-9 * abs(ref.bound)
    ref = orc.vi_run("rcg", ref_l.logl, ref_l.log_counts)
    got = res["rcg_part"]
    assert got.iters == ref.iters and np.max(np.abs(got.theta - ref.theta)) < 1e-6
    ref_mh = orc.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=50)
    assert np.array_equal(mask_mh, ref_mh.mask) and np.array_equal(hits_mh, ref_mh.hits)
    ref = orc.vi_run("rcg", ref_mh.logl, ref_mh.log_counts)
    assert res["rcg_mh"].iters == ref.iters and np.max(np.abs(res["rcg_mh"].theta - ref.theta)) < 1e-6
    how = "peer-memory exchange" if ctx.peer_active else "NCCL all-reduce"
    print(f"multi-GPU parity ok on {world} GPUs: {ec.n_ecs} ECs, rcg {res['rcg'].iters} iters, em {res['em'].iters} iters, collective: {how}")
dist.barrier()
ctx.close()
dist.finalize()
