"""Config-5-shaped bootstrap on one GPU acting as rank `--rank` of `--world`: time mswb_bootstrap_run with the generator
stepping through the other ranks' draws (MSWB_MT_JUMP=0) and jumping over them (=1); the replicates must be identical."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import msweep_b200 as M  # noqa: E402
from msweep_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ecs", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=1000)
ap.add_argument("--reps", type=int, default=100)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=3)
a = ap.parse_args()

wl = synth.generate_ec_patterns(a.ecs, a.k, 60, n_present=30, seed=5, dup_factor=10.0)
ctx = M.Context(0)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
out = {"ecs": int(aln.n_ecs), "draws_per_replicate": int(aln.n_aligned), "replicates": a.reps, "world": a.world, "rank": a.rank}
res = {}
for mode in ("0", "1", "0", "1"):
    os.environ["MSWB_MT_JUMP"] = mode
    t0 = time.perf_counter()
    th, it = lik.bootstrap_run(a.reps, seed=1, algo=M.ALGO_RCG, replica_rank=a.rank, replica_world=a.world)
    dt = time.perf_counter() - t0
    out.setdefault("seconds_jump" + mode, []).append(round(dt, 3))
    res[mode] = th
out["identical"] = bool(np.array_equal(res["0"], res["1"], equal_nan=True))
print(json.dumps(out))
