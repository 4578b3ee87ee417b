#!/usr/bin/env python
"""Latency probe for small problems (config 1 size): per-iteration and per-run overheads of the optimiser driver."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth
wl = synth.generate(300000, 3000, 50, n_present=5, n_templates=2000, p_noise=0.02, seed=20231017)
ctx = M.Context(0)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
print("ECs", aln.n_ecs, "K", lik.n_groups)
for algo, name in ((M.ALGO_RCG, "rcg"), (M.ALGO_EM, "em")):
    lik.vi_run(algo, max_iters=5, tol=-1e300)
    for iters in (8, 200):
        t0 = time.perf_counter(); r = lik.vi_run(algo, max_iters=iters, tol=-1e300 if name == "rcg" else 0.0); dt = time.perf_counter() - t0
        print(f"{name}: {r.iters} iterations in {dt*1e3:.2f} ms -> {dt/r.iters*1e6:.0f} us/iteration")
t0 = time.perf_counter(); th, it = lik.bootstrap_run(10, seed=5); dt = time.perf_counter() - t0
print(f"bootstrap: 10 replicates in {dt*1e3:.1f} ms ({np.mean(it):.0f} iterations each)")
