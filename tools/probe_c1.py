#!/usr/bin/env python
"""Config 1 shape (27,813 classes x 50 lineages): where does the optimiser's time go?  Cold run (first launches load the
kernels), warm run, and the steady-state time per iteration through the stepwise interface."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

wl = synth.generate(1_000_000, 3000, 50, n_present=5, n_templates=2000, p_noise=0.02, seed=20231017)
ctx = M.Context(0)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
out = {"ecs": aln.n_ecs}
for algo, code, storage in (("rcg", M.ALGO_RCG, M.STORE_F64), ("em", M.ALGO_EM, M.STORE_F64), ("rcg_sparse", M.ALGO_RCG, M.STORE_SPARSE),
                            ("em_sparse", M.ALGO_EM, M.STORE_SPARSE)):
    lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage)
    for tag in ("cold", "warm", "warm2"):
        ctx.sync(); t0 = time.perf_counter(); r = lik.vi_run(code); ctx.sync(); out[f"{algo}_{tag}_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    out[f"{algo}_iters"] = r.iters
    s = lik.vi_begin(code, tol=-1e300 if algo.startswith("rcg") else 0.0, max_iters=10**6)
    s.step(20); s.poll()
    for n in (100, 1000):
        l0 = M.launch_count(); t0 = time.perf_counter(); s.step(n); st = s.poll(); dt = time.perf_counter() - t0
        out[f"{algo}_us_per_iter_{n}"] = round(dt / n * 1e6, 2)
        out[f"{algo}_launches_per_iter"] = (M.launch_count() - l0) / n
    s.finish()
    lik.close()
print(json.dumps(out))
