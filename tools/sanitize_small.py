#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

ctx = M.Context(0)
for K, T, R in ((7, 70, 600), (300, 1500, 800), (1100, 3300, 300)):
    wl = synth.generate(R, T, K, n_present=3, n_templates=30, p_noise=0.05, seed=K)
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    for storage in (M.STORE_F64, M.STORE_F32, M.STORE_SPARSE):
        lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=1 if K == 7 else 0, storage=storage)
        r = lik.vi_run(M.ALGO_EM, max_iters=6, tol=0.0)
        assert abs(r.theta.sum() - 1) < 1e-6
        if storage == M.STORE_F64:
            r = lik.vi_run(M.ALGO_RCG, max_iters=6, tol=-1e300)
            assert abs(r.theta.sum() - 1) < 1e-9
            lik.posteriors(0, min(50, lik.n_ecs)); lik.export_logl(); lik.export_hit_counts()
            with np.errstate(divide="ignore"):
                bins = lik.assign(aln, np.log(r.theta))
            assert sum(b.size for b in bins) > 0
            lik.bootstrap_run(1, seed=3, max_iters=4)
            lik.bootstrap_resample(3, 1, rng_mode=M.RNG_PHILOX)
        lik.close()
    aln.close()
rng = np.random.default_rng(0)
logl = rng.normal(-5, 2, size=(33, 129)); lc = np.log(rng.integers(1, 9, size=129).astype(float))
lik = M.Likelihood.from_dense(ctx, logl, lc)
lik.vi_run(M.ALGO_RCG, max_iters=5, tol=-1e300); lik.vi_run(M.ALGO_EM, max_iters=5, tol=0.0)
print("sanitize run ok, launches:", M.launch_count())
