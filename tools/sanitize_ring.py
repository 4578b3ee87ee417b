#!/usr/bin/env python
"""Tiny RCG runs over the ring-fed sweep shapes for compute-sanitizer (racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M

ctx = M.Context(0)
rng = np.random.default_rng(0)
for K, N in ((100, 70), (300, 90), (1100, 40), (2100, 24)):
    logl = rng.normal(-5, 2, size=(K, N)); lc = np.log(rng.integers(1, 9, size=N).astype(float))
    lik = M.Likelihood.from_dense(ctx, logl, lc)
    r = lik.vi_run(M.ALGO_RCG, max_iters=3, tol=-1e300)
    assert abs(r.theta.sum() - 1) < 1e-9
    lik.close()
print("ring sanitize run ok, launches:", M.launch_count())
