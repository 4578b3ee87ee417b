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
        if storage == M.STORE_SPARSE:
            lik.posteriors(0, min(50, lik.n_ecs))
            r = lik.vi_run(M.ALGO_RCG, max_iters=6, tol=-1e300)          # sparse RCG: separable state off the hits (fused launch)
            assert abs(r.theta.sum() - 1) < 1e-9
            os.environ["MSWB_FUSED"] = "0"                               # ... and one launch per sweep
            r2 = lik.vi_run(M.ALGO_RCG, max_iters=6, tol=-1e300)
            os.environ.pop("MSWB_FUSED")
            assert np.array_equal(r.theta, r2.theta)
            lik.posteriors(0, min(50, lik.n_ecs))
            with np.errstate(divide="ignore"):
                bins = lik.assign(aln, np.log(r.theta))
            lik.bootstrap_run(2, seed=3, max_iters=4)                     # device MT19937-64 stream + sparse RCG replicates
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
# dense entry over the sub-warp tile shapes (rows of 1..32 pieces), both tails (last-CTA reduction + control step, and
# finalize_ctl_kernel with MSWB_TAIL_MAX=0), fp64 and fp32 storage
for K, N in ((3, 70), (8, 300), (13, 129), (30, 257), (50, 700), (33, 129)):
    logl = rng.normal(-5, 2, size=(K, N)); lc = np.log(rng.integers(1, 9, size=N).astype(float))
    for tail_max in (None, "0"):
        if tail_max is not None:
            os.environ["MSWB_TAIL_MAX"] = tail_max
        lik = M.Likelihood.from_dense(ctx, logl, lc)
        lik.vi_run(M.ALGO_RCG, max_iters=5, tol=-1e300); lik.vi_run(M.ALGO_EM, max_iters=5, tol=0.0)
        lik.close()
        lik = M.Likelihood.from_dense(ctx, logl, lc, storage=M.STORE_F32)
        lik.vi_run(M.ALGO_EM, max_iters=5, tol=0.0)
        lik.close()
        os.environ.pop("MSWB_TAIL_MAX", None)
# bootstrap replicates as device batches (EM, dense): ragged last slice, fp64 and fp32, three row shapes
for K, T, R in ((6, 60, 500), (300, 1200, 700), (1100, 3300, 300)):
    wl = synth.generate(R, T, K, n_present=3, n_templates=30, p_noise=0.05, seed=K + 1)
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    for storage in (M.STORE_F64, M.STORE_F32):
        lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage)
        th, _ = lik.bootstrap_run(5, seed=3, algo=M.ALGO_EM, max_iters=6, tol=0.0)
        assert np.all(np.abs(th.sum(axis=1) - 1) < 1e-6)
        lik.close()
    aln.close()
# segment-parallel std::mt19937_64 (mt64_chain_kernel + mt64_segments_kernel), the fused cooperative kernels with the partial
# vectors reduced by every CTA between two grid rendezvous (MSWB_TAIL_MAX=0 forces that path), and the tiled finalize kernels
wl = synth.generate(3000, 300, 12, n_present=3, n_templates=200, p_noise=0.05, seed=9)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
os.environ["MSWB_MT_SEGMENTS"] = "4"
lik.bootstrap_resample(5, 3); lik.bootstrap_resample(5, 2, bootstrap_count=1000)
lik.bootstrap_run(2, seed=3, max_iters=4)
os.environ.pop("MSWB_MT_SEGMENTS")
os.environ["MSWB_TAIL_MAX"] = "0"
for fused in ("1", "0"):
    os.environ["MSWB_FUSED"] = fused
    r = lik.vi_run(M.ALGO_EM, max_iters=6, tol=0.0)
    assert abs(r.theta.sum() - 1) < 1e-9
    r = lik.vi_run(M.ALGO_RCG, max_iters=6, tol=-1e300)
    assert abs(r.theta.sum() - 1) < 1e-9
os.environ.pop("MSWB_FUSED"); os.environ.pop("MSWB_TAIL_MAX")
lik.close(); aln.close()
print("sanitize run ok, launches:", M.launch_count())
