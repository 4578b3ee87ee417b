#!/usr/bin/env python
"""Quick per-kernel timing probe (not the bench): builds a config-3-like shard and times each VI sweep."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=2000)
ap.add_argument("--s", type=int, default=30)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--modes", default="em64,em32,rcg")
ap.add_argument("--no-events", action="store_true")
a = ap.parse_args()
t0 = time.time()
wl = synth.generate_ec_patterns(a.n, a.k, a.s)
print(f"gen {time.time()-t0:.1f}s nnz={len(wl.targets)}", flush=True)
ctx = M.Context(0)
t0 = time.time(); aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets); print(f"ec_build {time.time()-t0:.3f}s n_ecs={aln.n_ecs}", flush=True)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for mode in a.modes.split(","):
    storage = M.STORE_F32 if mode == "em32" else (M.STORE_SPARSE if mode in ("emsp", "rcgsp") else M.STORE_F64)
    t0 = time.time(); lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage); ctx.sync()
    print(f"[{mode}] lik_build {time.time()-t0:.3f}s K={lik.n_groups} N={lik.n_ecs}", flush=True)
    algo = M.ALGO_RCG if mode in ("rcg", "rcgsp") else M.ALGO_EM
    s = lik.vi_begin(algo, tol=0.0, max_iters=10**6, time_kernels=not a.no_events)
    s.step(3); st = s.poll()
    ms0, n0 = st.pass_ms_sum, st.pass_launches
    t0 = time.time(); s.step(a.iters); st = s.poll(); wall = time.time() - t0
    ms = (st.pass_ms_sum - ms0); nl = st.pass_launches - n0
    per_iter_kernel_ms = max(ms / a.iters, 1e-9)
    gbs = st.pass_bytes / (per_iter_kernel_ms * 1e-3) / 1e9
    print(f"[{mode}] iters={st.iters} bound={st.bound:.6f} wall/iter={wall/a.iters*1e3:.3f} ms  pass-kernels/iter={per_iter_kernel_ms:.3f} ms "
          f"({nl//a.iters} launches)  algorithmic={st.pass_bytes/1e9:.2f} GB -> {gbs:.0f} GB/s = {gbs/peak:.3f} of measured peak", flush=True)
    r = s.finish()
    print(f"[{mode}] theta sum={r.theta.sum():.12f} top={np.sort(r.theta)[-3:]}", flush=True)
    lik.close()
