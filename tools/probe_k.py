#!/usr/bin/env python
"""Sweep throughput across K (dense entry, synthetic logl): GB/s of algorithmic bytes per mode.  Not the bench."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M

ap = argparse.ArgumentParser()
ap.add_argument("--ks", default="8,16,30,50,64,100")
ap.add_argument("--bytes", type=float, default=2e9, help="fp64 logl bytes per case")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--modes", default="em64,em32,rcg")
a = ap.parse_args()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
ctx = M.Context(0)
rng = np.random.default_rng(1)
out = {}
for K in [int(x) for x in a.ks.split(",")]:
    N = int(a.bytes / 8 / K)
    logl = np.full((K, N), np.log(0.01))
    hit = rng.integers(0, K, size=(3, N))
    for h in hit:
        logl[h, np.arange(N)] = -rng.random(N) * 3
    lc = np.log(rng.integers(1, 20, N).astype(np.float64))
    for mode in a.modes.split(","):
        st = M.STORE_F32 if mode == "em32" else M.STORE_F64
        lik = M.Likelihood.from_dense(ctx, logl, lc, storage=st)
        s = lik.vi_begin(M.ALGO_RCG if mode == "rcg" else M.ALGO_EM, tol=-1e300 if mode == "rcg" else 0.0, max_iters=10**6, time_kernels=True)
        s.step(3); p0 = s.poll()
        l0 = M.launch_count(); t0 = time.perf_counter(); s.step(a.iters); p1 = s.poll(); wall = time.perf_counter() - t0
        launches = (M.launch_count() - l0) / max(1, p1.iters - p0.iters)
        ms = (p1.pass_ms_sum - p0.pass_ms_sum) / max(1, p1.iters - p0.iters)
        gbs = p1.pass_bytes / (ms * 1e-3) / 1e9
        out[f"{mode}_K{K}"] = round(gbs / peak, 3)
        print(f"K={K:5d} N={N:9d} {mode:5s} iters={p1.iters - p0.iters} kernels/iter={ms:.4f} ms wall/iter={wall / a.iters * 1e3:.4f} ms launches/iter={launches:.1f} "
              f"{gbs:7.0f} GB/s = {gbs / peak:.3f} of peak  bound={p1.bound:.6f}", flush=True)
        s.finish(); lik.close()
print(json.dumps(out))
