#!/usr/bin/env python
"""Stage timing of the end-to-end leg of bench.py (host CSR buffers -> theta) at the bench's default shard."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msweep_b200 as M
from msweep_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12_500_000)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--storage", default="f32")
a = ap.parse_args()
wl = synth.generate_ec_patterns(a.n, 2000, 16, seed=20231019)
for arr in (wl.row_ptr, wl.targets, wl.group_of_target, wl.group_sizes):
    torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
ctx = M.Context(0)
storage = {"f32": M.STORE_F32, "f64": M.STORE_F64, "sparse": M.STORE_SPARSE}[a.storage]
for rep in range(3):
    ctx.sync(); t0 = time.perf_counter()
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets); ctx.sync(); t1 = time.perf_counter()
    lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage); ctx.sync(); t2 = time.perf_counter()
    r = lik.vi_run(M.ALGO_EM, tol=0.0, max_iters=a.steps, poll_every=a.steps); t3 = time.perf_counter()
    lik.close(); aln.close(); ctx.sync(); t4 = time.perf_counter()
    print(f"rep {rep}: ec_build {t1-t0:.3f} s (h2d {(wl.row_ptr.nbytes + wl.targets.nbytes)/1e9:.2f} GB)  lik_build {t2-t1:.3f} s  vi_run({a.steps}) {t3-t2:.3f} s  free {t4-t3:.3f} s  total {t4-t0:.3f} s", flush=True)
