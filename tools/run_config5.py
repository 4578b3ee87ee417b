#!/usr/bin/env python
"""BASELINE config 5 shape: bootstrap on ~1e7 reads x 1000 lineages.  Times a few replicates on one GPU
(replicates are independent: N GPUs take r % N each) and checks one replicate's counts against the oracle stream."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--patterns", type=int, default=1_000_000)
ap.add_argument("--replicates", type=int, default=4)
ap.add_argument("--algo", default="rcg")
a = ap.parse_args()
wl = synth.generate_ec_patterns(a.patterns, 1000, 24, n_present=20, seed=55, dup_factor=9.0)
ctx = M.Context(0)
t0 = time.time(); aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets); t_ec = time.time() - t0
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
algo = M.ALGO_RCG if a.algo == "rcg" else M.ALGO_EM
t0 = time.time(); main = lik.vi_run(algo); t_main = time.time() - t0
t0 = time.time(); thetas, iters = lik.bootstrap_run(a.replicates, seed=11, algo=algo); t_boot = time.time() - t0
t0 = time.time(); counts = lik.bootstrap_resample(11, 1); t_res = time.time() - t0
from oracle import pyoracle as orc
ok = bool(np.array_equal(counts[0], orc.bootstrap_resample(aln.export().count, 11, 1)[0]))
print(json.dumps({"reads": wl.n_reads, "aligned": aln.n_aligned, "ecs": aln.n_ecs, "groups": lik.n_groups, "ec_build_s": round(t_ec, 3),
                  "main_estimate_s": round(t_main, 3), "main_iters": main.iters, "replicates": a.replicates,
                  "bootstrap_s": round(t_boot, 3), "s_per_replicate": round(t_boot / a.replicates, 3), "replicate_iters": iters,
                  "resample_only_s": round(t_res, 3), "counts_bit_exact_vs_oracle": ok,
                  "theta_spread_top": float(np.std(thetas[:, int(np.argmax(main.theta))]))}))
