#!/usr/bin/env python
"""BASELINE config 4 shape: 5e7 classes x 10,000 lineages, most of them empty, --min-hits 1.  Builds the classes and the
likelihood on one GPU (hit tallies over all 10,000 groups, the mask, the compacted K' x N matrix), runs the optimiser on the
kept groups and checks the mask against a host tally of the same patterns."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--patterns", type=int, default=50_000_000)
ap.add_argument("--groups", type=int, default=10_000)
ap.add_argument("--group-size", type=int, default=6)
ap.add_argument("--pool", type=int, default=100)
ap.add_argument("--algos", default="rcg,em")
ap.add_argument("--max-iters", type=int, default=5000)
a = ap.parse_args()
t0 = time.time()
wl = synth.generate_ec_patterns(a.patterns, a.groups, a.group_size, n_present=50, n_pool=a.pool, seed=20231021)
t_gen = time.time() - t0
# host tally: which groups receive any hit at all (every class has count >= 1, so hits[g] >= 1 <=> some pattern hits g)
hit_groups = np.zeros(a.groups, bool)
hit_groups[np.unique(wl.group_of_target[np.unique(wl.targets)])] = True
ctx = M.Context(0)
t0 = time.time(); aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets); t_ec = time.time() - t0
t0 = time.time(); lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=1); ctx.sync(); t_lik = time.time() - t0
mask, hits = lik.mask(want_hits=True)
out = {"patterns": a.patterns, "groups": a.groups, "targets": wl.n_targets, "generate_s": round(t_gen, 1), "ecs": aln.n_ecs,
       "ec_build_s": round(t_ec, 3), "likelihood_s": round(t_lik, 3), "groups_kept": int(lik.n_groups),
       "mask_matches_host_tally": bool(np.array_equal(mask.astype(bool), hit_groups)),
       "hits_sum_over_groups_ge_aligned": bool(int(hits.sum()) >= aln.n_aligned)}
kept = np.flatnonzero(mask)
for algo in a.algos.split(","):
    t0 = time.time()
    r = lik.vi_run(M.ALGO_RCG if algo == "rcg" else M.ALGO_EM, tol=1e-6, max_iters=a.max_iters)
    dt = time.time() - t0
    out[algo] = {"seconds": round(dt, 3), "iters": int(r.iters), "converged": bool(r.converged), "ms_per_iter": round(dt / max(1, r.iters) * 1e3, 3),
                 "theta_sum": float(r.theta.sum()), "bound": float(r.bound),
                 "max_abs_err_vs_generating_theta": float(np.max(np.abs(r.theta - wl.truth[kept])))}
    if algo == "rcg":
        theta_rcg = r.theta
    elif "rcg" in out:
        out["max_abs_theta_diff_em_vs_rcg"] = float(np.max(np.abs(r.theta - theta_rcg)))
print(json.dumps(out))
