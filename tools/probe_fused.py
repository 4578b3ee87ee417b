"""Where does a bootstrap replicate's time go with the fused cooperative kernels and the segmented generator on / off?
(config 2 / 5 size, 13 replicates = one GPU's share of config 5)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

wl = synth.generate_ec_patterns(1_000_000, 1000, 60, n_present=20, seed=55, dup_factor=9.0)
ctx = M.Context(0)
aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
out = {}
for rnd in range(2):
    for fused in ("1", "0"):
        for seg in ("16", "0"):
            os.environ["MSWB_FUSED"] = fused
            os.environ["MSWB_MT_SEGMENTS"] = seg
            for name, algo in (("rcg", M.ALGO_RCG), ("em", M.ALGO_EM)):
                ctx.sync(); t0 = time.perf_counter()
                th, it = lik.bootstrap_run(13, seed=11, algo=algo); ctx.sync(); t1 = time.perf_counter()
                out.setdefault(f"{name}_fused{fused}_seg{seg}", []).append(round((t1 - t0) * 1e3, 1))
print(json.dumps(out))
