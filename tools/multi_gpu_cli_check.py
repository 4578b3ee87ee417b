#!/usr/bin/env python
"""mSWEEP_b200 --gpus N against --gpus 1 on the same files: plain estimate (classes sharded over the GPUs, one
all-reduce per pass) and bootstrap (replicates spread over the GPUs).  Needs >= 2 GPUs."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from msweep_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cli = os.path.join(root, "msweep_b200", "bin", "mSWEEP_b200")
d = tempfile.mkdtemp()
wl = synth.generate(60000, 3000, 50, n_present=5, n_templates=400, p_noise=0.02, seed=5)
paths = synth.write_themisto(os.path.join(d, "aln"), wl, paired=True)
g = os.path.join(d, "g.txt")
synth.write_grouping(g, wl)


def run(tag, *extra):
    t0 = time.time()
    r = subprocess.run([cli, "--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-t", "8", "-o", os.path.join(d, tag), *extra],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split("\t") for l in open(os.path.join(d, tag + "_abundances.txt")).read().splitlines() if not l.startswith("#")]
    return np.array([[float(v) for v in x[1:]] for x in rows]), time.time() - t0


a1, t1 = run("p1")
aN, tN = run("pN", "--gpus", str(n))
assert np.max(np.abs(a1 - aN)) < 2e-6, np.max(np.abs(a1 - aN))
b1, tb1 = run("b1", "--iters", "8", "--seed", "3")
bN, tbN = run("bN", "--iters", "8", "--seed", "3", "--gpus", str(n))
assert np.array_equal(b1, bN), np.max(np.abs(b1 - bN))
m1, _ = run("m1", "--min-hits", "500", "--algorithm", "emb200")
mN, _ = run("mN", "--min-hits", "500", "--algorithm", "emb200", "--gpus", str(n))
assert np.max(np.abs(m1 - mN)) < 2e-6
print(f"multi-GPU CLI ok on {n} GPUs: plain {t1:.2f}s -> {tN:.2f}s, bootstrap x8 {tb1:.2f}s -> {tbN:.2f}s (identical replicates)")
