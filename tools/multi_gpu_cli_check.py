#!/usr/bin/env python
"""mSWEEP_b200 --gpus N against --gpus 1 on the same files: plain estimate (reads routed to the GPUs by pattern hash, every
GPU builds the classes of its own hash range, one all-reduce per pass), probabilities and read bins (class ids and read
ids mapped back to the input's), bootstrap (replicates spread over the GPUs), and the abort path (one GPU fails: the
others must not hang in a collective).  Needs >= 2 GPUs."""
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
# probabilities and read bins: the partitioned tables concatenate to the global table, read ids are the input's
os.makedirs(os.path.join(d, "bins1"), exist_ok=True); os.makedirs(os.path.join(d, "binsN"), exist_ok=True)
run("bins1/x", "--write-probs", "--bin-reads")
run("binsN/x", "--write-probs", "--bin-reads", "--gpus", str(n))
p1 = open(os.path.join(d, "bins1", "x_probs.tsv")).read().splitlines()
pN = open(os.path.join(d, "binsN", "x_probs.tsv")).read().splitlines()
assert p1[0] == pN[0] and len(p1) == len(pN)
for l1, lN in zip(p1[1:], pN[1:]):
    if not l1:
        continue
    c1, cN = l1.split("\t"), lN.split("\t")
    assert c1[0] == cN[0], "class ids"
    assert np.max(np.abs(np.array(c1[1:], float) - np.array(cN[1:], float))) < 2e-6
bins1 = sorted(f for f in os.listdir(os.path.join(d, "bins1")) if f.endswith(".bin"))
binsN = sorted(f for f in os.listdir(os.path.join(d, "binsN")) if f.endswith(".bin"))
assert bins1 == binsN and bins1
n_bin_diff = 0
for f in bins1:
    a = open(os.path.join(d, "bins1", f)).read().split()
    b = open(os.path.join(d, "binsN", f)).read().split()
    n_bin_diff += len(set(a) ^ set(b))
assert n_bin_diff <= 2, n_bin_diff          # (a class whose posterior sits within 1e-12 of the threshold may fall either way)
# one GPU fails: the process must end with the reference's message and exit code, not hang
t0 = time.time()
r = subprocess.run([cli, "--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-t", "8", "-o", os.path.join(d, "fail"), "--gpus", str(n)],
                   capture_output=True, text=True, timeout=120, env=dict(os.environ, MSWB_TEST_FAIL_GPU=str(n - 1)))
assert r.returncode == 1 and "injected failure" in r.stderr and "exiting" in r.stderr, (r.returncode, r.stderr[-500:])
t_fail = time.time() - t0
m1, _ = run("m1", "--min-hits", "500", "--algorithm", "emb200")
mN, _ = run("mN", "--min-hits", "500", "--algorithm", "emb200", "--gpus", str(n))
assert np.max(np.abs(m1 - mN)) < 2e-6
print(f"multi-GPU CLI ok on {n} GPUs: plain {t1:.2f}s -> {tN:.2f}s, bootstrap x8 {tb1:.2f}s -> {tbN:.2f}s (identical replicates), "
      f"abort path {t_fail:.2f}s")
