#!/usr/bin/env python
"""BASELINE config 1 end to end: synthetic Themisto paired-end pseudoalignment, 1e6 reads x 3000 refs in 50 lineages,
text files in -> <prefix>_abundances.txt out.  Runs the oracle CLI (reference restatement, rcgcpu, -t 8) and
mSWEEP_b200 on the same files, compares the outputs and prints per-stage timings."""
import argparse, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from msweep_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--iters", type=int, default=0)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
t0 = time.time()
wl = synth.generate(a.reads, 3000, 50, n_present=5, n_templates=2000, p_noise=0.02, seed=20231017)
paths = synth.write_themisto(os.path.join(d, "aln"), wl, paired=True, shuffle_frac=0.01)
g = os.path.join(d, "grouping.txt")
synth.write_grouping(g, wl)
t_gen = time.time() - t0
common = ["--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-t", str(a.threads)]
if a.iters:
    common += ["--iters", str(a.iters), "--seed", "7"]
t0 = time.time()
r = subprocess.run([os.path.join(root, "oracle", "msweep_oracle"), *common, "-o", os.path.join(d, "ref"), "--algorithm", "rcgcpu", "--print-timings"], capture_output=True, text=True)
t_ref = time.time() - t0
assert r.returncode == 0, r.stderr
oracle_stages = json.loads(r.stderr.strip().splitlines()[-1])
t0 = time.time()
r = subprocess.run([os.path.join(root, "msweep_b200", "bin", "mSWEEP_b200"), *common, "-o", os.path.join(d, "ours"), "--print-timings"], capture_output=True, text=True)
t_ours = time.time() - t0
assert r.returncode == 0, r.stderr
stages = json.loads(r.stderr.strip().splitlines()[-1])
parse_stages = [l.strip() for l in r.stderr.splitlines() if l.startswith("  [parse]")]     # with MSWB_PARSE_TIMING=1


def vals(p):
    rows = [l.split("\t") for l in open(p).read().splitlines() if not l.startswith("#")]
    return [x[0] for x in rows], np.array([[float(v) for v in x[1:]] for x in rows])


n1, v1 = vals(os.path.join(d, "ours_abundances.txt"))
n2, v2 = vals(os.path.join(d, "ref_abundances.txt"))
h1 = [l for l in open(os.path.join(d, "ours_abundances.txt")).read().splitlines() if l.startswith("#")][1:]
h2 = [l for l in open(os.path.join(d, "ref_abundances.txt")).read().splitlines() if l.startswith("#")][1:]
print(json.dumps({"config": f"1: {a.reads} paired reads x 3000 refs / 50 lineages, rcg, -t {a.threads}, bootstrap iters {a.iters}",
                  "input_mb": sum(os.path.getsize(p) for p in paths) / 1e6, "generate_s": round(t_gen, 1),
                  "oracle_cli_s": round(t_ref, 2), "oracle_stages": oracle_stages, "host_cores": os.cpu_count(), "msweep_b200_cli_s": round(t_ours, 2), "stages": stages, "parse_stages": parse_stages,
                  "same_header": h1 == h2, "same_names": n1 == n2, "max_abs_theta_diff": float(np.max(np.abs(v1 - v2)))}))
